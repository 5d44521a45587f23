#!/bin/bash
# GPU regression: UNet parity tests, a short training-step bench line and the sampler ms/step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_short.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "with_opt", round(d["with_optimizer"]["ms_per_step"],3))
PY
tail -2 gpurun_out/bench_short.err
timeout 300 python tools/sampler_ab.py 300 2>&1 | tail -1
