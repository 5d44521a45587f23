#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench25.json 2> gpurun_out/bench25.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench25.json")); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"])
PY
tail -2 gpurun_out/bench25.err
timeout 300 python tools/op_timing.py train 128 2>&1 | tail -75
