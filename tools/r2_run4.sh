#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -x 2>&1 | tail -6
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -2 gpurun_out/r2_bench5.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ddim50", d["sampler"]["ddim50"]["ms_per_step"], "hq", d["hq_train"]["ms_per_step"], d["hq_ddim100"]["ms_per_step"], d["roofline"]["achieved"])
PY
