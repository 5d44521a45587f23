#!/bin/bash
mkdir -p gpurun_out
for cl in 2 4; do
  export DDPM_GEMM_CLUSTER=$cl
  echo "== cluster $cl"; timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 2>&1 | tail -3
done
export DDPM_GEMM_CLUSTER=2
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s --timeout 300 > gpurun_out/unet_tests.txt 2>&1
grep -E "flat grad|eps rel|passed|failed|FAILED|Error" gpurun_out/unet_tests.txt | head -20
for cl in 1 2 4; do
  export DDPM_GEMM_CLUSTER=$cl
  timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench_cl$cl.json 2> gpurun_out/bench5.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cl$cl.json")); print("cluster $cl", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dominant TF/s", d["roofline"]["achieved"])
PY
done
tail -3 gpurun_out/bench5.err
