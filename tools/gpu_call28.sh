#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -3
for f in pdl nopdl; do
  if [ $f = nopdl ]; then export DDPM_NO_PDL=1; else unset DDPM_NO_PDL; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench28_$f.json 2> gpurun_out/bench28.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench28_$f.json")); print("$f", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d.get("sampler"))
PY
done
tail -2 gpurun_out/bench28.err
timeout 900 python -m pytest tests/test_gemm_gpu.py -q --timeout 300 -x 2>&1 | tail -3
