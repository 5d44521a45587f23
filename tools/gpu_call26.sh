#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -2
for f in dn nodn; do
  if [ $f = nodn ]; then export DDPM_GN_NO_DN=1; else unset DDPM_GN_NO_DN; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench26_$f.json 2> gpurun_out/bench26.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench26_$f.json")); print("$f", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"])
PY
done
tail -2 gpurun_out/bench26.err
