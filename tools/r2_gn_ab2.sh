#!/bin/bash
mkdir -p gpurun_out
for v in "DDPM_GN_ORDER=0" "DDPM_GN_ORDER=1" "DDPM_GN_ORDER=2" "DDPM_GN_ORDER=4" "DDPM_GN_ORDER=3" "DDPM_GN_ORDER=7" "DDPM_GN_ORDER=0"; do
  echo "== $v"
  env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq > gpurun_out/r2_bench5_ab.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5_ab.json")); print("BENCH ms/step", round(d["ms_per_step"],3), "ddim50", round(d["sampler"]["ddim50"]["ms_per_step"],3))
PY
done
