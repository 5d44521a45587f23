#!/bin/bash
mkdir -p gpurun_out
for v in "X=0" "DDPM_GN_BWD_MINB=3" "DDPM_GN_BWD_MINB=4" "DDPM_GN_BWD_MINB2=3" "DDPM_GN_BWD_MINB2=4" "DDPM_GN_BWD_MINB=3 DDPM_GN_BWD_MINB2=3" "DDPM_GN_NO_DN=1" "DDPM_GN_BWD_OCC=2" "DDPM_GN_BWD_OCC=3" "DDPM_GN_BWD_OCC=6" "X=0"; do
  echo "== $v"
  env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq --no-sampler > gpurun_out/r2_bench5_ab.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5_ab.json")); print("BENCH ms/step", round(d["ms_per_step"],3))
PY
done
