#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_tests6.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests6.txt
grep -E "FAILED|Error" gpurun_out/r2_tests6.txt | head
for v in "X=0" "DDPM_PACK_INLINE=1" "DDPM_NO_TC_INCONV=1" "X=0"; do
  echo "== $v"
  env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq > gpurun_out/r2_bench6_ab.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench6_ab.json")); print("BENCH ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "ddim50", round(d["sampler"]["ddim50"]["ms_per_step"],3), "anc", round(d["sampler"]["ancestral1000"]["ms_per_step"],3))
PY
done
