#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_tests8.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests8.txt
grep -E "FAILED|Error" gpurun_out/r2_tests8.txt | head
for v in "X=0" "DDPM_NO_UPFOLD=1" "X=0" "DDPM_NO_UPFOLD=1"; do
  env $v timeout 600 python bench.py --no-cpu-baseline --no-stock > gpurun_out/r2_bench8.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - "$v" <<PY
import json,sys
d=json.load(open("gpurun_out/r2_bench8.json")); print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "ddim50", round(d["sampler"]["ddim50"]["ms_per_step"],3), "anc", round(d["sampler"]["ancestral1000"]["ms_per_step"],3), "hq", round(d["hq_train"]["ms_per_step"],3), round(d["hq_ddim100"]["ms_per_step"],3))
PY
done
