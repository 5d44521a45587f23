#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_tests7.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests7.txt
grep -E "FAILED|Error" gpurun_out/r2_tests7.txt | head
bash tools/r2_ab_precise.sh X=0 DDPM_FUSED_ATTN_BWD=1
timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq > gpurun_out/r2_bench7.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench7.json")); print("BENCH ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "ddim50", round(d["sampler"]["ddim50"]["ms_per_step"],3), "anc", round(d["sampler"]["ancestral1000"]["ms_per_step"],3))
PY
