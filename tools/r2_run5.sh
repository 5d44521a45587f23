#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_tests5.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests5.txt
grep -E "FAILED|Error" gpurun_out/r2_tests5.txt | head
timeout 900 python bench.py --no-cpu-baseline --no-stock > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench5.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "launches", d["launches_per_step"], "ddim50 ms/step", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq", d["hq_train"]["ms_per_step"], d["hq_ddim100"]["ms_per_step"])
PY
for v in "DDPM_GN_BWD_NO_CLUSTER=1" "DDPM_NO_ATTN16=1" "DDPM_SPLIT_MAX_TILES=74"; do
  echo "== $v"; env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq > gpurun_out/r2_bench5_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5_ab.json")); print(round(d["ms_per_step"],3), "ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"])
PY
done
timeout 300 python tools/op_timing.py train 128 > gpurun_out/r2_op_timing_train5.txt 2>&1; head -30 gpurun_out/r2_op_timing_train5.txt
