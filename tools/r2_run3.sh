#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_tests3.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests3.txt
grep -E "FAILED|Error" gpurun_out/r2_tests3.txt | head
timeout 900 python bench.py --no-cpu-baseline --no-stock > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench4.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench4.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "launches", d["launches_per_step"], "ddim50 ms/step", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq", d["hq_train"]["ms_per_step"], d["hq_ddim100"]["ms_per_step"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd256_b.csv python tools/profile_step.py fwd 256 > gpurun_out/r2_ncu_fwd_b.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_fwd256_b.csv 16
timeout 300 python tools/op_timing.py train 128 > gpurun_out/r2_op_timing_train4.txt 2>&1; head -24 gpurun_out/r2_op_timing_train4.txt
