#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -x -k "fused_groupnorm_input" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench7.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench7.json")); print("XF: ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq_ddim", d["hq_ddim100"]["ms_per_step"])
PY
DDPM_NO_XF=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench7_noxf.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench7_noxf.json")); print("NO_XF: ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq_ddim", d["hq_ddim100"]["ms_per_step"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd256_d.csv python tools/profile_step.py fwd 256 > gpurun_out/r2_ncu_fwd_d.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_fwd256_d.csv 8
