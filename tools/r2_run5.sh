#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -x -k "fused_groupnorm_input or halo_concat" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; tail -2 gpurun_out/r2_bench6.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench6.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq", d["hq_train"]["ms_per_step"], d["hq_ddim100"]["ms_per_step"], d["roofline"]["achieved"])
PY
DDPM_NO_XF=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench6_noxf.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench6_noxf.json")); print("NO_XF: ddim50", d["sampler"]["ddim50"]["ms_per_step"], "hq_ddim", d["hq_ddim100"]["ms_per_step"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd256_c.csv python tools/profile_step.py fwd 256 > gpurun_out/r2_ncu_fwd_c.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_fwd256_c.csv 14
