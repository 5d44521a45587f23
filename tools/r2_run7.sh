#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -x -s -k "attention_fused" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench8.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench8.json")); print("fused attn: train", round(d["ms_per_step"],3), "ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq_ddim", d["hq_ddim100"]["ms_per_step"])
PY
DDPM_NO_FUSED_ATTN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench8_noattn.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench8_noattn.json")); print("unfused attn: train", round(d["ms_per_step"],3), "ddim50", d["sampler"]["ddim50"]["ms_per_step"], "anc", d["sampler"]["ancestral1000"]["ms_per_step"], "hq_ddim", d["hq_ddim100"]["ms_per_step"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd256_e.csv python tools/profile_step.py fwd 256 > gpurun_out/r2_ncu_fwd_e.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_fwd256_e.csv 12
