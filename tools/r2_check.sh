#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "halo_pair or pairmode" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported" | sort | uniq -c | head -8
bash tools/r2_ab_precise.sh X=0 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ms/step', round(d['ms_per_step'],3), 'ddim50', round(d['sampler']['ddim50']['ms_per_step'],3))"
