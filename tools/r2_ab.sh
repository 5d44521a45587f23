#!/bin/bash
# A/B of the CTA-pair modes on the whole training step
for m in 0 1 2 3; do
  DDPM_GEMM_PAIR=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock --no-hq --no-sampler > gpurun_out/ab_$m.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/ab_$m.json')); print('GEMM_PAIR=$m', {k:round(d[k],3) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']))
"
done
DDPM_HALO_PAIR=0 DDPM_GEMM_PAIR=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock --no-hq --no-sampler > gpurun_out/ab_nopair.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/ab_nopair.json')); print('no pair at all', {k:round(d[k],3) for k in ('value','ms_per_step')})
"
timeout 300 python tools/op_timing.py train 128 > gpurun_out/r2_op_timing_train5.txt 2>&1; head -12 gpurun_out/r2_op_timing_train5.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b.csv python tools/profile_step.py train 128 > gpurun_out/r2_ncu_train_b.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_train_b.csv 14
