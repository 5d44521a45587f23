#!/bin/bash
mkdir -p gpurun_out
export DDPM_GEMM_CLUSTER=1
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -s -k "halo" 2>&1 | grep -E "halo probe|passed|failed|Error|assert" | head -20
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s --timeout 300 > gpurun_out/unet_tests.txt 2>&1
grep -E "flat grad|eps rel|passed|failed|FAILED|Error" gpurun_out/unet_tests.txt | head -20
for halo in on off; do
  if [ $halo = off ]; then export DDPM_NO_HALO=1; else unset DDPM_NO_HALO; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench_halo_$halo.json 2> gpurun_out/bench6.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_halo_$halo.json")); print("halo $halo", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"])
PY
done
unset DDPM_NO_HALO
tail -3 gpurun_out/bench6.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train7.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train7.log 2>&1
