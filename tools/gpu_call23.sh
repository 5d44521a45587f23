#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s --timeout 300 > gpurun_out/unet_tests.txt 2>&1
grep -E "flat grad|eps rel|dropout grad|passed|failed|FAILED|Error" gpurun_out/unet_tests.txt | head -20
for f in on off; do
  if [ $f = off ]; then export DDPM_NO_FUSED_GN=1; else unset DDPM_NO_FUSED_GN; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench_fgn_$f.json 2> gpurun_out/bench13.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_fgn_$f.json")); print("fused_gn $f", {k:d[k] for k in ("value","ms_per_step","launches_per_step")}, "e2e", d["e2e"]["value"])
PY
done
unset DDPM_NO_FUSED_GN
tail -2 gpurun_out/bench13.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train12.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train12.log 2>&1
