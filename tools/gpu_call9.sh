#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -k "split_k or plain" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s --timeout 300 > gpurun_out/unet_tests.txt 2>&1
grep -E "flat grad|eps rel|passed|failed|FAILED|Error" gpurun_out/unet_tests.txt | head -20
for side in on off; do
  if [ $side = off ]; then export DDPM_NO_SIDE_STREAM=1; else unset DDPM_NO_SIDE_STREAM; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench_side_$side.json 2> gpurun_out/bench4.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_side_$side.json")); print("side $side", {k:d[k] for k in ("value","ms_per_step","launches_per_step")}, "e2e", d["e2e"]["value"], "dominant TF/s", d["roofline"]["achieved"])
PY
done
unset DDPM_NO_SIDE_STREAM
tail -3 gpurun_out/bench4.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train6.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train6.log 2>&1
wc -l gpurun_out/launches_train6.csv
