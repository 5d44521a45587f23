#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench12.json")); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dominant", d["roofline"]["achieved"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train11.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train11.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd256_c.csv python tools/profile_step.py fwd 256 > gpurun_out/ncu_fwd_c.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel" -c 4 -o gpurun_out/prof_halo python tools/profile_step.py train 128 > gpurun_out/ncu_halo.log 2>&1
ls -la gpurun_out/prof_halo.ncu-rep
