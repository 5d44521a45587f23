#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench11.json")); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dominant", d["roofline"]["achieved"])
PY
timeout 600 python tools/time_hq.py 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd256_b.csv python tools/profile_step.py fwd 256 > gpurun_out/ncu_fwd_b.log 2>&1
