#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 2>&1 | tail -2
python tools/dbg_dominant.py 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench8.json")); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dominant", d["roofline"]["achieved"])
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train8.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train8.log 2>&1
