#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim.py -q --timeout 300 -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench30.json 2> gpurun_out/bench30.err
tail -3 gpurun_out/bench30.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench30.json")); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["with_optimizer"], d["clocks"])
PY
