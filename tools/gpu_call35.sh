#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench35.json 2> gpurun_out/bench35.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench35.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "opt", round(d["with_optimizer"]["ms_per_step"],3))
PY
done
