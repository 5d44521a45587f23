#!/bin/bash
# usage: r2_ab_precise.sh "ENV_A" "ENV_B" ... ; alternates variants, 4 rounds, 100 timed steps each
mkdir -p gpurun_out
for r in 1 2 3 4; do
  for v in "$@"; do
    env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq --no-sampler --steps 100 --warmup 10 > gpurun_out/ab_p.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python - "$v" <<PY
import json,sys
d=json.load(open("gpurun_out/ab_p.json")); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4))
PY
  done
done
