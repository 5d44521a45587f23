#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench33.json 2> gpurun_out/bench33.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench33.json")); print("$1", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]))
PY
}
for occ in 3 1; do export DDPM_GN_BWD_OCC=$occ; run "bwd_occ=$occ"; done
export DDPM_GN_BWD_OCC=4
for occ in 2 3 4; do export DDPM_GN_STATS_OCC=$occ; run "bwd4 stats_occ=$occ"; done
unset DDPM_GN_STATS_OCC
for occ in 2 3; do export DDPM_GN_APPLY_OCC=$occ; run "bwd4 apply_occ=$occ"; done
