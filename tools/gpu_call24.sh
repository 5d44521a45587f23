#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -2
for f in side main; do
  if [ $f = main ]; then export DDPM_WGRAD_MAIN=1; else unset DDPM_WGRAD_MAIN; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-sampler --no-cpu-baseline > gpurun_out/bench_wg_$f.json 2> gpurun_out/bench14.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_wg_$f.json")); print("wgrad on $f", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"])
PY
done
tail -2 gpurun_out/bench14.err
