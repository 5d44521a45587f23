#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 2>&1 | tail -2
python tools/dbg_dominant.py 2>&1 | tail -6
DDPM_GEMM_SHALLOW=1 python tools/dbg_dominant.py 2>&1 | head -1
for d in 1 2 4 7; do DDPM_GEMM_DBG=$d python tools/dbg_dominant.py 2>&1 | tail -1; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-sampler --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench7.json")); print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "dominant", d["roofline"]["achieved"])
PY
