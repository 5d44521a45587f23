#!/bin/bash
export DDPM_GEMM_CLUSTER=1
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -k "halo" 2>&1 | tail -2
for d in 0 1 2 4 3 5 6 7; do DDPM_GEMM_DBG=$d python tools/dbg_dominant.py 2>&1 | tail -1; done
