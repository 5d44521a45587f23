#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 -s -k "halo" 2>&1 | grep -E "halo probe|passed|failed|Error|assert|FAILED" | head -40
