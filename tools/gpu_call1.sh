#!/bin/bash
# first GPU call: environment facts, engine parity tests, stock-torch baseline
mkdir -p gpurun_out
{ nvidia-smi; ls /root/reference 2>&1 | head -3; nproc; free -g | head -2; python -c "import matplotlib" 2>&1 | tail -1; } > gpurun_out/env.txt 2>&1
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 2>&1 | tail -40 > gpurun_out/gemm_tests.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py -q --timeout 120 2>&1 | tail -60 > gpurun_out/gemm_tests_all.txt
timeout 900 python tests/ref_cuda_timing.py > gpurun_out/ref_cuda.log 2>&1
tail -5 gpurun_out/gemm_tests.txt; tail -30 gpurun_out/gemm_tests_all.txt; tail -3 gpurun_out/ref_cuda.log
