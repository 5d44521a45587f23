#!/bin/bash
# compute-sanitizer over the -m gpu suite: memcheck on everything, racecheck (shared-memory hazards) on the operator tests and a
# small network pass.  Bounded by timeouts: the oracle's torch kernels run instrumented too.
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -q -m gpu -x > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r02_sanitizer_memcheck.txt | head -20
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gemm_gpu.py tests/test_bench_shapes_gpu.py -q -m gpu -x -k "attention or quad or celeba64 or optin" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/r02_sanitizer_racecheck.txt | head -20
