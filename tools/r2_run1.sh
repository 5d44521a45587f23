#!/bin/bash
# round-2 baseline: all GPU tests (incl. the new benchmarked-shape parity tests), the full bench line, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_env.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/r2_tests1.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.txt
tail -5 gpurun_out/r2_tests1.txt
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/r2_bench1_ref.err
head -c 1500 gpurun_out/r2_bench1.json; echo; head -c 600 gpurun_out/r2_bench1_ref.json
