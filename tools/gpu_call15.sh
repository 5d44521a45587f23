#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 2500 gpurun_out/bench_full.json; echo; tail -3 gpurun_out/bench_full.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
