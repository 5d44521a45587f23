#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s --timeout 300 > gpurun_out/unet_tests.txt 2>&1
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/unet_tests.txt | head -80
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
