#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -x -s --timeout 300 2>&1 | tail -80 > gpurun_out/unet_tests.txt
tail -80 gpurun_out/unet_tests.txt
