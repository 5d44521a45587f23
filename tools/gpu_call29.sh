#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -q --timeout 300 -x -k "sampler_steps_vs_golden and tiny" -s 2>&1 | grep -v "^$" | tail -30
echo ---- NO_PDL
DDPM_NO_PDL=1 timeout 600 python -m pytest tests/test_unet_gpu.py -q --timeout 300 -k "sampler_steps_vs_golden" 2>&1 | tail -3
