#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_gn_bwd_apply|k_gn_bwd_reduce|k_in_conv|k_gn_stats" -s 60 -c 8 -o gpurun_out/prof_gn2 python tools/profile_step.py train 128 > gpurun_out/ncu_gn2.log 2>&1
ls -la gpurun_out/prof_gn2.ncu-rep; tail -2 gpurun_out/ncu_gn2.log
