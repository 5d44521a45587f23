#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_gn_bwd_apply|k_gn_bwd_reduce|k_gn_apply|k_gn_stats" -c 12 -o gpurun_out/prof_gn python tools/profile_step.py train 128 > gpurun_out/ncu_gn.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"umma_gemm" -c 10 -o gpurun_out/prof_umma python tools/profile_step.py train 128 > gpurun_out/ncu_umma.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_gn.log
