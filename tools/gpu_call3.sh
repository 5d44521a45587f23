#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -q --timeout 300 -k "philox or dropout" 2>&1 | tail -5
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd256.csv python tools/profile_step.py fwd 256 > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/ncu_train.log
wc -l gpurun_out/launches_train.csv gpurun_out/launches_fwd256.csv
timeout 300 python tools/cpu_threads.py 2>&1 | tail -12
