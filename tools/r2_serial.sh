#!/bin/bash
mkdir -p gpurun_out
for v in "DDPM_GN_BWD_NO_SLICE=1" "DDPM_GN_BWD_NP=4"; do
  echo "== serial (no side stream) $v"
  env $v DDPM_NO_SIDE_STREAM=1 timeout 300 python tools/op_timing.py train 128 > "gpurun_out/serial_$v.txt" 2>&1
  head -45 "gpurun_out/serial_$v.txt"
  cp gpurun_out/op_timing_train.txt "gpurun_out/serial_raw_$v.txt"
done
