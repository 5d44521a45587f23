#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_shapes_gpu.py tests/test_unet_gpu.py -q -m gpu -x 2>&1 | tail -3
for v in "DDPM_GN_BWD_NP=4" "DDPM_GN_BWD_NP=8" "DDPM_GN_BWD_NO_SLICE=1"; do
  echo "== $v"
  env $v timeout 300 python tools/op_timing.py train 128 > gpurun_out/gn_ab_$v.txt 2>&1
  head -3 gpurun_out/gn_ab_$v.txt | tail -2
  cp gpurun_out/op_timing_train.txt "gpurun_out/gn_ab_raw_$v.txt"
  grep -E "gn_bwd" "gpurun_out/gn_ab_raw_$v.txt" | awk '{s+=$2} END {print "sum gn_bwd", s}'
  grep -E "upsamples.level_0.0.norm1|upsamples.level_0.0.norm2|downsamples.level_0.0.norm1|upsamples.level_1.0.0.norm2|downsamples.level_1.0.0.norm1|upsamples.level_2.0.norm1.gn|downsamples.level_2.0.norm1.gn|middle.0.norm1.gn|downsamples.level_3.0.norm2.gn" "gpurun_out/gn_ab_raw_$v.txt" | grep gn_bwd
  env $v timeout 600 python bench.py --no-cpu-baseline --no-stock --no-hq > gpurun_out/r2_bench5_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench5_ab.json")); print("BENCH ms/step", round(d["ms_per_step"],3), "ddim50", d["sampler"]["ddim50"]["ms_per_step"])
PY
done
