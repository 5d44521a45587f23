#!/bin/bash
# GN-epilogue fusion check: unit tests of the new epilogues, whole-network parity, short bench, op timeline, launch lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -q --timeout 600 -k "gn or quad or halo_concat or plain_gemm" > gpurun_out/r2_t2_gemm.txt 2>&1; echo "gemm rc=$?"; tail -3 gpurun_out/r2_t2_gemm.txt
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -q --timeout 600 -s > gpurun_out/r2_t2_unet.txt 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/r2_t2_unet.txt
grep -E "FAILED|Error|assert" gpurun_out/r2_t2_gemm.txt gpurun_out/r2_t2_unet.txt | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hq --no-stock > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench2.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "launches", d["launches_per_step"], "ddim50 ms/step", d["sampler"]["ddim50"]["ms_per_step"], "roofline", d["roofline"]["achieved"])
PY
DDPM_NO_GN_EPI=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hq --no-stock --no-sampler > gpurun_out/r2_bench2_noepi.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench2_noepi.json")); print("NO_GN_EPI", {k:round(d[k],3) for k in ("value","ms_per_step")})
PY
timeout 300 python tools/op_timing.py train 128 > gpurun_out/r2_op_timing_train.txt 2>&1; head -30 gpurun_out/r2_op_timing_train.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python tools/profile_step.py train 128 > gpurun_out/r2_ncu_train.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_train.csv 24
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fwd256.csv python tools/profile_step.py fwd 256 > gpurun_out/r2_ncu_fwd.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches_fwd256.csv 16
# host-thread sweep of the reference arm at bs=128 (one step each)
python - <<PY
import time, torch, sys
sys.path.insert(0, ".")
import bench
for th in (16, 32, 64):
    bench.CPU_THREADS = th
    step, kind, what = bench.reference_cpu_step(128)
    step(); t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
    print("cpu threads", th, "bs128 step s", round(dt, 2), "img/s", round(128 / dt, 1), kind, flush=True)
PY
