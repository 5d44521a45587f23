#!/bin/bash
# refresh the committed artefacts under profiles/ from the current build (run under gpurun, 1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; tail -1 gpurun_out/bench_n1_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_arm.json 2> gpurun_out/bench_ref_arm.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_final.csv python tools/profile_step.py train 128 > gpurun_out/ncu_train_final.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd256_final.csv python tools/profile_step.py fwd 256 > gpurun_out/ncu_fwd_final.log 2>&1
timeout 300 python tools/op_timing.py train 128 > gpurun_out/op_timing_train_summary.txt 2>&1
timeout 600 python tools/time_hq.py > gpurun_out/hq_timing.json 2> gpurun_out/hq_timing.err
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel" -c 2 -o gpurun_out/prof_halo_final python tools/profile_step.py train 128 > gpurun_out/ncu_halo_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_full.json")); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"], d["cpu_baseline"], d.get("sampler"), d["clocks"], d["with_optimizer"])
print(open("gpurun_out/bench_ref_arm.json").read()[:300])
print(open("gpurun_out/hq_timing.json").read()[:600])
PY
