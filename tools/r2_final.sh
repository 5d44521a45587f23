#!/bin/bash
# final refresh of the two driver-facing artefacts at the last commit: -m gpu test log and the N=1 bench line (+ reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/r02_gpu_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.txt; tail -3 gpurun_out/r02_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_full.json 2> gpurun_out/r02_bench_n1_full.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n1_full.json")); print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), "stock", round(d["vs_stock_cuda"]["train_speedup_vs_tf32_default"],2), "ddim50", round(d["sampler"]["ddim50"]["ms_per_step"],3), "anc", round(d["sampler"]["ancestral1000"]["ms_per_step"],3), "hq", round(d["hq_train"]["ms_per_step"],2), round(d["hq_ddim100"]["ms_per_step"],3))
PY
