#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-sampler > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n2b.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n2b.json")); print("N=2", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "hq_train", d["hq_train"]["ms_per_step"], d["hq_train"]["images_per_s"])
PY
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 20 --warmup 5 --no-sampler --no-cpu-baseline --no-stock > gpurun_out/r2_bench_n1b.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n1b.json")); print("N=1", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "hq_train", d["hq_train"]["ms_per_step"], d["hq_train"]["images_per_s"])
PY
