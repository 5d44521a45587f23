#!/bin/bash
# N=2 (or N=$1) bench exactly as the driver launches it, then N=1 on the same box for the efficiency ratio
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo rc=$?; tail -3 gpurun_out/r02_bench_n$N.err
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-stock > gpurun_out/r02_bench_n1_samebox.json 2>/dev/null
python - $N <<PY
import json,sys
N=sys.argv[1]
for f in (f"gpurun_out/r02_bench_n{N}.json","gpurun_out/r02_bench_n1_samebox.json"):
    d=json.load(open(f)); print(f, {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ddim50", round(d["sampler"]["ddim50"]["images_per_s"]), "hq_train", round(d["hq_train"]["ms_per_step"],3), round(d["hq_train"]["images_per_s"],1), "hq_ddim", round(d["hq_ddim100"]["images_per_s"],2))
PY
