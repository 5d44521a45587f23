#!/bin/bash
# round-2 evidence: tests, full bench lines, launch lists, op timeline, per-kernel ncu --set full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_env.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/r02_gpu_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.txt; tail -3 gpurun_out/r02_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_full.json 2> gpurun_out/r02_bench_n1_full.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_bs128.csv python tools/profile_step.py train 128 > gpurun_out/ncu_l1.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_train_bs128.csv 30 > gpurun_out/r02_launches_train_bs128_summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fwd_bs256.csv python tools/profile_step.py fwd 256 > gpurun_out/ncu_l2.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_fwd_bs256.csv 30 > gpurun_out/r02_launches_fwd_bs256_summary.txt
timeout 300 python tools/op_timing.py train 128 > gpurun_out/r02_op_timeline_train_bs128.txt 2>&1
timeout 300 python tools/op_timing.py fwd 256 > gpurun_out/r02_op_timeline_fwd_bs256.txt 2>&1
for k in conv3x3_halo2_kernel "umma_gemm_kernel<256, 1" "umma_gemm_kernel<256, 0" k_gn_bwd_reduce k_gn_bwd_apply k_gn_apply k_adam_ema k_out_gather k_softmax_rows; do
  f=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$k" -c 2 -o gpurun_out/r02_ncu_$f python tools/prof_kernels.py > gpurun_out/ncu_full_$f.log 2>&1
  echo "ncu $k rc=$?"
done
python tools/prof_halo_epi.py plain plain_pair > gpurun_out/r02_halo_pair_vs_single.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
