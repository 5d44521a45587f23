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
# (gpurun merges at most 64 MiB back: reports are summarised on the box and only the dominant kernel's .ncu-rep is kept)
for spec in "conv3x3_halo2_kernel:2" "umma_gemm_kernel:6" "attn_kernel:2" "k_attn16:2" "k_gn_bwd_reduce:1" "k_gn_bwd_apply:2" "k_gn_apply:1" "k_adam_ema:1"; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$k" -c $c -f -o gpurun_out/r02_ncu_$k python tools/prof_kernels.py > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_ncu_$k.ncu-rep > gpurun_out/r02_ncu_summary_$k.txt 2>&1
  if [ "$k" != "conv3x3_halo2_kernel" ]; then rm -f gpurun_out/r02_ncu_$k.ncu-rep; fi
  rm -f gpurun_out/ncu_full_$k.log
done
du -sh gpurun_out
python tools/prof_halo_epi.py plain plain_pair > gpurun_out/r02_halo_pair_vs_single.txt 2>&1
python tools/bench_wgrad.py > gpurun_out/r02_wgrad_microbench.txt 2>&1
python tools/cpu_enqueue.py > gpurun_out/r02_host_enqueue.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
