"""Text summary of an `ncu --set full` report: python tools/ncu_summary.py <file.ncu-rep> [...]  (runs `ncu -i ... --page raw --csv`)."""
import csv, io, subprocess, sys
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__cluster_dim_x", "cluster x"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
        ("smsp__inst_executed.sum", "warp instructions"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"), ("lts__t_bytes.sum", "L2 traffic bytes"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %")]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(path, ": no data"); continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("==", path)
    for r in rows[2:]:
        name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
        print("kernel:", name[:150])
        for k, label in KEYS:
            if k in col:
                print("   %-48s %s %s" % (label, r[col[k]], units[col[k]]))
        print()
