#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key counters per captured launch."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("smsp__inst_executed.sum", "warp insts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "ld sector use %"),
        ("derived__memory_l1_wavefronts_shared_excessive", "excess smem wavefronts")]
for r in rows[2:]:
    print("-" * 60)
    for k, nm in want:
        if k in idx:
            print(f"{nm:28s} {r[idx[k]][:90]} {units[idx[k]]}")
    st = [(float(r[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st.sort(reverse=True)
    print("top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))
