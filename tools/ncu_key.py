#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep:  python tools/ncu_key.py file.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = [("time", "gpu__time_duration.sum"), ("dram rd", "dram__bytes_read.sum"), ("dram wr", "dram__bytes_write.sum"),
        ("dram %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue act %", "smsp__issue_active.avg.pct"),
        ("inst", "smsp__inst_executed.sum"), ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("l1tex %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("lts %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("dyn smem", "launch__shared_mem_per_block_dynamic")]
stalls = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90])
    for lab, m in want:
        if m in h:
            print(f"   {lab:16s} {r[h.index(m)]} {rows[1][h.index(m)]}")
    st = sorted(((float(r[h.index(c)] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stalls), reverse=True)
    print("   stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
