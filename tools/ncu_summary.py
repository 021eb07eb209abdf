#!/usr/bin/env python
"""Summarise one gpurun_out/<tag>/ directory (tools/gpu_check.sh) into profiles/<tag>_*.

    python tools/ncu_summary.py gpurun_out/r1a r1a [bench-workload-key]

Writes profiles/<tag>_launches.md (per-kernel share of one step from the ncu launch list),
profiles/<tag>_kernels.md (ncu --set full metrics of the hand-written kernels) and updates
profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
wkey = sys.argv[3] if len(sys.argv) > 3 else "C4-f32"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
os.makedirs(prof, exist_ok=True)


def short(name):
    name = name.replace("void ", "").replace("hymd::", "")
    return name.split("(")[0][:60]


# ---- launch list ---------------------------------------------------------------------------
rows = list(csv.reader(open(os.path.join(src, "launches.csv"), errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui].strip(), 1e-3)
    seq.append((short(r[ki]), v * scale))
# one step = the launches between two consecutive count_kernel launches (last full step)
marks = [i for i, (k, _) in enumerate(seq) if k.startswith("count_kernel")]
step = seq[marks[-2]:marks[-1]] if len(marks) >= 2 else seq
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for k, v in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
with open(os.path.join(prof, f"{tag}_launches.md"), "w") as fh:
    fh.write(f"# {tag}: ncu launch list of one field-force cycle ({wkey})\n\n"
             "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 "
             "--warmup 3` (cold-cache, serialised: compare SHARES with bench.py's CUDA-event "
             "phases, not absolutes).\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {n} | {v:.1f} | {v / tot:.3f} |\n")
    fh.write(f"| **total** | {len(step)} | {tot:.1f} | 1.000 |\n")

# ---- full capture --------------------------------------------------------------------------
rep = os.path.join(src, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
            ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("smsp__issue_active.avg.pct", "issue active %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex (smem/L1) %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
            ("launch__block_size", "block"),
            ("launch__shared_mem_per_block_dynamic", "dyn smem"),
            ("smsp__inst_executed.sum", "warp insts"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
    idx = [(h.index(m), lbl) for m, lbl in want if m in h]
    kcol = h.index("Kernel Name")

    def to_bytes(val, unit):
        v = float(val.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    traffic = {}
    with open(os.path.join(prof, f"{tag}_kernels.md"), "w") as fh:
        fh.write(f"# {tag}: ncu --set full, hand-written kernels ({wkey})\n\n"
                 "`ncu --set full --clock-control none --import-source on` (one launch each, "
                 "after warm-up).\n\n")
        seen = set()
        for r in rows[2:]:
            name = short(r[kcol])
            if name in seen:
                continue
            seen.add(name)
            fh.write(f"## `{name}`\n\n| metric | value |\n|---|---|\n")
            for i, lbl in idx:
                fh.write(f"| {lbl} | {r[i]} {units[i]} |\n")
            pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
            st = sorted(((float(r[j] or 0), c[len(pre):-len(suf)]) for j, c in enumerate(h)
                         if c.startswith(pre) and c.endswith(suf)), reverse=True)[:6]
            fh.write("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |\n")
            fh.write("\n")
            try:
                rd = to_bytes(r[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")])
                wr = to_bytes(r[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")])
                key = {"paint_kernel<float, 0>": "paint", "paint_kernel<double, 0>": "paint"}.get(name)
                for frag, k in (("paint_kernel", "paint"), ("kspace_force", "kspace"),
                                ("xline_kernel", "kspace"), ("readout_kernel", "readout"),
                                ("readout_gather_kernel", "readout"),
                                ("plane_r2c_kernel", "fft_fwd"), ("plane_c2r_kernel", "fft_inv"),
                                ("plane_r2c_tmem_kernel", "fft_fwd"), ("plane_c2r_tmem_kernel", "fft_inv"),
                                ("count_kernel", "sort_count"), ("scatter_kernel", "sort_scatter"),
                                ("count2_kernel", "sort_count"), ("scatter2_kernel", "sort_scatter")):
                    if frag in name:
                        traffic[k] = int(rd + wr)
            except Exception:
                pass
    tpath = os.path.join(prof, "traffic.json")
    allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
    allt[wkey] = dict(traffic, source=f"profiles/{tag}_kernels.md")
    json.dump(allt, open(tpath, "w"), indent=1)
for f in ("bench.json", "bench_C2.json", "bench_C3.json", "bench_C4_f64.json"):
    p = os.path.join(src, f)
    if os.path.exists(p) and os.path.getsize(p):
        with open(p) as a, open(os.path.join(prof, f"{tag}_{f}"), "w") as b:
            b.write(a.read())
print("wrote profiles for", tag)
