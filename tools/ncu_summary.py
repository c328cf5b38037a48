#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_launches.csv and <tag>_prof.ncu-rep into profiles/<tag>_*.txt (tracked)."""
import csv
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

launches = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    order = []
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("egb::<unnamed>::", "")
        if name not in d:
            order.append(name)
        d[name].append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in d.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({tag}); cold-cache, serialised\n")
        f.write(f"# {'kernel':48s} {'launches':>8s} {'avg_us':>10s} {'total_us':>10s} {'share':>7s}\n")
        for k in order:
            v = d[k]
            f.write(f"{k:50s} {len(v):8d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/1e3:10.1f} {100*sum(v)/total:6.1f}%\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.txt")).read())

rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__cycles_active.avg", "sm__inst_executed.sum", "lts__t_bytes.sum",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
    with open(os.path.join(out_dir, f"{tag}_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({tag}); one column per captured launch\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w:70s} {units[i]:12s} " + "  ".join(r[i] for r in rows[2:]) + "\n")
    print(open(os.path.join(out_dir, f"{tag}_ncu_full.txt")).read())
