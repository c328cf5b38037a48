#!/usr/bin/env python
"""Summarise any .ncu-rep (ncu --set full) into profiles/<name>.txt: one column per captured launch."""
import csv, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}); one column per launch\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            vals = [r[i].split("(")[0].replace("void ", "").replace("unnamed>::", "")[:26] for r in rows[2:]]
            f.write(f"{w:75s} {units[i]:10s} " + "  ".join(vals) + "\n")
print(open(out).read())
