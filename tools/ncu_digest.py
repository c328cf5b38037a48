#!/usr/bin/env python
"""Digest of an ncu report for profiles/: per launch the duration, DRAM bytes / throughput, tensor-pipe and issue
utilisation and the top warp-stall reasons (ncu --page raw), i.e. the numbers DESIGN.md and bench.py quote.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep profiles/x.txt ["command that produced it"]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_dim_x", "cluster"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction")]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on   ({rep.split('/')[-1]})\n")
    if cmd:
        f.write(f"# {cmd}\n")
    f.write("# per-launch figures are cold-cache and serialised (ncu flushes caches between replay passes)\n\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("egb::<unnamed>::", "")
        f.write(f"== launch {r[hdr.index('ID')]}: {name[:110]}\n")
        for key, label in want:
            if key in hdr and r[hdr.index(key)] not in ("", "n/a"):
                f.write(f"   {label:38s} {r[hdr.index(key)]:>18s} {units[hdr.index(key)]}\n")
        st = sorted(((float(r[i].replace(",", "") or 0), hdr[i]) for i, h in enumerate(hdr)
                     if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:5]
        f.write("   top stalls (warps per issue):          " + ", ".join(f"{n.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, n in st) + "\n\n")
print(open(out).read()[:3000])
