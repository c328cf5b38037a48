"""Device timeline of the contraction launches inside the dense train step (CUDA-graph replay).
EGB_GEMM_TRACE makes CTA 0 of every contraction record clock stamps at its phase boundaries."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "gemm_trace.txt")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
os.environ["EGB_GEMM_TRACE"] = OUT
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL
import graphs as G
ctx = eg.new_gpu_context()
pm = eg.compile(*G.dense_net(F, PL), gpu=ctx, seed=0)
if os.environ.get("EGB_DP_FORCE"):   # the data-parallel plan layout on one rank (gradient bucket + exchange kernel)
    from exprgrad_b200 import dist as D
    comm = D.Comm(ctx, 0, 1)
    D.set_data_parallel(pm, comm)
for kv in sys.argv[1:]:          # planner options, e.g. concurrent=0 splitk=0
    k, v = kv.split("="); pm.set_option(k, int(v))
x, y, params = G.dense_inputs(1024)
dx, dy = eg.alloc_tensor(ctx, x.shape), eg.alloc_tensor(ctx, y.shape)
dx.write(x); dy.write(y)
STEPS = 6
for _ in range(STEPS):
    pm.apply("train", {"x": dx, "y": dy}, sync=False)
ctx.synchronize()
e0, e1 = eg.gpu.GpuEvent(ctx), eg.gpu.GpuEvent(ctx)
e0.record()
for _ in range(50):
    pm.apply("train", {"x": dx, "y": dy}, sync=False)
e1.record()
print("step us (with tracing):", e0.elapsed_ms(e1) / 50 * 1e3)
ctx.synchronize()
pm.apply("train", {"x": dx, "y": dy}, sync=True)   # one isolated replay: its stamps are the ones reported
print(pm.describe_plan())
pm.free(); ctx.destroy()
rows = [list(map(int, l.split())) for l in open(OUT)]
last = [r for r in rows if r[0]][-9:]   # graph replays overwrite the slots of the captured launches
last.sort(key=lambda r: r[0])
t0 = last[0][0]
print("  start_us   end_us | setup  pdlwait  1st-load  mainloop  epilogue  exit | M N K BN ck grid   (phase times in us at 1.9 GHz clk)")
for r in last:
    clk = lambda a, b: (r[b] - r[a]) / 1965.0 if r[a] and r[b] else float("nan")
    if r[13] == 0:   # head_rows kernel
        print(f"{(r[0]-t0)/1e3:9.2f} {(r[8]-t0)/1e3:9.2f} | head rows={r[9]} classes={r[10]} Kin={r[11]} Nout={r[12]} grid={r[14]}: tables-early {clk(1,2):.2f} pdlwait {clk(2,3):.2f} "
              f"tables+row {clk(3,4):.2f} forward {clk(4,5):.2f} softmax+stores {clk(5,15):.2f} adjoint+stores {clk(15,16):.2f} colsums {clk(16,6):.2f}")
        continue
    print(f"{(r[0]-t0)/1e3:9.2f} {(r[8]-t0)/1e3:9.2f} | {clk(1,2):5.2f} {clk(2,3):7.2f} {clk(3,4):8.2f} {clk(4,5):9.2f} {clk(5,6):9.2f} {clk(6,7):5.2f} | "
          + " ".join(str(v) for v in r[9:15]) + (f" | epi unit0: tmem_ld {clk(5,15):.2f} transpose {clk(15,16):.2f} finish {clk(16,17):.2f}" if r[13] <= 1 else f" | csk: stage+sync {clk(5,15):.2f} dsmem-reduce {clk(15,16):.2f} finish {clk(16,17):.2f} cluster-wait {clk(17,6):.2f}")
          + (f" | finish: inputs {clk(16,18):.2f} math {clk(18,19):.2f} fence {clk(19,20):.2f} tma-store {clk(20,21):.2f} colsum {clk(21,17):.2f}" if len(r) > 21 and r[18] else ""))
