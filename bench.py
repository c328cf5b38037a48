#!/usr/bin/env python
"""bench.py - headline measurement of the exprgrad hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload auto|matmul|dense|conv2|eltwise]

Workloads (BASELINE.json `configs`), all driven through the public model API
(exprgrad_b200.compile(...).call/apply/fit == exprgrad's compile[T] / Model.call / Model.apply / Model.fit):
  matmul   configs[1]  benchmarks/matmul: c[y,x] ++= a[y,it]*b[it,x], 4096x4096x4096 fp32
                       (metric matmul GFLOP/s = 2MNK / t). A single contraction does not shard: replicas only.
  dense    configs[2]/[4]  784->512->512->10 dense+relu, softmax+crossEntropy, SGD: one full train step
                       (fwd + bwd + update) on a batch of 1024 per GPU, data parallel over N GPUs: the
                       parameter-gradient bucket is exchanged by one fused peer-memory kernel over NVLink
                       (reduce-scatter + all-gather + SGD update), metric train samples/s, weak scaling;
                       plus the strong-scaling point (global batch 8192 fixed, 8192/N per GPU).
  conv2    configs[3]  NHWC 256x224x224x3 images, 64 3x3x3 filters: forward, d_filters and d_images (HBM-bound).
  eltwise  the HBM-streaming elementwise / optimizer kernels the north star names (relu and its adjoint, bias add,
           gradientDescent, adam) on 512 MiB tensors.
A "step" is one pass of the hot path over one batch of synthetic input.

Which workload is the primary JSON line (`--workload auto`): the configuration BASELINE.json's metric is quoted
on when it fits one GPU - the 4096^3 matmul - on a single-GPU box; on a multi-GPU box (more than one GPU visible,
or --gpus N > 1) the workload that actually shards, the data-parallel dense train step at EVERY N including 1, so
that the per-N values of one scaling run are one metric and their ratio is the data-parallel curve. The other
workloads ride along as extra keys of the same line (`dense_train`, `matmul_replicas`, `conv2_fwd_bwd`, `eltwise`).

`value`    device-resident throughput (inputs already in HBM when the timed region starts)
`e2e`      the same through the public API with HOST buffers: H2D of the step's inputs from pinned
           memory and D2H of the step's result inside the timed region
`roofline` the dominant kernel, timed live with CUDA events on the launching stream
`cpu_baseline` / `--impl reference`: the oracle's restatement of the reference's CPU (LLVM-JIT) path on
           this box's host cores (the reference itself needs Nim + LLVM 13, absent from this image).
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import graphs as GR  # noqa: E402  (graph builders shared with the tests; DSL-agnostic)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def visible_gpus():
    """GPUs this process could use (no torch / CUDA initialisation: also called by the CPU arm)."""
    env = os.environ.get("CUDA_VISIBLE_DEVICES")
    if env is not None and env.strip() != "":
        return len([x for x in env.split(",") if x.strip()])
    try:
        r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
        return sum(1 for line in r.stdout.splitlines() if line.startswith("GPU "))
    except (OSError, subprocess.TimeoutExpired):
        return 0


def resolve_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "dense" if (world > 1 or args.gpus > 1 or visible_gpus() > 1) else "matmul"


def oracle_threads(n=None):
    """All host cores for the oracle's OpenMP loop nests, whatever OMP_NUM_THREADS the launcher exported."""
    import ctypes
    n = n or os.cpu_count() or 1
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
    return n


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.15)  # first sample is out before the timed region starts
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0=None, t1=None):
        """Clocks seen between two perf_counter stamps (all samples so far when none fall inside)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        snap = list(self.samples)
        inside = [s for (t, s) in snap if t0 is None or (t0 - 0.02 <= t <= t1 + 0.03)]
        for s in inside or [s for (_, s) in snap]:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(power) if power else None, "samples": len(sm),
                "samples_inside_timed_region": len(inside)}

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()


class Timer:
    """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""

    def __init__(self, ctx, dist, local):
        self.ctx, self.dist, self.local = ctx, dist, local

    def barrier(self):
        self.ctx.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.ctx.synchronize()

    def run(self, fn, steps, warmup):
        from exprgrad_b200 import gpu as G
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = G.GpuEvent(self.ctx), G.GpuEvent(self.ctx)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        ms = e0.elapsed_ms(e1)
        self.barrier()
        t1 = time.perf_counter()
        if self.dist is not None:
            import torch
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{self.local}")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (t0, t1)


# ------------------------------------------------------------------------------ matmul workload
MATMUL_N = 4096
MATMUL_NAME = "benchmarks/matmul c[y,x] ++= a[y,it]*b[it,x] 4096x4096x4096 fp32 (BASELINE configs[1])"


def matmul_config(world):
    return {"workload": MATMUL_NAME, "inputs": "A,B ~ U(0,1) seed 0 (matmul_gpu.nim:69-70)",
            "parallelism": f"replicas only x{world}" if world > 1 else "1 GPU",
            "l2": "fp32 operands + bf16 planes + output = 320 MiB touched per step, larger than the 126 MB L2",
            "numerics": "fp32 in / fp32 out; device: bf16x3 split on tcgen05 (3 MMA passes); bar 1e-4 normalised max error",
            "api": "compile(c.target('c')).apply('c', {a, b}) == exprgrad Model.apply"}


def matmul_inputs(n=MATMUL_N):
    rng = np.random.default_rng(0)  # U(0,1) like benchmarks/matmul/matmul_gpu.nim:69-70
    return rng.uniform(0, 1, (n, n)).astype(np.float32), rng.uniform(0, 1, (n, n)).astype(np.float32)


def cpu_matmul_sample(budget_s=2.5, steps=1, warmup=1, keep_result=False):
    """Oracle (port of the reference's CPU path) on the 4096^3 workload: the whole product when one step fits
    `budget_s`, else the first `rows` rows of A against the full B (the reference parallelises over rows, so
    GFLOP/s on a row block is representative of the whole product)."""
    import oracle as o
    from oracle import layers as OL
    cores = oracle_threads()
    n = MATMUL_N
    a, b = matmul_inputs()
    m = o.compile(*GR.matmul(o, OL, ct="threads"))
    probe = max(cores, 16) * 8
    m.call("c", {"a": a[:probe], "b": b})  # page-in
    t0 = time.perf_counter(); m.call("c", {"a": a[:probe], "b": b}); t1 = time.perf_counter() - t0
    rows = n if t1 / probe * n <= budget_s else int(budget_s / max(t1 / probe, 1e-9))
    rows = max(min(n, rows) - min(n, rows) % cores, cores)
    out = None
    for _ in range(warmup):
        out = m.call("c", {"a": a[:rows], "b": b})
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); out = m.call("c", {"a": a[:rows], "b": b}); times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    cb = {"value": 2.0 * rows * n * n / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
          "sample": f"{'the whole product: all' if rows == n else 'first'} {rows} of {n} rows of A x full B (K=N={n}), "
                    f"oracle C loops y||,it,x (gcc -O3 -march=native -ffp-contract=off), {cores} OpenMP threads, "
                    f"{t:.2f} s per step, {steps} timed steps"}
    # the same loop nest through an LLVM JIT shaped like the reference's own (oracle/jit.py: llvmlite MCJIT,
    # default<O3>, host CPU features, llvmgen.nim:616-647): shows the gcc figure is representative of a JIT path
    try:
        from oracle import jit
        if jit.available():
            jr = min(rows, 1024)
            jit.matmul(a[:cores * 2], b, threads=cores)   # compile + page-in
            t0 = time.perf_counter(); jc = jit.matmul(a[:jr], b, threads=cores); tj = time.perf_counter() - t0
            cb["jit"] = {"value": 2.0 * jr * n * n / tj / 1e9, "unit": "GFLOP/s", "cores": cores,
                         "how": f"llvmlite MCJIT, default<O3>, host cpu + features; first {jr} rows of A x full B, {cores} threads, {tj:.2f} s",
                         "bit_identical_to_port": bool(out is not None and np.array_equal(jc, out[:jr]))}
    except Exception as e:   # the JIT leg is optional evidence, never a reason to lose the baseline
        cb["jit"] = {"unavailable": str(e)[:200]}
    return cb, t, (out if keep_result else None), rows


def run_matmul(args, ctx, timer, rank, world, sampler, cpu_check=True):
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, gpu as G, layers as PL
    peaks = load_peaks()
    n = MATMUL_N
    model = eg.compile(*GR.matmul(F, PL), gpu=ctx)
    a, b = matmul_inputs()
    ha, hb, hc = G.pinned_empty((n, n)), G.pinned_empty((n, n)), G.pinned_empty((n, n))
    ha[...] = a; hb[...] = b
    da, db = eg.alloc_tensor(ctx, (n, n)), eg.alloc_tensor(ctx, (n, n))
    da.write(ha); db.write(hb)
    dev_args, host_args = {"a": da, "b": db}, {"a": ha, "b": hb}

    def step_device():
        model.apply("c", dev_args, sync=False)

    def step_e2e():
        model.call("c", host_args, out=hc)

    step_device(); ctx.synchronize()
    l0 = ctx.launch_count
    ms, span = timer.run(step_device, args.steps, args.warmup)
    launches = (ctx.launch_count - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.window(*span) if sampler else None

    G.set_timing(ctx, True)  # dominant kernel, live, same stream, identical region
    for _ in range(args.steps):
        step_device()
    k_ms, k_n = G.kernel_time(ctx, "gemm")
    all_ms, _ = G.kernel_time(ctx, "all")
    G.set_timing(ctx, False)

    e2e_steps = max(3, min(args.steps, 10))
    if getattr(args, "no_e2e", False):
        step_e2e()   # one call only: the parity check below still needs the result
        e2e_ms = float("nan")
    else:
        e2e_ms, _ = timer.run(step_e2e, e2e_steps, 2)

    flop = 2.0 * n * n * n
    ms_step = ms / args.steps
    passes = 3
    t_kernel = k_ms / max(k_n, 1) * 1e-3
    achieved = passes * flop / t_kernel / 1e12
    # a >500 ms back-to-back region runs under the 1 kW power cap: the sustained cuBLAS figure is the
    # denominator (B200_PROFILING.md); short runs (--steps <= 100) compare against the burst figure
    sustained = ms > 500.0
    peak = peaks["bf16_sustained"] if sustained else peaks["bf16_burst"]
    out = {
        "metric": "matmul_gflops", "value": world * flop / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": matmul_config(world),
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16x3_2cta_kernel (cta_group::2 tcgen05, 256x256 pair tiles)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": 379.0e6, "traffic_source": "profiles/r02z_matmul_2cta_ncu.txt (dram read + write per launch: 323.6 + 55.3 MB; algorithmic 201 MB)",
                     "passes": passes, "algorithmic_tflops": flop / t_kernel / 1e12, "kernel_ms": t_kernel * 1e3,
                     "kernel_share_of_step": k_ms / max(all_ms, 1e-9),
                     "peak_source": peaks["source"] + (", sustained bf16 figure (kernel timed inside a %.0f ms back-to-back region)" % ms
                                                       if sustained else ", burst bf16 figure (short timed region)"),
                     "note": "achieved = passes x 2MNK / kernel time = tensor-pipe rate of the 3-pass fp32 scheme; "
                             "algorithmic_tflops = 2MNK / kernel time; fp32-equivalent ceiling = peak / 3"},
        "e2e": {"value": world * flop / (e2e_ms / e2e_steps * 1e-3) / 1e9, "unit": "GFLOP/s",
                "h2d_bytes_per_step": 2 * n * n * 4, "d2h_bytes_per_step": n * n * 4, "ms_per_step": e2e_ms / e2e_steps,
                "api": "model.call('c', {a, b}, out=c) with pinned host arrays -> egb_model_call_read: H2D of b, then a in "
                       "2 MiB row blocks; split + GEMM per block; D2H of each c block - copies and tensor cores overlap on three streams"},
        "clocks": clocks,
    }
    # parity of what was just timed: the e2e result (host buffer hc) against the oracle's product, element by element
    if rank == 0 and cpu_check and not args.no_cpu:
        cb, _, ref, rows = cpu_matmul_sample(keep_result=True)
        out["cpu_baseline"] = cb
        err = float(np.abs(hc[:rows].astype(np.float64) - ref.astype(np.float64)).max() / np.abs(ref).max())
        out["numerics"] = {"normalised_max_error_vs_oracle": err, "rows_compared": int(rows), "of_rows": n, "bar": 1e-4,
                           "what": "result of the timed e2e call vs the oracle product the cpu_baseline leg computed"}
    else:
        rows64 = a[:8].astype(np.float64) @ b.astype(np.float64).sum(1)
        err = float(np.abs(hc[:8].astype(np.float64).sum(1) - rows64).max() / np.abs(rows64).max())
        out["numerics"] = {"row_sum_error_vs_fp64": err, "bar": 1e-4, "what": "8 row sums (cpu leg skipped)"}
    model.free()
    for t in (da, db):
        t.buffer.dealloc()
    return out


# ------------------------------------------------------------------------------ dense train step
DENSE_SIZES = (784, 512, 512, 10)
DENSE_BATCH = 1024
DENSE_NAME = "synthetic fashion_mnist dense net 784->512->512->10, relu, softmax+crossEntropy, SGD train step (BASELINE configs[2]/[4])"
DENSE_FLOP_PER_SAMPLE = 3286237184 / 1024      # fwd + dW + dX (layer-1 dX is dead)
DENSE_BYTES_PER_STEP = 38.7e6                  # minimal-fusion HBM traffic at batch 1024
DENSE_BUCKET_BYTES = 669706 * 4                # parameter-gradient bucket exchanged per step


def dense_config(world, batch=DENSE_BATCH):
    return {"workload": DENSE_NAME, "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": (f"dp{world}: batch rows sharded, 669706-float gradient bucket averaged across ranks every step"
                            if world > 1 else "1 GPU"),
            "inputs": "x ~ U(0,1) seed 0, one-hot labels seed 1, params U(-0.1,0.1) seed 2, rate 0.01",
            "l2": "working set (~40 MB) is L2-resident by design; every step rewrites all activations and parameters",
            "numerics": "fp32 tensors; device contractions bf16x3 on tcgen05, everything else fp32; bar 1e-4 normalised max error"}


def cpu_dense_sample(steps=5, warmup=1, batch=DENSE_BATCH):
    import oracle as o
    from oracle import layers as OL
    cores = oracle_threads()
    m = o.compile(*GR.dense_net(o, OL, DENSE_SIZES, ct="threads"))
    x, y, params = GR.dense_inputs(batch, DENSE_SIZES)
    for tid, v in zip(sorted(m.params), params):
        m.params[tid][...] = v
    for _ in range(max(warmup, 1)):
        m.apply("train", {"x": x, "y": y})
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); m.apply("train", {"x": x, "y": y}); times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return {"value": batch / t, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} full train steps at batch {batch}, oracle C loop nests (row-split OpenMP, {cores} threads), "
                      f"{t * 1e3:.1f} ms per step"}, t


def run_dense(args, ctx, timer, rank, world, comm, sampler=None, batch=DENSE_BATCH, light=False, exact=False):
    """One data-parallel train step per `step`: every rank holds `batch` rows of the global batch. `exact`: time
    exactly --steps steps after --warmup warm-up steps (the primary line); otherwise at least 200 / 10."""
    import exprgrad_b200 as eg
    from exprgrad_b200 import dist as D, frontend as F, gpu as G, layers as PL
    from exprgrad_b200._ffi import check, lib
    peaks = load_peaks()
    B = batch
    model = eg.compile(*GR.dense_net(F, PL, DENSE_SIZES), gpu=ctx, seed=0)
    x, y, params = GR.dense_inputs(B * world, DENSE_SIZES)
    for tid, v in zip(model.params.ids(), params):
        model.params[tid] = v
    if comm is not None and world > 1:
        if os.environ.get("EGB_DP_NCCL"):
            model.set_option("dp_peer", 0)   # comparison arm: ncclAllReduce(avg) + separate optimizer kernels
        D.set_data_parallel(model, comm)
    lo, hi = D.shard_rows(B * world, rank, world)
    hx, hy = G.pinned_empty((B, DENSE_SIZES[0])), G.pinned_empty((B, DENSE_SIZES[-1]))
    hx[...] = x[lo:hi]; hy[...] = y[lo:hi]
    dx, dy = eg.alloc_tensor(ctx, hx.shape), eg.alloc_tensor(ctx, hy.shape)
    dx.write(hx); dy.write(hy)
    dev_args, host_args = {"x": dx, "y": dy}, {"x": hx, "y": hy}
    last_bias = model.params.ids()[-1]
    hb = G.pinned_empty((DENSE_SIZES[-1],))

    def step_device():
        model.apply("train", dev_args, sync=False)

    def step_e2e():
        model.apply("train", host_args, sync=False)
        check(lib.egb_model_read_tensor(model.handle, last_bias, hb.ctypes.data, hb.nbytes))  # blocking D2H

    steps = args.steps if exact else max(args.steps, 100 if light else 200)
    warm = args.warmup if exact else max(args.warmup, 10)
    step_device(); ctx.synchronize()
    l0 = ctx.launch_count
    ms, span = timer.run(step_device, steps, warm)
    launches = (ctx.launch_count - l0) // (steps + warm)
    clocks = sampler.window(*span) if sampler else None
    plan = model.describe_plan()
    ms_step = ms / steps
    flop = DENSE_FLOP_PER_SAMPLE * B
    out = {"metric": "dense_train_samples_per_s", "value": world * B / (ms_step * 1e-3), "unit": "samples/s",
           "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dense_config(world, B), "gpu_launches": int(launches)}
    if light:
        out["clocks"] = clocks
        model.free()
        return out

    # kernel-class breakdown, eager launches with events (graphs cannot carry the per-launch events)
    reps = 20
    G.set_timing(ctx, True)
    for _ in range(reps):
        step_device()
    classes = {}
    for c in ("gemm", "split", "interp", "eltwise", "reduce", "fill", "exchange", "other"):
        t, k = G.kernel_time(ctx, c)
        if k:
            classes[c] = {"ms_per_step": t / reps, "launches_per_step": k / reps}
    all_ms, _ = G.kernel_time(ctx, "all")
    G.set_timing(ctx, False)

    e2e_steps = 100
    e2e_ms, _ = timer.run(step_e2e, e2e_steps, 5)
    t_tensor = 3 * flop / (peaks["bf16_sustained"] * 1e12)
    t_hbm = DENSE_BYTES_PER_STEP * (B / DENSE_BATCH) / (peaks["hbm_gbs"] * 1e9)
    g = classes.get("gemm", {"ms_per_step": 0.0, "launches_per_step": 0})
    out.update({
        "plan_nodes": plan.count("\n  #"), "cuda_graph": "graph yes" in plan,
        "roofline": {"bound": "tensor", "kernel": "gemm_lat_kernel / gemm_bf16x3_kernel (%d contraction launches per step: cluster split-K, fused epilogues; the two 10-class contractions run inside head_rows_kernel - the rate below divides ALL contraction flops of the step by the contraction launches' time)" % int(g["launches_per_step"]),
                     "achieved": 3 * flop / max(g["ms_per_step"], 1e-9) / 1e9, "peak": peaks["bf16_sustained"],
                     "unit": "TFLOP/s", "frac": 3 * flop / max(g["ms_per_step"], 1e-9) / 1e9 / peaks["bf16_sustained"],
                     "traffic": None, "passes": 3,
                     "step_roofline_us": {"tensor_3pass": t_tensor * 1e6, "hbm_min_fusion": t_hbm * 1e6},
                     "step_frac_of_roofline": max(t_tensor, t_hbm) / (ms_step * 1e-3),
                     "peak_source": peaks["source"] + ", sustained bf16 figure (kernels timed inside a long step)",
                     "kernel_classes": classes, "eager_ms_per_step": all_ms / reps},
        "e2e": {"value": world * B / (e2e_ms / e2e_steps * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int(hx.nbytes + hy.nbytes), "d2h_bytes_per_step": int(hb.nbytes),
                "ms_per_step": e2e_ms / e2e_steps,
                "api": "model.apply('train', {x, y}) with pinned host arrays + read of the updated output bias"},
        "clocks": clocks,
    })
    if world > 1:
        ex = classes.get("exchange")
        # push-only protocol: reduce-scatter and all-gather each store (world-1)/world of the bucket into peer memory,
        # every 16 payload bytes travel as a 32-byte line that carries its own epoch tag (csrc/exchange.cu)
        payload = int(DENSE_BUCKET_BYTES * (world - 1) / world)
        out["exchange"] = {"kernel": "dp_exchange_sgd_kernel x2 (early part of the bucket beside the last contraction, last gradient "
                                     "behind it): reduce-scatter + all-gather by tagged peer-memory stores (NVLink 5 / NVSwitch), "
                                     "fused with the SGD update",
                           "nvlink_bytes_per_rank_per_step": {"read_from_peers": 0, "written_to_peers": 4 * payload,
                                                              "payload_written_to_peers": 2 * payload},
                           "eager_ms_per_step": ex["ms_per_step"] if ex else None,
                           "launches_per_step": ex["launches_per_step"] if ex else None}
    model.free()
    return out


# ------------------------------------------------------------------------------ conv2 forward + backward
CONV_IMG = (256, 224, 224, 3)
CONV_FIL = (64, 3, 3, 3)
CONV_NAME = "benchmarks/conv2: NHWC 256x224x224x3 images, 64 3x3x3 filters, valid, fp32 forward + d_filters + d_images (BASELINE configs[3])"


def cpu_conv2_sample(images=4):
    """Oracle loop nests (threads) for forward, d_filters and d_images on the first `images` images."""
    import oracle as o
    from oracle import layers as OL
    cores = oracle_threads()
    om = o.compile(*GR.conv2_net(o, OL, ct="threads", filters=CONV_FIL), seed=0)
    om.params[sorted(om.params)[0]][...] = np.random.default_rng(1).uniform(-2, 2, CONV_FIL).astype(np.float32)
    img = np.random.default_rng(0).uniform(0, 1, (images,) + CONV_IMG[1:]).astype(np.float32)
    t = {}
    for target in ("conv", "dw", "dimg"):
        om.call(target, {"img": img})
        t0 = time.perf_counter(); om.call(target, {"img": img}); t[target] = time.perf_counter() - t0
    # the adjoint targets recompute the forward pass and the loss adjoint; count each kernel once
    total = t["conv"] + (t["dw"] - t["conv"]) + (t["dimg"] - t["conv"])
    return {"value": images / max(total, 1e-9), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"first {images} of {CONV_IMG[0]} images: oracle forward {t['conv']:.2f} s, dw target {t['dw']:.2f} s, "
                      f"dimg target {t['dimg']:.2f} s ({cores} OpenMP threads; only `image`/`chan` loops of d_images are independent)"}


def run_conv2(args, ctx, timer, rank, world, cpu=True):
    """Times the conv2 targets (forward; loss=sum(out^2) -> d_filters, d_images) and, inside them, the three
    convolution kernels with their own CUDA events (timing classes conv_fwd / conv_dw / conv_dimg)."""
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, gpu as G, layers as PL
    peaks = load_peaks()
    model = eg.compile(*GR.conv2_net(F, PL, filters=CONV_FIL), gpu=ctx, seed=0)
    w = np.random.default_rng(1).uniform(-2, 2, CONV_FIL).astype(np.float32)   # conv2.nim:337-338 ranges
    model.params[model.params.ids()[0]] = w
    n, h, wd, c = CONV_IMG
    f, kh, kw, _ = CONV_FIL
    oh, ow = h - kh + 1, wd - kw + 1
    himg = G.pinned_empty(CONV_IMG)
    himg[...] = np.random.default_rng(0).uniform(0, 1, CONV_IMG).astype(np.float32)
    dimg = eg.alloc_tensor(ctx, CONV_IMG)
    dimg.write(himg)
    out_bytes, in_bytes = n * oh * ow * f * 4, n * h * wd * c * 4
    flop = 2.0 * n * oh * ow * f * kh * kw * c
    steps = max(3, min(args.steps, 10))
    targets, kern = {}, {}
    for target, cls in (("conv", "conv_fwd"), ("dw", "conv_dw"), ("dimg", "conv_dimg")):
        fn = lambda: model.apply(target, {"img": dimg}, sync=False)
        fn(); ctx.synchronize()
        ms, _ = timer.run(fn, steps, 3)
        G.set_timing(ctx, True)
        for _ in range(steps):
            fn()
        k_ms, k_n = G.kernel_time(ctx, cls)
        all_ms, all_n = G.kernel_time(ctx, "all")
        G.set_timing(ctx, False)
        targets[target] = {"target_ms": ms / steps, "launches_per_run": all_n / steps, "all_kernels_ms": all_ms / steps}
        # per run of the target (d_images is two launches: zero fill of the image gradient + the kernel)
        kern[{"conv": "forward", "dw": "d_filters", "dimg": "d_images"}[target]] = k_ms / steps
    total = sum(kern.values())
    hdw, hdi = np.empty(CONV_FIL, np.float32), G.pinned_empty(CONV_IMG)

    def e2e_fn():
        model.call("dw", {"img": himg}, out=hdw)
        model.call("dimg", {"img": himg}, out=hdi)
    e2e_ms, _ = timer.run(e2e_fn, 2, 1)
    alg = in_bytes + out_bytes
    out = {"metric": "conv2_fwd_bwd_images_per_s", "value": world * n / (total * 1e-3), "unit": "images/s",
           "n_gpus": world, "steps": steps, "ms_per_step": total, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
           "data": "synthetic", "config": {"workload": CONV_NAME, "parallelism": f"replicas x{world}" if world > 1 else "1 GPU",
                                           "l2": "3.2 GB output / output gradient per kernel, far larger than the L2"},
           "value_definition": "images / (forward + d_filters + d_images kernel time), each kernel timed by its own CUDA events "
                               "inside its target (timing classes conv_fwd / conv_dw / conv_dimg); whole-target times are in `targets`",
           "kernels_ms": kern,
           "roofline": {"bound": "hbm", "kernel": "conv2_fwd_tc_kernel / conv2_dw_tc_kernel / conv2_dimg_tc_kernel (tcgen05 implicit GEMM)",
                        "achieved": 3 * alg / (total * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": 3 * alg / (total * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
                        "traffic": 3.328e9 + 3.392e9 + 3.544e9,
                        "per_kernel_traffic": {"forward": 3.328e9, "d_filters": 3.392e9, "d_images": 3.544e9},
                        "traffic_source": "profiles/r02d_conv2_fwd_tma_store_ncu.txt (157 MB read + 3.171 GB written), "
                                          "r02z_conv2_dw_ncu.txt (3.387 GB + 5 MB), r03e_conv2_dimg_ncu.txt (3.389 GB + 155 MB)",
                        "algorithmic_bytes_per_kernel": alg,
                        "per_kernel_frac": {k: alg / (v * 1e-3) / 1e9 / peaks["hbm_gbs"] for k, v in kern.items()},
                        "tflops_fp32": {k: flop / (v * 1e-3) / 1e12 for k, v in kern.items()},
                        "peak_source": peaks["source"]},
           "targets": targets,
           "e2e": {"value": world * n / (e2e_ms / 2 * 1e-3), "unit": "images/s", "h2d_bytes_per_step": 2 * in_bytes,
                   "d2h_bytes_per_step": in_bytes + int(hdw.nbytes), "ms_per_step": e2e_ms / 2,
                   "api": "model.call('dw', {img}, out) + model.call('dimg', {img}, out) with a pinned host image batch: "
                          "forward + loss adjoint + d_filters, then forward + loss adjoint + d_images, gradients read back"}}
    if cpu and rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = cpu_conv2_sample()
    model.free()
    dimg.buffer.dealloc()
    return out


# ------------------------------------------------------------------------------ streaming elementwise / optimizer kernels
ELT_N = 1 << 27   # 512 MiB per fp32 tensor
ELT_NAME = ("HBM-streaming elementwise / optimizer kernels on 512 MiB tensors: relu, relu adjoint, bias add, gradientDescent, adam "
            "(layers/dnn.nim:22-27, layers/base.nim:37-53)")


def _elt_graphs(F, PL, rows, cols):
    n = rows * cols
    x = F.input("x", [rows, cols]); g = F.input("g", [rows, cols])
    relu = PL.relu(x)
    radj = F.Fun(); it = F.Iter("it")
    radj.raw[it] += F.select(x.raw[it] >= 0.0, g.raw[it], 0.0)      # derive() of relu (passes.nim:471-476)
    radj.copy_shape(x)
    b = F.param([cols], name="bias")
    biased = F.Fun(); i, j = F.Iter("y"), F.Iter("x")
    biased[i, j] += x[i, j]
    i, j = F.Iter("y"), F.Iter("x")
    biased[i, j] += b[j]
    gflat = F.input("gflat", [n])
    p1 = F.param([n], name="p_sgd"); e1 = F.Fun("Effect", effect=p1); PL.gradient_descent(0.01)(e1, gflat)
    p2 = F.param([n], name="p_adam"); e2 = F.Fun("Effect", effect=p2); PL.adam(0.01)(e2, gflat)
    return [relu.target("relu", "gpu"), radj.target("relu_adjoint", "gpu"), biased.target("bias_add", "gpu"),
            e1.target("sgd", "gpu"), e2.target("adam", "gpu")]


def run_eltwise(args, ctx, timer, rank, world):
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, gpu as G, layers as PL
    peaks = load_peaks()
    rows, cols = 16384, ELT_N // 16384
    n = rows * cols
    model = eg.compile(*_elt_graphs(F, PL, rows, cols), gpu=ctx, seed=0)
    model.set_option("epoch", 1)
    dxt, dgt = eg.alloc_tensor(ctx, (rows, cols)), eg.alloc_tensor(ctx, (rows, cols))
    chunk = np.random.default_rng(0).uniform(-1, 1, (256, cols)).astype(np.float32)
    big = np.tile(chunk, (rows // 256, 1))
    dxt.write(big); dgt.write(big[::-1].copy())
    dgflat = dgt.view((n,))
    # bytes each target's kernels move (4 B x elements read + written)
    cases = [("relu", {"x": dxt}, 8), ("relu_adjoint", {"x": dxt, "g": dgt}, 12), ("bias_add", {"x": dxt}, 16),
             ("sgd", {"gflat": dgflat}, 12), ("adam", {"gflat": dgflat}, 40)]
    steps = max(5, min(args.steps, 20))
    res = {}
    for target, targs, bpe in cases:
        fn = lambda: model.apply(target, targs, sync=False)
        fn(); ctx.synchronize()
        ms, _ = timer.run(fn, steps, 3)
        G.set_timing(ctx, True)
        for _ in range(steps):
            fn()
        k_ms, k_n = G.kernel_time(ctx, "eltwise")
        all_ms, all_n = G.kernel_time(ctx, "all")
        G.set_timing(ctx, False)
        plan = model.describe_plan()
        # adam: the reference's three kernels (first moment, second moment, step) touch 40 B per element; run as one
        # pass (eltwise_stream.cu, adam_fused_kernel) they move 28 - the roofline fraction is taken on what is moved
        moved = 28 if (target == "adam" and "in one pass" in plan) else bpe
        gbs = moved * n / (k_ms / steps * 1e-3) / 1e9
        res[target] = {"target_ms": ms / steps, "kernels_ms": k_ms / steps, "launches": k_n / steps, "other_launches": (all_n - k_n) / steps,
                       "algorithmic_bytes": moved * n, "reference_kernels_bytes": bpe * n, "achieved_gbs": gbs,
                       "frac": gbs / peaks["hbm_gbs"], "specialised": " interp " not in plan}
    # spot check: relu of the resident tensor
    y = model.call("relu", {"x": dxt})
    assert np.array_equal(y[:256], np.maximum(chunk, 0)), "relu result mismatch"
    total_bytes = sum(v["algorithmic_bytes"] for v in res.values())
    total_ms = sum(v["kernels_ms"] for v in res.values())
    dom = res["sgd"]
    out = {"metric": "eltwise_stream_gbs", "value": total_bytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
           "steps": steps, "ms_per_step": total_ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
           "config": {"workload": ELT_NAME, "elements": n, "l2": "every tensor is 512 MiB, four times the L2"},
           "roofline": {"bound": "hbm", "kernel": "elt_stream_kernel<sgd-axpy> (P += (0 - g) * rate, base.nim:37-38)",
                        "achieved": dom["achieved_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dom["frac"],
                        "traffic": 1.565e9, "traffic_source": "profiles/r02c_eltwise_sgd_ncu.txt (dram read 1.074 GB + write 0.491 GB per launch; "
                                                                "the tail of the written lines is still in L2 when the kernel ends)",
                        "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "peak_source": peaks["source"]},
           "per_target": res}
    model.free()
    for t in (dxt, dgt):
        t.buffer.dealloc()
    return out


# ------------------------------------------------------------------------------ entry points
def run_reference(args, rank, world, workload):
    """The reference's CPU path (restated by the oracle) on this box's host cores: W warm-up + K timed steps of
    the same workload and config as our arm. Rank 0 only."""
    if rank != 0:
        return
    if workload == "dense":
        batch = DENSE_BATCH * world
        cb, t = cpu_dense_sample(steps=args.steps, warmup=args.warmup, batch=batch)
        metric, config = "dense_train_samples_per_s", dense_config(world)
    elif workload == "conv2":
        cb = cpu_conv2_sample()
        t = 4 / cb["value"]
        metric, config = "conv2_fwd_bwd_images_per_s", {"workload": CONV_NAME, "parallelism": "1 GPU",
                                                         "l2": "3.2 GB output / output gradient per kernel, far larger than the L2"}
    else:
        # keep the whole --steps K --warmup W run within a few minutes: one step is the whole product when it
        # takes <= 2.5 s on this box, otherwise a row block of it
        cb, t, _, _ = cpu_matmul_sample(budget_s=2.5, steps=args.steps, warmup=args.warmup)
        metric, config = "matmul_gflops", matmul_config(world)
    out = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": cb["unit"], "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "note": "reference CPU path restated by the oracle (Nim + LLVM 13 are not in this image), all host cores; rank 0 only",
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "matmul", "dense", "conv2", "eltwise"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense_train block of the matmul line")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (used for ncu launch lists of the timed region)")
    ap.add_argument("--no-conv", action="store_true", help="skip the conv2_fwd_bwd block")
    ap.add_argument("--no-eltwise", action="store_true", help="skip the eltwise block")
    ap.add_argument("--no-extras", action="store_true", help="primary workload only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = resolve_workload(args, world)
    if args.impl == "reference":
        run_reference(args, rank, world, workload)
        return
    args.warmup = max(args.warmup, 3)
    if args.no_extras:
        args.no_dense = args.no_conv = args.no_eltwise = True
    import exprgrad_b200 as eg
    from exprgrad_b200 import dist as D
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl")
        dist = dist_
    ctx = eg.new_gpu_context(eg.GpuDevice(local))
    timer = Timer(ctx, dist, local)
    comm = D.Comm(ctx, rank, world, dist) if world > 1 else None
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    def extra(fn):
        """a secondary block of a single-GPU run must not take the primary line down with it: report the failure in place
        (multi-rank blocks are not wrapped - a rank that skips a collective would leave the others waiting)"""
        if world > 1:
            return fn()
        try:
            return fn()
        except Exception as e:   # noqa: BLE001
            return {"error": f"{type(e).__name__}: {e}"[:500]}

    if workload == "matmul":
        out = run_matmul(args, ctx, timer, rank, world, sampler)
        if not args.no_dense:
            def dense_block():
                d = run_dense(args, ctx, timer, rank, world, comm, sampler)
                if rank == 0 and not args.no_cpu:
                    d["cpu_baseline"], _ = cpu_dense_sample()
                return d
            out["dense_train"] = extra(dense_block)
        if world == 1 and not args.no_conv:
            out["conv2_fwd_bwd"] = extra(lambda: run_conv2(args, ctx, timer, rank, world))
        if world == 1 and not args.no_eltwise:
            out["eltwise"] = extra(lambda: run_eltwise(args, ctx, timer, rank, world))
    elif workload == "dense":
        out = run_dense(args, ctx, timer, rank, world, comm, sampler, exact=True)
        if rank == 0 and world == 1 and not args.no_cpu:
            out["cpu_baseline"], _ = cpu_dense_sample()
        if not args.no_extras:
            # strong-scaling point of BASELINE configs[4]: global batch 8192 fixed, 8192 / N rows per GPU
            if 8192 % world == 0:
                strong = run_dense(args, ctx, timer, rank, world, comm, sampler, batch=8192 // world, light=True)
                strong["scaling"] = "strong"
                out["dense_strong"] = strong
            mm = run_matmul(args, ctx, timer, rank, world, sampler, cpu_check=False)
            out["matmul_replicas"] = {k: mm[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e", "numerics")}
    elif workload == "conv2":
        out = run_conv2(args, ctx, timer, rank, world)
        out["warmup"] = args.warmup
        out["vs_baseline"] = None
        out["clocks"] = sampler.window() if sampler else None
    else:
        out = run_eltwise(args, ctx, timer, rank, world)
        out["warmup"] = args.warmup
        out["vs_baseline"] = None
        out["clocks"] = sampler.window() if sampler else None
        out["gpu_launches"] = None
    if sampler:
        sampler.stop()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if comm is not None:
        comm.destroy()
    ctx.destroy()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
