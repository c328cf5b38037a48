"""Data-parallel dense train steps through the library against the oracle's global-batch steps (BASELINE
config 5 at reduced size and at 1024 rows per GPU). The gradient bucket is exchanged by the fused peer-memory
kernel (csrc/exchange.cu: reduce-scatter + all-gather over NVLink + gradientDescent update) inside the plan's
CUDA graph; option dp_peer=0 selects the ncclAllReduce(avg) + separate optimizer kernels it replaced.
Multi-GPU cases need >= 2 B200s (skipped otherwise); the single-rank cases run the same plan layout and the
same kernel on one GPU (EGB_DP_FORCE)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fused_updates(plan):
    """gradientDescent updates that run inside the exchange kernel(s) of a plan (the bucket may travel in two launches)"""
    import re
    return sum(int(m) for m in re.findall(r"\+ (\d+) fused gradientDescent updates", plan))


def _worker(rank, world, port, sizes, per_gpu, steps, out, mode="peer"):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        import exprgrad_b200 as eg
        from exprgrad_b200 import dist as D, frontend as F, layers as PL
        import graphs as G
        ctx = eg.new_gpu_context(eg.GpuDevice(rank))
        comm = D.Comm(ctx, rank, world, dist)
        graphs = G.dense_net(F, PL, sizes) if mode != "adam" else _adam_net(F, PL, sizes)
        pm = eg.compile(*graphs, gpu=ctx, seed=0)
        x, y, params = G.dense_inputs(per_gpu * world, sizes)
        for tid, v in zip(pm.params.ids(), params):
            pm.params[tid] = v
        if mode == "resume":
            # checkpoint round trip BEFORE going data parallel: the reloaded (already compiled) program must still
            # know its parameter gradients, else the replicas would silently train without averaging
            import tempfile
            path = os.path.join(tempfile.mkdtemp(), f"ckpt{rank}.egb")
            pm.save(path)
            pm.free()
            pm = eg.load_model(path, gpu=ctx)
        if mode == "nccl":
            pm.set_option("dp_peer", 0)
        D.set_data_parallel(pm, comm)
        lo, hi = D.shard_rows(per_gpu * world, rank, world)
        for step in range(steps):
            if mode == "adam":
                pm.set_option("epoch", step + 1)
            pm.apply("train", {"x": x[lo:hi], "y": y[lo:hi]})
        plan = pm.describe_plan()
        assert ("allreduce" if mode == "nccl" else " exchange peer exchange") in plan and "graph yes" in plan, plan
        if mode in ("peer", "resume"):
            assert _fused_updates(plan) == 6, plan
        got = [pm.params[t] for t in pm.params.ids()]
        if rank == 0:
            out.put(got)
        pm.free(); comm.destroy(); ctx.destroy()
    finally:
        dist.destroy_process_group()


def _adam_net(d, L, sizes):
    x = d.input("x", [-1, sizes[0]]); y = d.input("y", [-1, sizes[-1]])
    h = L.relu(L.dense(x, sizes[0], sizes[1]))
    h = L.relu(L.dense(h, sizes[1], sizes[2]))
    p = L.softmax(L.dense(h, sizes[2], sizes[3]))
    # eps = 1e-2: with the default 1e-8 every element's step is ~rate * sign(g) and elements whose shard gradients
    # nearly cancel amplify summation-order differences without bound (exact fp64 arithmetic is already 9e-5 away
    # from the oracle's fp32 loop after 3 steps); the default eps is covered on one GPU by test_fashion_net_adam_fit
    return [L.cross_entropy(p, y).backprop(L.adam(0.01, eps=1e-2)).target("train", "gpu")]


@pytest.mark.parametrize("sizes,per_gpu,steps,mode", [((64, 48, 32, 10), 24, 3, "peer"), ((784, 512, 512, 10), 1024, 2, "peer"),
                                                      ((64, 48, 32, 10), 24, 3, "nccl"), ((64, 48, 32, 10), 24, 2, "resume"),
                                                      ((64, 48, 32, 10), 24, 3, "adam")])
def test_data_parallel_matches_global_batch_oracle(sizes, per_gpu, steps, mode):
    import exprgrad_b200 as eg
    world = min(len(eg.list_devices()), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from parity_cases import assert_close
    import oracle as o
    from oracle import layers as OL
    import graphs as G
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, per_gpu, steps, out, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    x, y, params = G.dense_inputs(per_gpu * world, sizes)
    om = o.compile(*(G.dense_net(o, OL, sizes, ct="threads") if mode != "adam" else _adam_net(o, OL, sizes)), seed=0)
    ids = sorted(om.params)
    for tid, v in zip(ids, params):
        om.params[tid][...] = v
    for _ in range(steps):
        om.epoch += 1
        om.apply("train", {"x": x, "y": y})
    for g, tid, v in zip(got, ids, params):
        assert_close(g, om.params[tid], what=f"param tensor{tid - 1}")
        assert_close(g - v, om.params[tid] - v, tol=2e-3, what=f"update of tensor{tid - 1}")


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_single_rank_exchange_plan_matches_plain_plan(opt):
    """The data-parallel plan on ONE rank (EGB_DP_FORCE): gradient bucket + exchange kernel with the fused
    gradientDescent updates (or, for adam, the exchange followed by the optimizer kernels) must reproduce the
    plain single-GPU plan, whose SGD runs in the contraction epilogues."""
    import exprgrad_b200 as eg
    from exprgrad_b200 import dist as D, frontend as F, layers as PL
    from parity_cases import assert_close
    import graphs as G
    sizes = (64, 48, 32, 10)
    ctx = eg.new_gpu_context()
    x, y, params = G.dense_inputs(96, sizes)
    outs = []
    for dp in (0, 1):
        pm = eg.compile(*(G.dense_net(F, PL, sizes) if opt == "sgd" else _adam_net(F, PL, sizes)), gpu=ctx, seed=0)
        for tid, v in zip(pm.params.ids(), params):
            pm.params[tid] = v
        comm = None
        if dp:
            os.environ["EGB_DP_FORCE"] = "1"
            comm = D.Comm(ctx, 0, 1)
            D.set_data_parallel(pm, comm)
        try:
            for step in range(3):
                pm.set_option("epoch", step + 1)
                pm.apply("train", {"x": x, "y": y})
            plan = pm.describe_plan()
        finally:
            os.environ.pop("EGB_DP_FORCE", None)
        if dp:
            assert " exchange peer exchange" in plan, plan
            assert _fused_updates(plan) == (6 if opt == "sgd" else 0), plan
        outs.append([pm.params[t] for t in pm.params.ids()])
        pm.free()
        if comm is not None:
            comm.destroy()
    for a, b, v in zip(outs[0], outs[1], params):
        # (adam: the bias gradients are column sums with a different summation order in the two plans, and adam's
        # normalised step amplifies their last-bit differences)
        assert_close(b, a, tol=1e-6 if opt == "sgd" else 1e-5, what=f"{opt}: exchange plan vs plain plan")
        assert_close(b - v, a - v, tol=1e-3, what=f"{opt}: update, exchange plan vs plain plan")
    ctx.destroy()


def test_bucket_is_found_without_the_gradient_table():
    """A program compiled elsewhere (exprgrad's own passes.nim) carries no gradient table: the planner then takes the
    result tensors of the backward block that the optimizer kernels read. Simulated by stripping the table from
    the serialised program; the data-parallel plan must be the same as with it."""
    import exprgrad_b200 as eg
    from exprgrad_b200 import dist as D, frontend as F, layers as PL
    from exprgrad_b200.model import Model, Program
    import graphs as G
    sizes = (64, 48, 32, 10)
    ctx = eg.new_gpu_context()
    text = Program.from_graphs(G.dense_net(F, PL, sizes)).compile().serialize()
    assert "\ngrads " in text
    stripped = text[:text.index("\ngrads ")] + "\nend\n"
    x, y, params = G.dense_inputs(32, sizes)
    plans = []
    os.environ["EGB_DP_FORCE"] = "1"
    try:
        for t in (text, stripped):
            pm = Model([], gpu=ctx, program=Program(t))
            for tid, v in zip(pm.params.ids(), params):
                pm.params[tid] = v
            comm = D.Comm(ctx, 0, 1)
            D.set_data_parallel(pm, comm)
            pm.apply("train", {"x": x, "y": y})
            plans.append(pm.describe_plan())
            pm.free(); comm.destroy()
    finally:
        os.environ.pop("EGB_DP_FORCE", None)
    assert _fused_updates(plans[0]) == 6
    assert plans[0] == plans[1]
    ctx.destroy()
