"""Multi-GPU (needs >= 2 B200s; skipped otherwise): data-parallel dense train steps through the
library (NCCL all-reduce of the gradient bucket inside the plan's CUDA graph) against the oracle's
global-batch steps (BASELINE config 5 at reduced size and at 1024 rows per GPU)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, sizes, per_gpu, steps, out):
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
        pm = eg.compile(*G.dense_net(F, PL, sizes), gpu=ctx, seed=0)
        x, y, params = G.dense_inputs(per_gpu * world, sizes)
        for tid, v in zip(pm.params.ids(), params):
            pm.params[tid] = v
        D.set_data_parallel(pm, comm)
        lo, hi = D.shard_rows(per_gpu * world, rank, world)
        for _ in range(steps):
            pm.apply("train", {"x": x[lo:hi], "y": y[lo:hi]})
        plan = pm.describe_plan()
        assert "allreduce" in plan and "graph yes" in plan, plan
        got = [pm.params[t] for t in pm.params.ids()]
        if rank == 0:
            out.put(got)
        pm.free(); comm.destroy(); ctx.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes,per_gpu,steps", [((64, 48, 32, 10), 24, 3), ((784, 512, 512, 10), 1024, 2)])
def test_data_parallel_matches_global_batch_oracle(sizes, per_gpu, steps):
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
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, per_gpu, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    x, y, params = G.dense_inputs(per_gpu * world, sizes)
    om = o.compile(*G.dense_net(o, OL, sizes, ct="threads"), seed=0)
    ids = sorted(om.params)
    for tid, v in zip(ids, params):
        om.params[tid][...] = v
    for _ in range(steps):
        om.apply("train", {"x": x, "y": y})
    for g, tid, v in zip(got, ids, params):
        assert_close(g, om.params[tid], what=f"param tensor{tid - 1}")
        assert_close(g - v, om.params[tid] - v, tol=2e-3, what=f"update of tensor{tid - 1}")
