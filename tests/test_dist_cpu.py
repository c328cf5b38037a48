"""CPU (gloo, world_size 2): the host-side data-parallel plumbing and the arithmetic of the scheme.
No GPU: the per-rank compute is done by the oracle; the exchange goes through torch.distributed/gloo
exactly where the product uses one ncclAllReduce(avg) on the gradient bucket."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from exprgrad_b200 import dist as D
        import oracle as o
        from oracle import layers as OL
        import graphs as G
        # 1. unique-id exchange: rank 0 makes it, everybody gets the same 128 bytes
        uid = D.exchange_unique_id(lambda: bytes(range(128)), rank, world, dist)
        assert uid == bytes(range(128))
        # 2. contiguous equal shards
        sizes = (12, 8, 6, 4)
        gb = 16
        lo, hi = D.shard_rows(gb, rank, world)
        assert (hi - lo) * world == gb and lo == rank * (gb // world)
        with pytest.raises(ValueError):
            D.shard_rows(gb + 1, rank, world)
        # 3. DP step == global-batch step: local SGD deltas are (-rate * local mean gradient); averaging
        #    the deltas over equal shards is the all-reduce(avg) of the gradient bucket.
        x, y, params = G.dense_inputs(gb, sizes)
        m = o.compile(*G.dense_net(o, OL, sizes), seed=0)
        ids = sorted(m.params)
        for tid, v in zip(ids, params):
            m.params[tid][...] = v
        m.apply("train", {"x": x[lo:hi], "y": y[lo:hi]})
        new = []
        for tid, v in zip(ids, params):
            delta = torch.from_numpy(m.params[tid] - v)
            dist.all_reduce(delta)
            new.append(v + (delta / world).numpy())
        if rank == 0:
            g = o.compile(*G.dense_net(o, OL, sizes), seed=0)
            for tid, v in zip(ids, params):
                g.params[tid][...] = v
            g.apply("train", {"x": x, "y": y})
            err = max(float(np.abs(a - g.params[t]).max() / max(np.abs(g.params[t]).max(), 1e-30)) for a, t in zip(new, ids))
            upd = max(float(np.abs((a - v) - (g.params[t] - v)).max() / max(np.abs(g.params[t] - v).max(), 1e-30))
                      for a, t, v in zip(new, ids, params))
            out.put((err, upd))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_data_parallel_equivalence():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    err, upd = out.get(timeout=5)
    # the update is recovered as (new - old) in fp32, which costs ~1e-4 of its own magnitude
    assert err < 1e-6 and upd < 2e-3, (err, upd)
