"""Data-parallel host plumbing: one process per GPU, `torch.distributed` only for
rendezvous (who is rank 0, broadcasting the 128-byte NCCL unique id); the data path - one averaged
all-reduce of the parameter-gradient bucket per train step - runs inside libegb200.so on the model's
own stream / CUDA graph (csrc/dist.cu)."""
import ctypes
import os
from typing import Tuple

from ._ffi import check, lib


def shard_rows(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block of rank `rank` (the `viewFirst` convention, exprgrad/tensors.nim:290-297).
    The global batch must divide evenly: the reference's losses normalise by the local shape[0]
    (exprgrad/layers/base.nim:57-67), so only equal shards make mean-of-means equal the global mean."""
    if total % world != 0:
        raise ValueError(f"global batch {total} is not divisible by {world} ranks")
    per = total // world
    return rank * per, (rank + 1) * per


def exchange_unique_id(make_id, rank: int, world: int, dist=None) -> bytes:
    """Rank 0 calls make_id() and everybody receives the bytes through torch.distributed."""
    if world == 1:
        return make_id()
    if dist is None:
        import torch.distributed as dist
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


class Comm:
    def __init__(self, ctx, rank: int = None, world: int = None, dist=None):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        self.ctx = ctx

        def make_id():
            buf = ctypes.create_string_buffer(128)
            check(lib.egb_comm_unique_id(buf, 128))
            return buf.raw
        uid = exchange_unique_id(make_id, self.rank, self.world, dist) if self.world > 1 else b"\0" * 128
        h = ctypes.c_void_p()
        check(lib.egb_comm_create(ctx.handle, uid, self.rank, self.world, ctypes.byref(h)))
        self.handle = h

    def nccl_version(self) -> int:
        v = ctypes.c_int(0)
        check(lib.egb_comm_info(self.handle, None, None, ctypes.byref(v)))
        return v.value

    def allreduce_avg(self, gpu_tensor):
        n = gpu_tensor.buffer.size // 4
        check(lib.egb_comm_allreduce_avg_f32(self.handle, gpu_tensor.buffer.device_ptr, n))

    def destroy(self):
        if self.handle:
            check(lib.egb_comm_destroy(self.handle))
            self.handle = None


def set_data_parallel(model, comm: Comm):
    check(lib.egb_model_set_data_parallel(model.handle, comm.handle if comm is not None else None))
