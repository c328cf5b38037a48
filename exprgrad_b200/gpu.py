"""Host-side mirror of exprgrad's backend-neutral device interface (exprgrad/runtimes/gpu.nim:25-76,
implemented for OpenCL in exprgrad/runtimes/cl.nim:45-207), on top of the C-ABI in include/egb200.h.

Same names and argument meaning as the reference (snake_case), same error behaviour: every failure
raises GpuError with the library's message; size mismatches on write/read raise like cl.nim:112-113,
134-135, 141-142."""
import ctypes
from typing import List, Sequence

import numpy as np

from . import _ffi
from ._ffi import GpuError, check, lib


class GpuDevice:
    def __init__(self, index: int):
        self.index = index

    def _str(self, fn) -> str:
        buf = ctypes.create_string_buffer(256)
        check(fn(self.index, buf, 256))
        return buf.value.decode()

    @property
    def name(self) -> str:  # cl.nim:74
        return self._str(lib.egb_device_name)

    @property
    def vendor(self) -> str:  # cl.nim:75
        return self._str(lib.egb_device_vendor)

    @property
    def version(self) -> str:  # cl.nim:76
        return self._str(lib.egb_device_version)

    @property
    def is_gpu(self) -> bool:  # cl.nim:78-81
        v = ctypes.c_int(0)
        check(lib.egb_device_is_gpu(self.index, ctypes.byref(v)))
        return bool(v.value)


def list_devices() -> List[GpuDevice]:  # cl.nim:63-65
    n = ctypes.c_int(0)
    check(lib.egb_device_count(ctypes.byref(n)))
    return [GpuDevice(i) for i in range(n.value)]


class GpuBuffer:
    """cl.nim:27-30: value object {ctx, size, mem}; freed explicitly (cl.nim:108-109) or with its context."""

    def __init__(self, ctx: "GpuContext", size: int):
        self.ctx = ctx
        self.size = int(size)
        h = ctypes.c_void_p()
        check(lib.egb_alloc_buffer(ctx.handle, self.size, ctypes.byref(h)))
        self.handle = h

    @property
    def device_ptr(self) -> int:
        return lib.egb_buffer_device_ptr(self.handle) or 0

    def dealloc(self):
        if self.handle:
            check(lib.egb_buffer_free(self.handle))
            self.handle = None

    def write(self, data, size: int = None):  # cl.nim:111-120 (blocking)
        if isinstance(data, np.ndarray):
            arr = np.ascontiguousarray(data)
            ptr, nbytes = arr.ctypes.data, arr.nbytes if size is None else size
        else:
            ptr, nbytes = data, size
        check(lib.egb_buffer_write(self.handle, ptr, nbytes))

    def fill(self, value, dtype=np.float32):  # cl.nim:122-126 (asynchronous)
        v = np.array([value], dtype=dtype)
        check(lib.egb_buffer_fill(self.handle, v.ctypes.data, v.itemsize))

    def read_into(self, out: np.ndarray):  # cl.nim:128-138 (blocking)
        if not out.flags["C_CONTIGUOUS"]:
            raise GpuError("read_into needs a contiguous array")
        check(lib.egb_buffer_read_into(self.handle, out.ctypes.data, out.nbytes))

    def read(self, dtype=np.float32) -> np.ndarray:  # cl.nim:140-147
        item = np.dtype(dtype).itemsize
        if self.size % item != 0:
            raise GpuError("Buffer size is not divisible by item type size")
        out = np.empty(self.size // item, dtype)
        if self.size:
            self.read_into(out)
        return out


class GpuContext:
    """cl.nim:22-25, 83-99: one device, one in-order queue (= one CUDA stream)."""

    def compile(self, name: str, source: str) -> "GpuKernel":  # gpu.nim:46
        return GpuKernel(self, name, source)

    def __init__(self, device: GpuDevice = None):
        h = ctypes.c_void_p()
        check(lib.egb_context_create(-1 if device is None else device.index, ctypes.byref(h)))
        self.handle = h
        self.device = device or GpuDevice(0)

    def alloc_buffer(self, size: int) -> GpuBuffer:  # cl.nim:101-106
        return GpuBuffer(self, size)

    def synchronize(self):
        check(lib.egb_context_synchronize(self.handle))

    @property
    def stream(self) -> int:
        return lib.egb_context_stream(self.handle) or 0

    @property
    def launch_count(self) -> int:
        return int(lib.egb_context_launch_count(self.handle))

    def destroy(self):
        if self.handle:
            check(lib.egb_context_destroy(self.handle))
            self.handle = None


class GpuKernel:
    """cl.nim:37-39, 149-207: compile / arg / run. `source` is the text of a compiled program (see
    egb_compile in include/egb200.h); arguments are the target's tensors in order."""

    def __init__(self, ctx: GpuContext, name: str, source: str):
        h = ctypes.c_void_p()
        check(lib.egb_compile(ctx.handle, name.encode(), source.encode(), ctypes.byref(h)))
        self.handle = h
        self.ctx = ctx

    @property
    def arg_count(self) -> int:
        n = ctypes.c_int(0)
        check(lib.egb_kernel_arg_count(self.handle, ctypes.byref(n)))
        return n.value

    def arg(self, index: int, value, shape=None) -> "GpuKernel":
        if isinstance(value, GpuTensor):
            shape, value = value.shape, value.buffer
        if isinstance(value, GpuBuffer):
            check(lib.egb_kernel_arg_buffer(self.handle, index, value.handle))
            if shape is not None:
                dims = (ctypes.c_int64 * max(len(shape), 1))(*[int(d) for d in shape])
                check(lib.egb_kernel_arg_shape(self.handle, index, len(shape), dims))
        else:
            check(lib.egb_kernel_arg_index(self.handle, index, int(value)))
        return self

    def run(self, group_size: Sequence[int], local_size: Sequence[int]):
        if len(group_size) != len(local_size) and len(group_size) > 0:
            raise GpuError("Dimension of group size must equal dimension of local size")  # cl.nim:193-194
        n = len(group_size)
        g = (ctypes.c_int64 * max(n, 1))(*group_size)
        l = (ctypes.c_int64 * max(n, 1))(*local_size)
        check(lib.egb_kernel_run(self.handle, n, g, l))

    def free(self):
        if self.handle:
            check(lib.egb_kernel_free(self.handle))
            self.handle = None


def compile_kernel(ctx: GpuContext, name: str, source: str) -> GpuKernel:  # gpu.nim:46
    return GpuKernel(ctx, name, source)


def new_gpu_context(device: GpuDevice = None) -> GpuContext:  # cl.nim:83-99
    return GpuContext(device)


class GpuTensor:
    """gpu.nim:54-76: {shape, buffer} with alloc/read/write/fill helpers."""

    def __init__(self, ctx: GpuContext, shape: Sequence[int], dtype=np.float32):
        self.shape = [int(s) for s in shape]
        self.dtype = np.dtype(dtype)
        n = 1
        for s in self.shape:
            n *= s
        self.buffer = ctx.alloc_buffer(n * self.dtype.itemsize)

    def view(self, shape: Sequence[int]) -> "GpuTensor":
        """The same buffer under another shape of equal length (tensors.nim `reshape` semantics; no copy)."""
        v = GpuTensor.__new__(GpuTensor)
        v.shape = [int(s) for s in shape]
        v.dtype = self.dtype
        v.buffer = self.buffer
        n = 1
        for s in v.shape:
            n *= s
        if n * self.dtype.itemsize != self.buffer.size:
            raise GpuError("view: shape does not match the buffer size")
        return v

    def read_into(self, tensor: np.ndarray):
        assert list(tensor.shape) == self.shape
        self.buffer.read_into(tensor)

    def read(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if out.size:
            self.buffer.read_into(out)
        return out

    def write(self, tensor: np.ndarray):
        self.buffer.write(np.ascontiguousarray(tensor, dtype=self.dtype))

    def fill(self, value):
        self.buffer.fill(value, self.dtype)


def alloc_tensor(ctx: GpuContext, shape: Sequence[int], dtype=np.float32) -> GpuTensor:  # gpu.nim:58-62
    return GpuTensor(ctx, shape, dtype)


# ---------------------------------------------------------------- measurement helpers (bench.py)
class GpuEvent:
    """CUDA event on the context's stream."""

    def __init__(self, ctx: GpuContext):
        self.ctx = ctx
        h = ctypes.c_void_p()
        check(lib.egb_event_create(ctx.handle, ctypes.byref(h)))
        self.handle = h

    def record(self):
        check(lib.egb_event_record(self.ctx.handle, self.handle))

    def elapsed_ms(self, stop: "GpuEvent") -> float:
        ms = ctypes.c_double(0)
        check(lib.egb_event_elapsed_ms(self.handle, stop.handle, ctypes.byref(ms)))
        return ms.value


KERNEL_CLASSES = {"gemm": 0, "split": 1, "fill": 2, "interp": 3, "reduce": 4, "eltwise": 5, "conv": 6, "conv_fwd": 6,
                  "other": 7, "conv_dw": 8, "conv_dimg": 9, "exchange": 10, "all": -1}


def set_timing(ctx: GpuContext, enabled: bool):
    check(lib.egb_context_set_timing(ctx.handle, 1 if enabled else 0))


def kernel_time(ctx: GpuContext, kernel_class: str):
    """(total device ms, launches) of one kernel class since set_timing(ctx, True)."""
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    check(lib.egb_context_kernel_time(ctx.handle, KERNEL_CLASSES[kernel_class], ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


_pinned = []


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked host memory (kept alive for the life of the process)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if len(shape) else 1
    p = ctypes.c_void_p()
    check(lib.egb_host_alloc(max(n * dt.itemsize, 1), ctypes.byref(p)))
    buf = (ctypes.c_char * (n * dt.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    _pinned.append((p, buf))
    return arr
