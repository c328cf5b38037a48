"""Host-side mirror of exprgrad's `Model[T]` interface (exprgrad/model.nim:29-42, 262-273, 392-454) for
CompileGpu targets: compile / call / apply / fit / params / caches / epoch. Every method is a thin
call into libegb200.so (include/egb200.h, group 3); numpy arrays stand in for the reference's host
`Tensor[T]` (exprgrad/tensors.nim:19-25)."""
import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import frontend
from ._ffi import RuntimeError_, ShapeError, ValueError_, check, lib
from .gpu import GpuContext, new_gpu_context

MAX_RANK = 8


class Program:
    """Handle of a parsed program inside the library."""

    def __init__(self, text: str):
        raw = text.encode("utf-8")
        h = ctypes.c_void_p()
        check(lib.egb_program_parse(raw, len(raw), ctypes.byref(h)))
        self.handle = h

    @staticmethod
    def from_graphs(graphs: Sequence[frontend.Fun]) -> "Program":
        source = frontend.to_program(list(graphs))
        return Program(frontend.serialize(source))

    def compile(self):
        check(lib.egb_program_compile(self.handle))
        return self

    def serialize(self) -> str:
        need = ctypes.c_size_t(0)
        check(lib.egb_program_serialize(self.handle, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.egb_program_serialize(self.handle, buf, need.value, ctypes.byref(need)))
        return buf.value.decode("utf-8")

    def describe(self, target: str) -> str:
        need = ctypes.c_size_t(0)
        check(lib.egb_program_describe(self.handle, target.encode(), None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.egb_program_describe(self.handle, target.encode(), buf, need.value, ctypes.byref(need)))
        return buf.value.decode()

    def classify(self, target: str, input_shapes: Dict[str, Sequence[int]]) -> List[str]:
        """Device kernel family per IR kernel of `target` (host only; see egb_program_classify)."""
        names, ranks, dims = _pack_shapes(input_shapes)
        need = ctypes.c_size_t(0)
        check(lib.egb_program_classify(self.handle, target.encode(), len(input_shapes), names, ranks, dims, None, 0,
                                       ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.egb_program_classify(self.handle, target.encode(), len(input_shapes), names, ranks, dims, buf,
                                       need.value, ctypes.byref(need)))
        return [line.split(": ", 1)[1] for line in buf.value.decode().splitlines()]

    def lower_dump(self, target: str, input_shapes: Dict[str, Sequence[int]], strict: bool = True, epoch: int = 0) -> List[dict]:
        """The device program of every IR kernel of `target` (host only; see egb_program_lower_dump)."""
        import json
        names, ranks, dims = _pack_shapes(input_shapes)
        need = ctypes.c_size_t(0)
        check(lib.egb_program_lower_dump(self.handle, target.encode(), len(input_shapes), names, ranks, dims, int(strict), epoch,
                                         None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.egb_program_lower_dump(self.handle, target.encode(), len(input_shapes), names, ranks, dims, int(strict), epoch,
                                         buf, need.value, ctypes.byref(need)))
        return [json.loads(line) for line in buf.value.decode().splitlines()]

    def tensor_count(self) -> int:
        n = ctypes.c_int(0)
        check(lib.egb_program_tensor_count(self.handle, ctypes.byref(n)))
        return n.value

    def tensor_info(self, tensor_id: int):
        kind, rank = ctypes.c_int(0), ctypes.c_int(0)
        dims = (ctypes.c_int64 * MAX_RANK)()
        name = ctypes.create_string_buffer(256)
        check(lib.egb_program_tensor_info(self.handle, tensor_id, ctypes.byref(kind), ctypes.byref(rank), dims, name, 256))
        return {"kind": ["result", "input", "param", "cache", "random"][kind.value],
                "shape": [dims[i] for i in range(rank.value)], "name": name.value.decode()}

    def target_output(self, target: str) -> int:
        t = ctypes.c_int(0)
        check(lib.egb_program_target_output(self.handle, target.encode(), ctypes.byref(t)))
        return t.value

    def infer_shapes(self, target: str, input_shapes: Dict[str, Sequence[int]], tensor_id: int = 0) -> List[int]:
        """inferShapes (exprgrad/passes.nim:1386-1436) - host only."""
        names, ranks, dims = _pack_shapes(input_shapes)
        out_rank = ctypes.c_int(0)
        out_dims = (ctypes.c_int64 * MAX_RANK)()
        check(lib.egb_program_infer_shapes(self.handle, target.encode(), len(input_shapes), names, ranks, dims,
                                           tensor_id, ctypes.byref(out_rank), out_dims))
        if out_rank.value < 0:
            return None
        return [out_dims[i] for i in range(out_rank.value)]

    def free(self):
        if self.handle:
            check(lib.egb_program_free(self.handle))
            self.handle = None


def _pack_shapes(shapes: Dict[str, Sequence[int]]):
    n = len(shapes)
    names = (ctypes.c_char_p * max(n, 1))(*[k.encode() for k in shapes])
    ranks = (ctypes.c_int * max(n, 1))(*[len(s) for s in shapes.values()])
    flat = [int(d) for s in shapes.values() for d in s]
    dims = (ctypes.c_int64 * max(len(flat), 1))(*flat)
    return names, ranks, dims


class _StateTable:
    """`model.params` / `model.caches` (exprgrad/model.nim:37-38): tensor id -> array, backed by HBM."""

    def __init__(self, model: "Model", kind: str):
        self.model, self.kind = model, kind

    def ids(self) -> List[int]:
        p = self.model.program
        return [t for t in range(1, p.tensor_count() + 1) if p.tensor_info(t)["kind"] == self.kind]

    def keys(self):
        return self.ids()

    def __iter__(self):
        return iter(self.ids())

    def __len__(self):
        return len(self.ids())

    def __getitem__(self, tensor_id: int) -> np.ndarray:
        return self.model.read_tensor(tensor_id)

    def __setitem__(self, tensor_id: int, value):
        self.model.write_tensor(tensor_id, value)

    def items(self):
        return [(t, self[t]) for t in self.ids()]


MODEL_MAGIC = b"EGB200-MODEL-1\n"


class Model:
    def __init__(self, graphs: Sequence[frontend.Fun], gpu: Optional[GpuContext] = None, seed: int = 0,
                 strict: bool = False, program: Optional[Program] = None):
        self.ctx = gpu or new_gpu_context()
        self.program = program if program is not None else Program.from_graphs(graphs)
        self.program.compile()
        h = ctypes.c_void_p()
        check(lib.egb_model_create(self.ctx.handle, self.program.handle, seed, ctypes.byref(h)))
        self.handle = h
        self.params = _StateTable(self, "param")
        self.caches = _StateTable(self, "cache")
        if strict:
            self.set_option("strict", 1)

    # -- state
    def set_option(self, key: str, value: int):
        check(lib.egb_model_set_option(self.handle, key.encode(), int(value)))

    @property
    def epoch(self) -> int:
        e = ctypes.c_int64(0)
        check(lib.egb_model_epoch(self.handle, ctypes.byref(e)))
        return e.value

    @epoch.setter
    def epoch(self, value: int):
        self.set_option("epoch", value)

    def tensor_shape(self, tensor_id: int) -> List[int]:
        rank = ctypes.c_int(0)
        dims = (ctypes.c_int64 * MAX_RANK)()
        check(lib.egb_model_tensor_shape(self.handle, tensor_id, ctypes.byref(rank), dims))
        return [dims[i] for i in range(rank.value)]

    def read_tensor(self, tensor_id: int) -> np.ndarray:
        out = np.empty(self.tensor_shape(tensor_id), np.float32)
        check(lib.egb_model_read_tensor(self.handle, tensor_id, out.ctypes.data, out.nbytes))
        return out

    def write_tensor(self, tensor_id: int, value):
        arr = np.ascontiguousarray(value, dtype=np.float32)
        check(lib.egb_model_write_tensor(self.handle, tensor_id, arr.ctypes.data, arr.nbytes))

    def tensor_device_ptr(self, tensor_id: int) -> int:
        p = ctypes.c_void_p()
        check(lib.egb_model_tensor_device_ptr(self.handle, tensor_id, ctypes.byref(p)))
        return p.value or 0

    def plan_count(self) -> int:
        n = ctypes.c_int(0)
        check(lib.egb_model_plan_count(self.handle, ctypes.byref(n)))
        return n.value

    def describe_plan(self) -> str:
        need = ctypes.c_size_t(0)
        check(lib.egb_model_describe_plan(self.handle, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.egb_model_describe_plan(self.handle, buf, need.value, ctypes.byref(need)))
        return buf.value.decode()

    # -- execution (model.nim:392-454)
    def _pack(self, args):
        """args: name -> numpy array (host) or GpuTensor (device-resident)."""
        from .gpu import GpuTensor
        keep, shapes, ptrs, dev = [], {}, [], []
        for name, v in args.items():
            if isinstance(v, GpuTensor):
                shapes[name] = v.shape
                ptrs.append(v.buffer.device_ptr)
                dev.append(1)
            else:
                arr = v if (isinstance(v, np.ndarray) and v.dtype == np.float32 and v.flags["C_CONTIGUOUS"]) \
                    else np.ascontiguousarray(v, dtype=np.float32)
                keep.append(arr)
                shapes[name] = arr.shape
                ptrs.append(arr.ctypes.data)
                dev.append(0)
        names, ranks, dims = _pack_shapes(shapes)
        n = len(args)
        data = (ctypes.c_void_p * max(n, 1))(*ptrs)
        on_dev = (ctypes.c_int * max(n, 1))(*dev)
        return n, names, data, ranks, dims, on_dev, keep

    def call(self, target: str, args: Optional[dict] = None, out: Optional[np.ndarray] = None):
        n, names, data, ranks, dims, on_dev, keep = self._pack(args or {})
        out_rank = ctypes.c_int(0)
        out_dims = (ctypes.c_int64 * MAX_RANK)()
        if out is not None and isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]:
            # Model.call into a caller-provided host tensor: one blocking ABI call (streams H2D / compute /
            # D2H for single-contraction targets); a wrong-sized buffer raises GpuError like cl.nim:134-135
            check(lib.egb_model_call_read(self.handle, target.encode(), n, names, data, ranks, dims, on_dev,
                                          out.ctypes.data, out.nbytes, ctypes.byref(out_rank), out_dims))
            shape = [out_dims[i] for i in range(out_rank.value)]
            if list(out.shape) != shape:
                raise RuntimeError_(f"output buffer has shape {list(out.shape)}, the target produces {shape}")
            return out
        check(lib.egb_model_call(self.handle, target.encode(), n, names, data, ranks, dims, on_dev,
                                 ctypes.byref(out_rank), out_dims))
        if out_rank.value < 0:
            self.ctx.synchronize()
            return None
        shape = [out_dims[i] for i in range(out_rank.value)]
        if out is None:
            out = np.empty(shape, np.float32)
        elif list(out.shape) != shape:
            raise RuntimeError_(f"output buffer has shape {list(out.shape)}, the target produces {shape}")
        check(lib.egb_model_read_output(self.handle, out.ctypes.data, out.nbytes))
        return out

    def apply(self, target: str, args: Optional[dict] = None, sync: bool = True):
        n, names, data, ranks, dims, on_dev, keep = self._pack(args or {})
        check(lib.egb_model_call(self.handle, target.encode(), n, names, data, ranks, dims, on_dev, None, None))
        if sync:
            self.ctx.synchronize()

    def fit(self, target: str, args: dict, batch_size: int = 32, log_status: bool = False) -> int:
        n, names, data, ranks, dims, on_dev, keep = self._pack(args)
        if any(on_dev[i] for i in range(n)):
            raise RuntimeError_("Model.fit takes host tensors")
        done = ctypes.c_int64(0)
        check(lib.egb_model_fit(self.handle, target.encode(), n, names, data, ranks, dims, batch_size, ctypes.byref(done)))
        self.ctx.synchronize()
        if log_status:
            print(f"{done.value}/{done.value}")
        return done.value

    # -- checkpoint (exprgrad/io/serialize.nim:344-379: store(model) = program, params, caches)
    def save(self, path: str):
        """`model.save(path)`: the compiled program plus the parameter and cache tensors, read back from HBM the
        way flushStateTensors(to = CompileCpu) does before the reference serialises (model.nim:326-345). The
        container is this backend's own (the reference's is a Nim binary stream); the epoch is kept as well."""
        import json
        text = self.program.serialize().encode("utf-8")
        tensors, blobs = [], []
        for kind, table in (("param", self.params), ("cache", self.caches)):
            for tid in table.ids():
                arr = np.ascontiguousarray(table[tid], dtype=np.float32)
                tensors.append({"id": tid, "kind": kind, "shape": list(arr.shape)})
                blobs.append(arr.tobytes())
        header = json.dumps({"program_bytes": len(text), "epoch": self.epoch, "tensors": tensors}).encode("utf-8")
        with open(path, "wb") as f:
            f.write(MODEL_MAGIC)
            f.write(len(header).to_bytes(8, "little"))
            f.write(header)
            f.write(text)
            for b in blobs:
                f.write(b)

    def free(self):
        if self.handle:
            check(lib.egb_model_free(self.handle))
            self.handle = None


def load_model(path: str, gpu: Optional[GpuContext] = None, strict: bool = False) -> Model:
    """`loadModel[float32](path)` (exprgrad/io/serialize.nim:376-379) onto the device: parse the stored program
    (already compiled: the passes are skipped), create the model and upload parameters and caches - the
    flushStateTensors(to = CompileGpu) direction of model.nim:337-344."""
    import json
    with open(path, "rb") as f:
        if f.read(len(MODEL_MAGIC)) != MODEL_MAGIC:
            raise ValueError_(f"{path} is not an exprgrad_b200 model file")
        header = json.loads(f.read(int.from_bytes(f.read(8), "little")).decode("utf-8"))
        text = f.read(header["program_bytes"]).decode("utf-8")
        model = Model([], gpu=gpu, strict=strict, program=Program(text))
        for t in header["tensors"]:
            n = int(np.prod(t["shape"])) if t["shape"] else 1
            raw = f.read(4 * n)
            if len(raw) != 4 * n:
                raise ValueError_(f"{path} is truncated (tensor {t['id']})")
            if model.tensor_shape(t["id"]) != list(t["shape"]):
                raise ShapeError(f"tensor {t['id']} of {path} has shape {t['shape']}, the program declares "
                                 f"{model.tensor_shape(t['id'])}")
            model.write_tensor(t["id"], np.frombuffer(raw, np.float32).reshape(t["shape"]))
        model.epoch = header.get("epoch", 0)
    return model


def compile(*graphs, gpu: Optional[GpuContext] = None, seed: int = 0, strict: bool = False) -> Model:
    """compile[float32](graphs, gpu=ctx) (exprgrad/model.nim:270-273)."""
    gs = []
    for g in graphs:
        if isinstance(g, (list, tuple)):
            gs.extend(g)
        else:
            gs.append(g)
    return Model(gs, gpu=gpu, seed=seed, strict=strict)
