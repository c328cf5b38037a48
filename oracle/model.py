"""ORACLE - TEST INFRASTRUCTURE ONLY. Never imported by the product (exprgrad_b200/).

Restatement of exprgrad's CPU `Model` runtime (exprgrad/model.nim): parameter initialisation
U(initRange) (model.nim:232-251; the Nim RNG is not reproducible, so tests inject params through the
public `params` table), per-call shape inference + zero-filled result tensors (model.nim:275-300,
392-406), `fit` batching over `viewFirst` slices with `epoch += 1` and result re-zeroing
(model.nim:413-454; tensors.nim:290-297), output hand-off (model.nim:370-376).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import numpy as np

from . import cgen
from .ir import Fun, Program, RuntimeError_, to_program
from .passes import compile_program, infer_shapes


class Model:
    def __init__(self, graphs: List[Fun], scalar="float32", seed: int = 0, openmp: bool = True):
        self.source = to_program(list(graphs))
        self.program: Program = self.source
        self.program.scalar_type = scalar
        compile_program(self.program)
        self.dtype = np.float32 if scalar == "float32" else np.float64
        rng = np.random.default_rng(seed)
        self.params: Dict[int, np.ndarray] = {}
        self.caches: Dict[int, np.ndarray] = {}
        for i, td in enumerate(self.program.tensors):
            tid = i + 1
            if td.kind == "Param":
                self.params[tid] = rng.uniform(td.init_range[0], td.init_range[1], td.shape).astype(self.dtype)
            elif td.kind == "Cache":
                self.caches[tid] = np.zeros(td.shape, self.dtype)
        self.epoch = 0
        self.c_source = cgen.emit_program(self.program)
        self.lib = cgen.build(self.c_source, openmp=openmp)
        self.tensors: Dict[int, np.ndarray] = {}
        self.shapes: Dict[int, List[int]] = {}
        self._rng = np.random.default_rng(seed + 1)

    # -- helpers
    def param_ids(self) -> List[int]:
        return list(self.program.params)

    def _alloc(self, target, shapes):
        for tid, shape in shapes.items():
            self.shapes[tid] = list(shape)
            td = self.program.tdef(tid)
            required = tid in target.tensors
            if td.kind == "Param":
                self.tensors[tid] = self.params[tid]
            elif td.kind == "Cache":
                self.tensors[tid] = self.caches[tid]
            elif td.kind == "Random" and required:
                lo, hi = td.random_range
                self.tensors[tid] = self._rng.uniform(lo, hi, shape).astype(self.dtype)
            elif td.kind == "Result" and required:
                self.tensors[tid] = np.zeros(shape, self.dtype)

    def _run(self, target_name):
        target = self.program.targets[target_name]
        n = len(self.program.tensors) + 1
        ptrs = (ctypes.c_void_p * n)()
        shapes = (ctypes.c_long * (n * cgen.MAX_RANK))()
        ranks = (ctypes.c_long * n)()
        lens = (ctypes.c_long * n)()
        for tid, arr in self.tensors.items():
            if arr is None:
                continue
            assert arr.flags["C_CONTIGUOUS"] and arr.dtype == self.dtype
            ptrs[tid] = arr.ctypes.data
        for tid, shape in self.shapes.items():
            ranks[tid] = len(shape)
            lens[tid] = int(np.prod(shape)) if len(shape) else 1
            for d, s in enumerate(shape):
                shapes[tid * cgen.MAX_RANK + d] = s
        fn = getattr(self.lib, cgen.target_symbol(target_name))
        fn.restype = None
        fn(ptrs, shapes, ranks, lens, ctypes.c_long(self.epoch))
        if target.output:
            out = self.tensors[target.output]
            if self.program.tdef(target.output).kind == "Result":
                self.tensors[target.output] = None
            return out
        return None

    # -- public API (model.nim:392-454)
    def call(self, target_name: str, args: Optional[Dict[str, np.ndarray]] = None):
        args = args or {}
        if target_name not in self.program.targets:
            raise RuntimeError_(f"{target_name} is not a target of the model")
        target = self.program.targets[target_name]
        input_shapes = {}
        for name, arr in args.items():
            if name not in self.program.inputs:
                raise RuntimeError_(f"{name} is not an input to the model")
            arr = np.ascontiguousarray(arr, dtype=self.dtype)
            tid = self.program.inputs[name]
            self.tensors[tid] = arr
            input_shapes[tid] = list(arr.shape)
        shapes = infer_shapes(self.program, target_name, input_shapes)
        self._alloc(target, shapes)
        return self._run(target_name)

    def apply(self, target_name, args=None):
        self.call(target_name, args)

    def fit(self, target_name: str, args: Dict[str, np.ndarray], batch_size: int = 32):
        if not args:
            raise RuntimeError_("Model.fit requires at least one input tensor.")
        if target_name not in self.program.targets:
            raise RuntimeError_(f"{target_name} is not a target of the model")
        target = self.program.targets[target_name]
        first = next(iter(args.values()))
        batch_count = first.shape[0] // batch_size
        input_shapes = {}
        for name, arr in args.items():
            if name not in self.program.inputs:
                raise RuntimeError_(f"{name} is not an input to the model")
            input_shapes[self.program.inputs[name]] = [batch_size] + list(arr.shape[1:])
        shapes = infer_shapes(self.program, target_name, input_shapes)
        self._alloc(target, shapes)
        self.epoch += 1
        for b in range(batch_count):
            off = b * batch_size
            for name, arr in args.items():
                self.tensors[self.program.inputs[name]] = np.ascontiguousarray(arr[off:off + batch_size], dtype=self.dtype)
            self._run(target_name)
            for tid in target.tensors:
                if self.program.tdef(tid).kind == "Result":
                    if self.tensors.get(tid) is None:
                        self.tensors[tid] = np.zeros(self.shapes[tid], self.dtype)
                    else:
                        self.tensors[tid].fill(0)


def compile(*graphs, scalar="float32", seed=0, openmp=True) -> Model:
    gs = []
    for g in graphs:
        if isinstance(g, (list, tuple)):
            gs.extend(g)
        else:
            gs.append(g)
    return Model(gs, scalar=scalar, seed=seed, openmp=openmp)
