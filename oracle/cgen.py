"""ORACLE - TEST INFRASTRUCTURE ONLY. Never imported by the product (exprgrad_b200/).

C emitter that restates exprgrad's CPU lowering (exprgrad/llvmgen.nim:193-361, 518-563): every kernel
becomes a plain loop nest in the reference's loop order with

  * InstrRead   -> load                      (llvmgen.nim:277-293)
  * InstrWrite  -> load, fadd, store         (llvmgen.nim:294-297)   [accumulate, never fused]
  * Scalar ops  -> separate fmul/fadd/fdiv, `0 - x` negate, ordered compares, select evaluating
                   both arms (llvmgen.nim:219-262), libm expf/logf/sinf/cosf/powf/sqrtf (139-145)
  * Index ops   -> int64 add/sub/mul, truncating sdiv/srem, wrap = ((a % b) + b) % b (223-228)
  * literals    -> f64 rounded to T after f64 constant folding with the identities of
                   propagateConstants (exprgrad/passes.nim:1614-1706)
  * loops       -> `for (i = start; i != stop ...)` entered only if i != stop; emitted as `<`
                   (llvmgen.nim:322-361)
  * threads     -> the first independent loop is hoisted outermost and split into contiguous chunks
                   (passes.nim:1801-1823, model.nim:110-132): `omp parallel for schedule(static)`.

The generated C is compiled with gcc -O3 -march=native -ffp-contract=off -fno-fast-math, i.e. strict
IEEE fp32 like the reference's default JIT (O3, host features, fast-math off: llvmgen.nim:616-647,
wrappers/llvm.nim:486-491). Index expression row-major flattening follows passes.nim:782-843.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from typing import Dict, List

from .ir import GeneratorError, Instr, Kernel, LinearIndex, Program, Target

MAX_RANK = 8
BUILD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")


def _flit(v: float, ctype: str) -> str:
    if v != v:
        return f"(({ctype})NAN)"
    if v in (float("inf"), float("-inf")):
        return f"(({ctype}){'-' if v < 0 else ''}INFINITY)"
    return f"(({ctype}){float(v).hex()})"


class _KernelEmitter:
    def __init__(self, kernel: Kernel, ctype: str, out: List[str]):
        self.k = kernel
        self.ctype = ctype
        self.out = out
        self.types: Dict[int, str] = {}
        self.const: Dict[int, tuple] = {}  # reg -> (kind, value) compile-time constants
        self.alias: Dict[int, int] = {}
        self.arrays: Dict[int, tuple] = {}
        self.extra = 0
        for loop in kernel.loops:
            self.types[loop.iter] = "long"

    # -- registers
    def r(self, reg: int) -> str:
        reg = self._res(reg)
        if reg in self.const:
            kind, v = self.const[reg]
            if kind == "scalar":
                return _flit(v, self.ctype)
            if kind == "index":
                return f"{int(v)}L"
            return "1" if v else "0"
        return f"r{reg}"

    def _res(self, reg):
        while reg in self.alias:
            reg = self.alias[reg]
        return reg

    def cval(self, reg):
        return self.const.get(self._res(reg))

    def typ(self, reg):
        reg = self._res(reg)
        if reg in self.const:
            return {"scalar": self.ctype, "index": "long", "bool": "int"}[self.const[reg][0]]
        return self.types[reg]

    def line(self, depth, s):
        self.out.append("  " * depth + s)

    # -- one instruction
    def instr(self, ins: Instr, depth: int):
        k, a = ins.kind, ins.args
        T = self.ctype
        fsuf = "f" if T == "float" else ""

        def define(typ, expr):
            self.types[ins.res] = typ
            self.line(depth, f"const {typ} r{ins.res} = {expr};")

        def setc(kind, v):
            self.const[ins.res] = (kind, v)

        def is_zero(c):
            return c is not None and ((c[0] == "bool" and c[1] is False) or (c[0] != "bool" and c[1] == 0))

        def is_one(c):
            return c is not None and ((c[0] == "bool" and c[1] is True) or (c[0] != "bool" and c[1] == 1))

        if k == "Scalar":
            return setc("scalar", float(ins.lit or 0.0))
        if k == "Index":
            return setc("index", int(ins.lit or 0))
        if k == "Boolean":
            return setc("bool", bool(ins.lit))
        c = [self.cval(x) for x in a]
        if k in ("Add", "Sub", "Mul", "Div", "IndexDiv", "Mod"):
            # identities + folding (passes.nim:1656-1706); scalars fold in float64
            if k == "Add" and is_zero(c[0]):
                self.alias[ins.res] = a[1]; return
            if k in ("Add", "Sub") and is_zero(c[1]):
                self.alias[ins.res] = a[0]; return
            if k == "Mul":
                if is_zero(c[0]):
                    self.alias[ins.res] = a[0]; return
                if is_zero(c[1]):
                    self.alias[ins.res] = a[1]; return
                if is_one(c[0]):
                    self.alias[ins.res] = a[1]; return
                if is_one(c[1]):
                    self.alias[ins.res] = a[0]; return
            if k in ("Div", "IndexDiv"):
                if is_zero(c[0]) or is_one(c[1]):
                    self.alias[ins.res] = a[0]; return
            if k == "Mod" and is_zero(c[0]):
                self.alias[ins.res] = a[0]; return
            if c[0] is not None and c[1] is not None:
                x, y = c[0][1], c[1][1]
                kind = c[0][0]
                if k == "Add": return setc(kind, x + y)
                if k == "Sub": return setc(kind, x - y)
                if k == "Mul": return setc(kind, x * y)
                if k == "Div":
                    if y == 0:
                        v = float("nan") if x == 0 else (float("inf") if x > 0 else float("-inf"))
                    else:
                        v = x / y
                    return setc("scalar", v)
                if k == "IndexDiv":
                    q = abs(x) // abs(y)
                    return setc("index", q if (x >= 0) == (y >= 0) else -q)
                if k == "Mod":
                    q = abs(x) // abs(y)
                    q = q if (x >= 0) == (y >= 0) else -q
                    return setc("index", x - q * y)
            t = self.typ(a[0])
            op = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "IndexDiv": "/", "Mod": "%"}[k]
            return define(t, f"{self.r(a[0])} {op} {self.r(a[1])}")
        if k in ("Eq", "Lt", "Le"):
            if k == "Eq" and c[0] is None and c[1] is None and self._res(a[0]) == self._res(a[1]):
                return setc("bool", True)
            if c[0] is not None and c[1] is not None:
                x, y = c[0][1], c[1][1]
                return setc("bool", {"Eq": x == y, "Lt": x < y, "Le": x <= y}[k])
            op = {"Eq": "==", "Lt": "<", "Le": "<="}[k]
            return define("int", f"{self.r(a[0])} {op} {self.r(a[1])}")
        if k in ("And", "Or"):
            if c[0] is not None and c[1] is not None:
                return setc("bool", (c[0][1] and c[1][1]) if k == "And" else (c[0][1] or c[1][1]))
            return define("int", f"{self.r(a[0])} {'&' if k == 'And' else '|'} {self.r(a[1])}")
        if k == "Select":
            if c[0] is not None:
                self.alias[ins.res] = a[1] if c[0][1] else a[2]
                return
            return define(self.typ(a[1]), f"{self.r(a[0])} ? {self.r(a[1])} : {self.r(a[2])}")
        if k == "Wrap":
            return define("long", f"(({self.r(a[0])} % {self.r(a[1])}) + {self.r(a[1])}) % {self.r(a[1])}")
        if k == "Negate":
            t = self.typ(a[0])
            zero = "0L" if t == "long" else _flit(0.0, T)
            return define(t, f"{zero} - {self.r(a[0])}")
        if k in ("Sin", "Cos", "Exp", "Sqrt", "Ln"):
            fn = {"Sin": "sin", "Cos": "cos", "Exp": "exp", "Sqrt": "sqrt", "Ln": "log"}[k] + fsuf
            return define(T, f"{fn}({self.r(a[0])})")
        if k == "Pow":
            return define(T, f"pow{fsuf}({self.r(a[0])}, {self.r(a[1])})")
        if k in ("Log", "Log10", "Log2"):
            # exprgrad/llvmgen.nim:501-502 - no CPU lowering exists for these opcodes
            raise GeneratorError(f"Unable to generate LLVM IR for Instr{k}")
        if k == "ToScalar":
            return define(T, f"({T}){self.r(a[0])}")
        if k == "ToIndex":
            return define("long", f"(long){self.r(a[0])}")
        if k == "Shape":
            d = ins.dim
            idx = f"ranks[{ins.tensor}] + ({d})" if d < 0 else str(d)
            return define("long", f"shapes[{ins.tensor} * {MAX_RANK} + {idx}]")
        if k == "Len":
            return define("long", f"lens[{ins.tensor}]")
        if k == "ShapeLen":
            return define("long", f"ranks[{ins.tensor}]")
        if k == "Epoch":
            return define("long", "epoch")
        if k == "Array":
            items = ", ".join(self.r(x) for x in a)
            inner = self.typ(a[0]) if a else T
            self.types[ins.res] = "array"
            if inner == "array":
                # nested arrays: flatten as pointers to rows
                self.arrays[ins.res] = ("nested", list(a))
                return
            self.arrays[ins.res] = ("flat", len(a))
            self.line(depth, f"const {inner} r{ins.res}[{max(len(a), 1)}] = {{{items}}};")
            return
        if k == "ArrayLen":
            return setc("index", self.arrays[self._res(a[0])][1] if self.arrays[self._res(a[0])][0] == "flat"
                        else len(self.arrays[self._res(a[0])][1]))
        if k == "ArrayRead":
            arr = self._res(a[0])
            kind, info = self.arrays[arr]
            if kind == "flat":
                return define(T, f"r{arr}[{self.r(a[1])}]")
            if kind == "ptr":
                return define(T, f"({info})[{self.r(a[1])}]")
            # nested: select the row by index -> emit a switch-free lookup via pointer table
            rows = ", ".join(f"r{self._res(x)}" for x in info)
            self.extra += 1
            self.line(depth, f"const {T}* rows{ins.res}[] = {{{rows}}};")
            self.types[ins.res] = "array"
            self.arrays[ins.res] = ("ptr", f"rows{ins.res}[{self.r(a[1])}]")
            return
        raise GeneratorError(f"Unable to generate code for Instr{k}")

    def index_expr(self, li: LinearIndex, depth: int) -> str:
        for ins in li.setup:
            if ins.res not in self.types and ins.res not in self.const and ins.res not in self.alias:
                self.instr(ins, depth)
        terms = []
        for reg, f in li.factors.items():
            terms.append(self.r(reg) if f == 1 else f"{self.r(reg)} * {f}L")
        if li.constant != 0 or not terms:
            terms.append(f"{li.constant}L")
        return " + ".join(terms)

    def flat_index(self, op, depth: int) -> str:
        if op.is_raw:
            return self.index_expr(op.dims[0], depth)
        expr = ""
        for i, d in enumerate(op.dims):
            e = f"({self.index_expr(d, depth)})"
            expr = e if i == 0 else f"({expr} * shapes[{op.tensor} * {MAX_RANK} + {i}] + {e})"
        return expr if expr else "0L"

    def emit(self, parallel: bool):
        k = self.k
        loops = list(k.loops)
        par_loop = None
        if parallel:
            for i, l in enumerate(loops):
                if l.mode >= 1:
                    par_loop = loops.pop(i)
                    loops.insert(0, par_loop)
                    break
        depth = 1
        self.line(depth, "{")
        depth += 1
        for loop in loops:
            if not loop.has_bounds:
                raise GeneratorError("loop without bounds")
            start = self.index_expr(loop.start, depth)
            stop = self.index_expr(loop.stop, depth)
            if loop is par_loop:
                self.line(depth, f"const long lo{loop.iter} = {start}, hi{loop.iter} = {stop};")
                self.line(depth, "#pragma omp parallel for schedule(static)")
                self.line(depth, f"for (long r{loop.iter} = lo{loop.iter}; r{loop.iter} < hi{loop.iter}; r{loop.iter} += {loop.step}) {{")
            else:
                self.line(depth, f"for (long r{loop.iter} = {start}, e{loop.iter} = {stop}; r{loop.iter} < e{loop.iter}; r{loop.iter} += {loop.step}) {{")
            depth += 1
        T = self.ctype
        for rd in k.reads:
            idx = self.flat_index(rd, depth)
            self.types[rd.data] = T
            self.line(depth, f"const {T} r{rd.data} = t{rd.tensor}[{idx}];")
        for ins in k.instrs:
            self.instr(ins, depth)
        idx = self.flat_index(k.write, depth)
        self.line(depth, f"const long w = {idx};")
        self.line(depth, f"t{k.write.tensor}[w] = t{k.write.tensor}[w] + {self.r(k.write.data)};")
        for _ in loops:
            depth -= 1
            self.line(depth, "}")
        depth -= 1
        self.line(depth, "}")


def emit_program(prog: Program) -> str:
    T = "float" if prog.scalar_type == "float32" else "double"
    out = ["#include <math.h>", "#include <stdint.h>", ""]
    for name, target in prog.targets.items():
        fn = "target_" + hashlib.md5(name.encode()).hexdigest()[:12]
        out.append(f"void {fn}({T}** tensors, const long* shapes, const long* ranks, const long* lens, long epoch) {{")
        for ki, kern in enumerate(target.kernels):
            out.append(f"  // kernel {ki}")
            body: List[str] = []
            # tensors may alias (a kernel may read and write the same tensor) -> plain pointers per kernel
            ts = []
            for _, op in kern.tensor_ops():
                if op.tensor not in ts:
                    ts.append(op.tensor)
            body.append("  {")
            for t in ts:
                body.append(f"    {T}* t{t} = tensors[{t}];")
            em = _KernelEmitter(kern, T, body)
            em.emit(parallel=(target.compile_target == "threads"))
            body.append("  }")
            out.extend(body)
        out.append("}")
        out.append("")
    return "\n".join(out)


def target_symbol(name: str) -> str:
    return "target_" + hashlib.md5(name.encode()).hexdigest()[:12]


_MARCH = None


def native_march() -> str:
    """What `-march=native` resolves to on this host. Built libraries are cached in the source tree and travel with
    it (to the GPU box): the cache key names the instruction set they were compiled for, so a host with another CPU
    rebuilds instead of loading code it cannot execute."""
    global _MARCH
    if _MARCH is None:
        _MARCH = "unknown"
        try:
            out = subprocess.run(["gcc", "-march=native", "-Q", "--help=target"], capture_output=True, text=True).stdout
            for line in out.splitlines():
                parts = line.split()
                if len(parts) == 2 and parts[0] == "-march=":
                    _MARCH = parts[1]
                    break
        except OSError:
            pass
    return _MARCH


def build(source: str, openmp: bool = True):
    os.makedirs(BUILD_DIR, exist_ok=True)
    h = hashlib.sha1((source + str(openmp) + native_march()).encode()).hexdigest()[:16]
    so = os.path.join(BUILD_DIR, f"oracle_{h}.so")
    if not os.path.exists(so):
        c = os.path.join(BUILD_DIR, f"oracle_{h}.c")
        with open(c, "w") as f:
            f.write(source)
        cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w",
               "-o", so + ".tmp", c, "-lm"]
        if openmp:
            cmd.insert(1, "-fopenmp")
        subprocess.run(cmd, check=True)
        os.replace(so + ".tmp", so)
    return ctypes.CDLL(so)
