"""Host-side front-end of exprgrad_b200: builds the `Fun` graph / `++=` kernels the way exprgrad's Nim
macros do and serialises the resulting source `Program` into the text form that libegb200.so parses
(egb_program_parse, csrc/program.cpp). In a real integration this role is played by exprgrad's own
Nim front-end (exprgrad/parser.nim, exprgrad/dsl.nim stay untouched) plus the ~100-line serialiser
shown in INTEGRATION.md; this Python mirror exists so that the library can be driven, tested and
benchmarked without a Nim toolchain. It does no numerical work: passes, shape inference, planning and
execution all happen behind the C ABI.

Mirrors (behaviour, not code):
  * data model           - exprgrad/ir.nim:41-270
  * expression builders  - exprgrad/dsl.nim:21-146 (`>`/`>=` are swapped `<`/`<=`; max/min are selects,
                           dsl.nim:138-142; `or` builds InstrAnd, dsl.nim:50)
  * Fun graph + kernels  - exprgrad/parser.nim:67-97, 137-255, 261-417, 713-817

Python surface:
    c = Fun();  c[y, x] += a[y, it] * b[it, x]          # c[y, x] ++= a[y, it] * b[it, x] | (y, x, it)
    r = Fun();  r.raw[it] += select(0.0 <= x.raw[it], x.raw[it], 0.0)   # r{it} ++= ... | it
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


class CompilerError(Exception): pass
class ParserError(CompilerError): pass
class TypeError_(CompilerError): pass
class GradientError(CompilerError): pass
class GeneratorError(CompilerError): pass
class RuntimeError_(CompilerError): pass
class ShapeError(CompilerError): pass
class ValidationError(CompilerError): pass


# ----------------------------------------------------------------------------- IR (ir.nim:41-270)

# opcode names follow ir.nim:51-76 without the "Instr" prefix
LITERALS = ("Index", "Scalar", "Boolean")


@dataclass
class Instr:
    kind: str
    args: List[int] = field(default_factory=list)
    res: int = 0
    tensor: int = 0
    lit: object = None  # indexLit / scalarLit / booleanLit / array literal items
    dim: int = 0

    def clone(self):
        return Instr(self.kind, list(self.args), self.res, self.tensor, self.lit, self.dim)

    def key(self):
        return (self.kind, tuple(self.args), self.res, self.tensor, self.lit, self.dim)


class LinearIndex:
    """sum(factor * reg) + constant, plus the setup instructions that define non-iterator regs
    (ir.nim:109-112, arithmetic 620-663)."""

    def __init__(self, setup=None, factors=None, constant=0):
        self.setup: List[Instr] = list(setup or [])
        self.factors: Dict[int, int] = dict(factors or {})
        self.constant = constant

    @staticmethod
    def const(c):
        return LinearIndex(constant=c)

    @staticmethod
    def reg(r):
        return LinearIndex(factors={r: 1})

    def clone(self):
        return LinearIndex([i.clone() for i in self.setup], dict(self.factors), self.constant)

    def scale(self, b: int):
        if b == 0:
            return LinearIndex()
        return LinearIndex(self.setup, {r: f * b for r, f in self.factors.items()}, self.constant * b)

    def __add__(self, o):
        if isinstance(o, int):
            o = LinearIndex.const(o)
        res = LinearIndex(self.setup + o.setup, self.factors, self.constant + o.constant)
        for r, f in o.factors.items():
            if r in res.factors:
                res.factors[r] += f
                if res.factors[r] == 0:
                    del res.factors[r]
            else:
                res.factors[r] = f
        return res

    def __neg__(self):
        return self.scale(-1)

    def __sub__(self, o):
        if isinstance(o, int):
            o = LinearIndex.const(o)
        return self + o.scale(-1)

    def mul(self, o):
        if not self.factors:
            return o.scale(self.constant)
        if not o.factors:
            return self.scale(o.constant)
        raise ValueError("non-linear index")

    def eval(self, values: Dict[int, int]) -> int:
        return self.constant + sum(f * values[r] for r, f in self.factors.items())

    def only_register(self) -> int:
        """passes.nim:995-999"""
        if self.constant == 0 and len(self.factors) == 1 and next(iter(self.factors.values())) == 1:
            return next(iter(self.factors.keys()))
        return 0

    def key(self):
        return (tuple(i.key() for i in self.setup), tuple(sorted(self.factors.items())), self.constant)


@dataclass
class Loop:
    iter: int
    has_bounds: bool = False
    start: LinearIndex = field(default_factory=LinearIndex)
    stop: LinearIndex = field(default_factory=LinearIndex)
    step: int = 1
    mode: int = 0  # 0 LoopNone, 1 LoopIndependent, 2 LoopParallel (ir.nim:114)

    def clone(self):
        return Loop(self.iter, self.has_bounds, self.start.clone(), self.stop.clone(), self.step, self.mode)


@dataclass
class TensorOp:
    tensor: int = 0
    is_raw: bool = False
    dims: List[LinearIndex] = field(default_factory=list)
    data: int = 0

    def clone(self):
        return TensorOp(self.tensor, self.is_raw, [d.clone() for d in self.dims], self.data)

    def key_without_data(self):
        return (self.tensor, self.is_raw, tuple(d.key() for d in self.dims))


@dataclass
class Generator:
    kind: str = "None"  # None | Backwards | Gradient | Reshape (ir.nim:187-195)
    tensor: int = 0
    reshape: List[int] = field(default_factory=list)


class Kernel:
    def __init__(self):
        self.generator = Generator()
        self.custom_grad = None  # dict(tensors={tensor: gradId}, kernels=[Kernel], subs={}) (ir.nim:197-203)
        self.nregs = 0
        self.loops: List[Loop] = []
        self.reads: List[TensorOp] = []
        self.instrs: List[Instr] = []  # expr.instrs
        self.res = 0  # expr.res
        self.write = TensorOp()

    def alloc_reg(self) -> int:
        self.nregs += 1
        return self.nregs

    def clone(self):
        k = Kernel()
        k.generator = Generator(self.generator.kind, self.generator.tensor, list(self.generator.reshape))
        if self.custom_grad is not None:
            k.custom_grad = dict(tensors=dict(self.custom_grad["tensors"]),
                                 kernels=[g.clone() for g in self.custom_grad["kernels"]],
                                 subs=dict(self.custom_grad["subs"]))
        k.nregs = self.nregs
        k.loops = [l.clone() for l in self.loops]
        k.reads = [r.clone() for r in self.reads]
        k.instrs = [i.clone() for i in self.instrs]
        k.res = self.res
        k.write = self.write.clone()
        return k

    def tensor_ops(self):
        for r in self.reads:
            yield "read", r
        yield "write", self.write

    def substitute_tensors(self, subs: Dict[int, int]):
        """Kernel.substitute over tensor ids (used for customGrad, passes.nim:624-634)."""
        def sub(t):
            return subs.get(t, t)
        for r in self.reads:
            r.tensor = sub(r.tensor)
        self.write.tensor = sub(self.write.tensor)
        for lst in [self.instrs] + [d.setup for op in list(self.reads) + [self.write] for d in op.dims] + \
                   [l.start.setup for l in self.loops] + [l.stop.setup for l in self.loops]:
            for ins in lst:
                if ins.tensor:
                    ins.tensor = sub(ins.tensor)


@dataclass
class ShapeConstraint:
    kind: str  # Dims | Linear | Copy | Rank (ir.nim:166-185)
    dest: int
    priority: int  # 0 Condition < 1 Inferred < 2 User (ir.nim:178-179)
    rank: int = 0
    dims: List[LinearIndex] = field(default_factory=list)
    reads: Dict[int, List[List[LinearIndex]]] = field(default_factory=dict)
    write: List[LinearIndex] = field(default_factory=list)
    src: int = 0


PRIO_CONDITION, PRIO_INFERRED, PRIO_USER = 0, 1, 2


@dataclass
class TensorDef:
    kind: str  # Result | Input | Param | Cache | Random (ir.nim:222-233)
    shape: List[int] = field(default_factory=list)
    name: str = ""
    init_range: Tuple[float, float] = (-0.1, 0.1)
    random_range: Tuple[float, float] = (0.0, 1.0)
    cache: int = 0


class Target:
    def __init__(self, name, output, compile_target):
        self.name = name
        self.output = output
        self.tensors: List[int] = []  # insertion-ordered set
        self.shapes: List[ShapeConstraint] = []
        self.kernels: List[Kernel] = []
        self.compile_target = compile_target


class Program:
    def __init__(self):
        self.tensors: List[TensorDef] = []  # id = index + 1
        self.inputs: Dict[str, int] = {}
        self.params: List[int] = []
        self.caches: List[int] = []
        self.targets: Dict[str, Target] = {}
        self.scalar_type = "float32"

    def alloc_tensor(self, td: TensorDef) -> int:
        self.tensors.append(td)
        return len(self.tensors)

    def tdef(self, tid: int) -> TensorDef:
        return self.tensors[tid - 1]


# ----------------------------------------------------------------------------- expression DSL

class Expr:
    """ExprBuilder (parser.nim:30-58): kind in {read, iter, instr}; typ in {scalar, index, bool, array}."""

    def __init__(self, kind, typ, children=(), **kw):
        self.kind = kind
        self.typ = typ
        self.children = list(children)
        self.instr = kw.get("instr")
        self.lit = kw.get("lit")
        self.tensor = kw.get("tensor")  # Fun
        self.is_raw = kw.get("is_raw", False)
        self.dim = kw.get("dim", 0)
        self.name = kw.get("name")

    # --- arithmetic (dsl.nim:39-72)
    def _bin(self, other, instr, typ=None, swap=False):
        other = lift(other, self.typ)
        a, b = (other, self) if swap else (self, other)
        return Expr("instr", typ or self.typ, [a, b], instr=instr)

    def __add__(self, o): return self._bin(o, "Add")
    def __radd__(self, o): return self._bin(o, "Add", swap=True)
    def __sub__(self, o): return self._bin(o, "Sub")
    def __rsub__(self, o): return self._bin(o, "Sub", swap=True)
    def __mul__(self, o): return self._bin(o, "Mul")
    def __rmul__(self, o): return self._bin(o, "Mul", swap=True)
    def __truediv__(self, o): return self._bin(o, "Div")
    def __rtruediv__(self, o): return self._bin(o, "Div", swap=True)
    def __floordiv__(self, o): return self._bin(o, "IndexDiv")
    def __mod__(self, o): return self._bin(o, "Mod")
    def __neg__(self): return Expr("instr", self.typ, [self], instr="Negate")
    def __lt__(self, o): return self._bin(o, "Lt", "bool")
    def __le__(self, o): return self._bin(o, "Le", "bool")
    # Nim rewrites a > b to b < a and a >= b to b <= a
    def __gt__(self, o): return self._bin(o, "Lt", "bool", swap=True)
    def __ge__(self, o): return self._bin(o, "Le", "bool", swap=True)
    def eq(self, o): return self._bin(o, "Eq", "bool")
    def and_(self, o): return self._bin(o, "And", "bool")
    def or_(self, o): return self._bin(o, "And", "bool")  # sic: dsl.nim:50 builds InstrAnd for `or`

    def __getitem__(self, idx):  # array read (dsl.nim:84-88)
        return Expr("instr", "scalar", [self, lift(idx, "index")], instr="ArrayRead")


def lift(v, typ="scalar") -> Expr:
    if isinstance(v, Expr):
        return v
    if isinstance(v, bool):
        return Expr("instr", "bool", instr="Boolean", lit=bool(v))
    if isinstance(v, int) and typ == "index":
        return Expr("instr", "index", instr="Index", lit=int(v))
    if isinstance(v, (int, float)):
        return Expr("instr", "scalar", instr="Scalar", lit=float(v))
    if isinstance(v, (list, tuple)):
        return Expr("instr", "array", [lift(x) for x in v], instr="Array")
    raise ParserError(f"cannot lift {v!r}")


def _unop(name):
    def f(x):
        return Expr("instr", "scalar", [lift(x)], instr=name)
    return f


sin, cos, exp, sqrt, ln, log10, log2 = map(_unop, ["Sin", "Cos", "Exp", "Sqrt", "Ln", "Log10", "Log2"])


def pow_(a, b): return Expr("instr", "scalar", [lift(a), lift(b)], instr="Pow")
def log(a, b): return Expr("instr", "scalar", [lift(a), lift(b)], instr="Log")
def to_scalar(i): return Expr("instr", "scalar", [lift(i, "index")], instr="ToScalar")
def to_index(s): return Expr("instr", "index", [lift(s)], instr="ToIndex")
def epoch(): return Expr("instr", "index", instr="Epoch")
def sq(x): return x * x  # dsl.nim:135-136 (same node twice -> one register used twice)
def array_len(a): return Expr("instr", "index", [a], instr="ArrayLen")
def wrap(a, b): return Expr("instr", "index", [lift(a, "index"), lift(b, "index")], instr="Wrap")


def select(cond, a, b):
    a = lift(a)
    b = lift(b, a.typ)
    return Expr("instr", a.typ, [cond, a, b], instr="Select")


def max_(x, y):  # dsl.nim:138-139
    x, y = lift(x), lift(y)
    return select(x > y, x, y)


def min_(x, y):  # dsl.nim:141-142
    x, y = lift(x), lift(y)
    return select(x < y, x, y)


def Iter(name, start=None, stop=None) -> Expr:
    ch = []
    if start is not None or stop is not None:
        ch = [lift(start, "index"), lift(stop, "index")]
    return Expr("iter", "index", ch, name=name)


class _Shape:
    def __init__(self, fun): self.fun = fun
    def __getitem__(self, dim):  # negative dims count from the end (dsl.nim:118-123)
        return Expr("instr", "index", instr="Shape", tensor=self.fun, dim=dim)
    def len(self):
        return Expr("instr", "index", instr="ShapeLen", tensor=self.fun)


class _Accum:
    def __init__(self, value): self.value = value


class _ReadExpr(Expr):
    """A tensor read that also supports `+=` so that `t[y, x] += v` registers a kernel."""
    def __iadd__(self, value):
        return _Accum(lift(value))


class _RawAccessor:
    def __init__(self, fun): self.fun = fun
    def __getitem__(self, idx):
        return _ReadExpr("read", "scalar", [lift(idx, "index")], tensor=self.fun, is_raw=True)
    def __setitem__(self, idx, acc):
        if not isinstance(acc, _Accum):
            raise ParserError("use `t.raw[it] += value`")
        self.fun.add_kernel([lift(idx, "index")], acc.value, is_raw=True)


class KernelBuilder:
    def __init__(self, target, dims, value, is_raw):
        self.target, self.dims, self.value, self.is_raw = target, dims, value, is_raw
        self.custom_grads: List[KernelBuilder] = []
        self.has_custom_grad = False


class Fun:
    """Graph node (parser.nim:67-97)."""
    _ids = itertools.count()

    def __init__(self, kind="Result", **kw):
        self.kind = kind
        self.uid = next(Fun._ids)
        self.targets = set()
        self.tensor = 0
        self.children: List[Fun] = list(kw.get("children", []))
        self.name = kw.get("name", "")
        self.locked = False
        self.kernels: List[KernelBuilder] = []
        self.shape_constr = None  # ("copy", Fun) | ("dims", [Expr])
        self.effect: Optional[Fun] = kw.get("effect")
        self.input_shape = list(kw.get("shape", []))
        self.param_shape = list(kw.get("shape", []))
        self.init_range = kw.get("init_range", (-0.1, 0.1))
        self.random_range = kw.get("random_range", (0.0, 1.0))
        self.cache: Optional[Fun] = kw.get("cache")
        self.reshape_dims = list(kw.get("reshape", []))
        self.cond: Dict[str, Fun] = kw.get("cond", {})
        self.cond_else: Optional[Fun] = kw.get("cond_else")
        self.compile_target = kw.get("compile_target", "cpu")

    # --- reads
    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return _ReadExpr("read", "scalar", [lift(i, "index") for i in idx], tensor=self, is_raw=False)

    def __setitem__(self, idx, acc):
        if not isinstance(acc, _Accum):
            raise ParserError("use `t[...] += value`")
        if not isinstance(idx, tuple):
            idx = (idx,)
        self.add_kernel([lift(i, "index") for i in idx], acc.value, is_raw=False)

    @property
    def raw(self):
        return _RawAccessor(self)

    @property
    def shape(self):
        return _Shape(self)

    def len(self):
        return Expr("instr", "index", instr="Len", tensor=self)

    # --- kernels (parser.nim: ensureInit / addKernel)
    def add_kernel(self, dims, value, is_raw, custom_grad=None):
        if self.kind not in ("Result", "Effect"):
            raise ParserError(f"Cannot add a kernel to {self.kind}")
        if self.locked:
            raise ParserError("Unable to add kernel to locked function")
        kb = KernelBuilder(self, dims, value, is_raw)
        if custom_grad:
            kb.has_custom_grad = True
            kb.custom_grads = custom_grad
        self.kernels.append(kb)
        for e in [value] + list(dims):
            _collect_children(e, self)
        return kb

    def lock(self):
        self.locked = True

    def copy_shape(self, src):  # parser.nim:683-688
        if self.kind != "Result":
            raise ParserError("Cannot set shape of " + self.kind)
        self.shape_constr = ("copy", src)
        if src not in self.children:
            self.children.append(src)

    def with_shape(self, *dims):  # parser.nim:690-695
        if self.kind != "Result":
            raise ParserError("Cannot set shape of " + self.kind)
        if len(dims) == 1 and isinstance(dims[0], (list, tuple)):
            dims = tuple(dims[0])
        dims = [lift(d, "index") for d in dims]
        self.shape_constr = ("dims", dims)
        for d in dims:
            _collect_children(d, self)

    # --- graph ops (parser.nim:713-817)
    def target(self, name, compile_target="threads"):
        return Fun("Target", name=name, children=[self], compile_target=compile_target)

    def backwards(self):
        return Fun("Backwards", children=[self])

    def grad(self, fun):
        return Fun("Gradient", children=[self, fun])

    def reshape(self, shape):
        return Fun("Reshape", name="reshape", children=[self], reshape=list(shape))

    def params(self, stop=frozenset()):
        out: List[Fun] = []
        _params(self, set(stop), out, set())
        return out

    def optimize(self, optim, params=None):
        res = Fun("Multiple")
        for p in (self.params() if params is None else params):
            effect = Fun("Effect", effect=p)
            g = Fun("Gradient", children=[self, p])
            optim(effect, g)
            res.children.append(effect)
        return res

    def backprop(self, optim):
        return self.backwards().optimize(optim)


def _params(fun, stop, out, seen):
    if fun.kind == "Target" and fun.name in stop:
        return
    for c in fun.children:
        _params(c, stop, out, seen)
    if fun.kind == "Param" and fun.uid not in seen:
        seen.add(fun.uid)
        out.append(fun)
    if fun.kind == "Cond":
        for c in list(fun.cond.values()) + ([fun.cond_else] if fun.cond_else else []):
            _params(c, stop, out, seen)


def _collect_children(e: Expr, fun: Fun):
    if e.tensor is not None and e.tensor is not fun:
        t = e.tensor
        if t.kind == "GradientArg":
            pass
        elif t not in fun.children:
            fun.children.append(t)
    for c in e.children:
        _collect_children(c, fun)


def input(name, shape=()): return Fun("Input", name=name, shape=shape)
def param(shape, init_range=(-0.1, 0.1), name=""): return Fun("Param", shape=shape, init_range=init_range, name=name)
def rand(fun, rng): return Fun("Random", children=[fun], random_range=rng)
def cache(fun, name=""): return Fun("Effect", effect=Fun("Cache", cache=fun, name=name))
def grad_arg(fun): return Fun("GradientArg", children=[fun])  # `grad(x)` inside customGrad
def cond(branches: Dict[str, Fun], otherwise=None): return Fun("Cond", cond=dict(branches), cond_else=otherwise)


# ----------------------------------------------------------------------------- Fun graph -> Program

class _BuildCtx:
    def __init__(self, compile_target):
        self.kernel: Kernel = None
        self.iters: Dict[str, int] = {}
        self.grads: Dict[int, int] = {}
        self.blocks = 0
        self.memo: Dict[Tuple[int, int], int] = {}
        self.compile_target = compile_target

    def alloc_block(self):
        self.blocks += 1
        return self.blocks - 1

    def lookup_tensor(self, fun: Fun) -> int:  # parser.nim:140-147
        if fun.kind == "GradientArg":
            tid = self.lookup_tensor(fun.children[0])
            if tid not in self.grads:
                self.grads[tid] = -len(self.grads) - 1
            return self.grads[tid]
        return fun.tensor


def _build_linear(e: Expr, ctx: _BuildCtx) -> LinearIndex:
    li = LinearIndex()
    reg = _build(e, li.setup, ctx.alloc_block(), ctx)
    li.factors = {reg: 1}
    return li


def _build(e: Expr, instrs: List[Instr], block: int, ctx: _BuildCtx) -> int:
    key = (id(e), block)
    if key in ctx.memo:
        return ctx.memo[key]
    k = ctx.kernel
    if e.kind == "read":
        dims = [_build_linear(d, ctx) for d in e.children]
        res = k.alloc_reg()
        k.reads.append(TensorOp(ctx.lookup_tensor(e.tensor), e.is_raw, dims, res))
    elif e.kind == "iter":
        if e.name not in ctx.iters:
            reg = k.alloc_reg()
            ctx.iters[e.name] = reg
            loop = Loop(reg)
            if e.children:
                loop.has_bounds = True
                loop.start = _build_linear(e.children[0], ctx)
                loop.stop = _build_linear(e.children[1], ctx)
                loop.step = 1
            k.loops.append(loop)
        res = ctx.iters[e.name]
    else:
        ins = Instr(e.instr)
        for c in e.children:
            ins.args.append(_build(c, instrs, block, ctx))
        if e.tensor is not None:
            ins.tensor = ctx.lookup_tensor(e.tensor)
        if e.instr in LITERALS:
            ins.lit = e.lit
        if e.instr == "Shape":
            ins.dim = e.dim
        ins.res = k.alloc_reg()
        res = ins.res
        instrs.append(ins)
    ctx.memo[key] = res
    return res


def build_kernel(kb: KernelBuilder, ctx: _BuildCtx) -> Kernel:
    k = Kernel()
    ctx.kernel = k
    block = ctx.alloc_block()
    k.res = _build(kb.value, k.instrs, block, ctx)
    k.write = TensorOp(ctx.lookup_tensor(kb.target), kb.is_raw, [], k.res)
    for d in kb.dims:
        k.write.dims.append(_build_linear(d, ctx))
    if kb.has_custom_grad:
        grads: Dict[int, int] = {}
        gk = []
        for g in kb.custom_grads:
            gctx = _BuildCtx(ctx.compile_target)
            gctx.grads = grads
            gk.append(build_kernel(g, gctx))
            grads = gctx.grads
        k.custom_grad = dict(tensors=dict(grads), kernels=gk, subs={})
    return k


def _alloc_tensors(fun: Fun, prog: Program):
    if fun.tensor != 0:
        return
    k = fun.kind
    if k == "Input":
        if fun.name not in prog.inputs:
            prog.inputs[fun.name] = prog.alloc_tensor(TensorDef("Input", list(fun.input_shape), fun.name))
        fun.tensor = prog.inputs[fun.name]
        if prog.tdef(fun.tensor).shape != list(fun.input_shape):
            raise ParserError(f'Expected shapes for input "{fun.name}" do not match.')
    elif k == "Param":
        fun.tensor = prog.alloc_tensor(TensorDef("Param", list(fun.param_shape), fun.name, init_range=fun.init_range))
    elif k == "Random":
        fun.tensor = prog.alloc_tensor(TensorDef("Random", [], fun.name, random_range=fun.random_range))
    elif k in ("Result", "Gradient", "Reshape"):
        fun.tensor = prog.alloc_tensor(TensorDef("Result", [], fun.name))
    elif k == "Effect":
        _alloc_tensors(fun.effect, prog)
        fun.tensor = fun.effect.tensor
    elif k == "Cache":
        _alloc_tensors(fun.cache, prog)
        fun.tensor = prog.alloc_tensor(TensorDef("Cache", [], fun.name, cache=fun.cache.tensor))
    elif k == "Cond":
        for c in fun.cond.values():
            _alloc_tensors(c, prog)
        if fun.cond_else is not None:
            _alloc_tensors(fun.cond_else, prog)
    for c in fun.children:
        _alloc_tensors(c, prog)
    if k == "Target":
        fun.tensor = fun.children[0].tensor


def _flatten(fun: Fun, target: Target):
    if target.name in fun.targets:
        return
    for c in fun.children:
        _flatten(c, target)
    if fun.kind == "Effect":
        _flatten(fun.effect, target)
    fun.targets.add(target.name)
    k = fun.kind
    if k in ("Result", "Effect"):
        for kb in fun.kernels:
            target.kernels.append(build_kernel(kb, _BuildCtx(target.compile_target)))
        if fun.shape_constr is not None:
            kind, val = fun.shape_constr
            if kind == "copy":
                target.shapes.append(ShapeConstraint("Copy", fun.tensor, PRIO_USER, src=val.tensor))
            else:
                sc = ShapeConstraint("Dims", fun.tensor, PRIO_USER)
                for d in val:
                    ctx = _BuildCtx(target.compile_target)
                    ctx.kernel = Kernel()
                    sc.dims.append(_build_linear(d, ctx))
                target.shapes.append(sc)
    elif k == "Backwards":
        kern = Kernel()
        kern.generator = Generator("Backwards", fun.children[0].tensor)
        target.kernels.append(kern)
    elif k == "Gradient":
        kern = Kernel()
        kern.generator = Generator("Gradient", fun.children[1].tensor)
        kern.write = TensorOp(fun.tensor)
        target.kernels.append(kern)
    elif k == "Reshape":
        kern = Kernel()
        kern.generator = Generator("Reshape", fun.children[0].tensor, list(fun.reshape_dims))
        kern.write = TensorOp(fun.tensor)
        target.kernels.append(kern)
    elif k == "Cond":
        child = fun.cond.get(target.name, fun.cond_else)
        if child is None:
            raise ParserError(f'Conditional node does not have a branch for the target "{target.name}"')
        _flatten(child, target)
        fun.tensor = child.tensor
    elif k == "Random":
        target.shapes.append(ShapeConstraint("Copy", fun.tensor, PRIO_USER, src=fun.children[0].tensor))


def _collect_targets(fun: Fun, targets: Dict[str, Fun], seen):
    if fun.uid in seen:
        return
    seen.add(fun.uid)
    if fun.kind == "Target":
        if fun.name in targets:
            if targets[fun.name] is not fun:
                raise ParserError(f'There are multiple targets named "{fun.name}".')
            return
        targets[fun.name] = fun
    elif fun.kind == "Cond":
        for c in list(fun.cond.values()) + ([fun.cond_else] if fun.cond_else else []):
            _collect_targets(c, targets, seen)
    for c in fun.children:
        _collect_targets(c, targets, seen)
    if fun.kind == "Effect" and fun.effect is not None:
        pass


def to_program(graphs: List[Fun]) -> Program:
    """parser.nim:404-417"""
    prog = Program()
    targets: Dict[str, Fun] = {}
    for g in graphs:
        _alloc_tensors(g, prog)
        _collect_targets(g, targets, set())
    for name, fun in targets.items():
        t = Target(name, fun.tensor, fun.compile_target)
        _flatten(fun, t)
        prog.targets[name] = t
    return prog


# ----------------------------------------------------------------------------- serialisation

_TENSOR_KINDS = {"Result": 0, "Input": 1, "Param": 2, "Cache": 3, "Random": 4}
_COMPILE_TARGETS = {"cpu": 0, "threads": 1, "gpu": 2}
_GEN_KINDS = {"None": 0, "Backwards": 1, "Gradient": 2, "Reshape": 3}


def _s_str(s: str) -> str:
    if not s:
        return "-"
    out = []
    for ch in s.encode("utf-8"):
        if ch <= 32 or ch in (37, 45) or ch >= 127:
            out.append("%%%02x" % ch)
        else:
            out.append(chr(ch))
    return "".join(out)


def _s_instr(i, out):
    scalar, index = 0.0, 0
    if i.kind == "Scalar":
        scalar = float(i.lit or 0.0)
    elif i.kind == "Index":
        index = int(i.lit or 0)
    elif i.kind == "Boolean":
        index = 1 if i.lit else 0
    out += ["I", i.kind, i.res, i.tensor, i.dim, len(i.args)] + list(i.args) + [float(scalar).hex(), index]


def _s_li(li, out):
    out += ["LI", len(li.setup), len(li.factors), li.constant]
    for s in li.setup:
        _s_instr(s, out)
    for reg in sorted(li.factors):
        out += [reg, li.factors[reg]]


def _s_op(op, tag, out):
    out += [tag, op.tensor, 1 if op.is_raw else 0, op.data, len(op.dims)]
    for d in op.dims:
        _s_li(d, out)


def _s_kernel(k, out):
    cg = k.custom_grad
    out += ["K", _GEN_KINDS[k.generator.kind], k.generator.tensor, len(k.generator.reshape)] + list(k.generator.reshape)
    out += [k.nregs, len(k.loops), len(k.reads), len(k.instrs), k.res, 1 if cg is not None else 0, "\n"]
    for l in k.loops:
        out += ["L", l.iter, 1 if l.has_bounds else 0, l.step, l.mode]
        _s_li(l.start, out)
        _s_li(l.stop, out)
        out.append("\n")
    for r in k.reads:
        _s_op(r, "R", out)
        out.append("\n")
    for i in k.instrs:
        _s_instr(i, out)
    out.append("\n")
    _s_op(k.write, "W", out)
    out.append("\n")
    if cg is not None:
        out += ["C", len(cg["tensors"])]
        for t in sorted(cg["tensors"]):
            out += [t, cg["tensors"][t]]
        out.append(len(cg["subs"]))
        for t in sorted(cg["subs"]):
            out += [t, cg["subs"][t]]
        out += [len(cg["kernels"]), "\n"]
        for g in cg["kernels"]:
            _s_kernel(g, out)


def serialize(prog, compiled: bool = False) -> str:
    """Program -> text (format documented in csrc/program.cpp `parse_program`)."""
    out = ["egbprog", 1, "f32" if prog.scalar_type == "float32" else "f64", 1 if compiled else 0, "\n"]
    out += ["tensors", len(prog.tensors), "\n"]
    for t in prog.tensors:
        lo, hi = t.random_range if t.kind == "Random" else t.init_range
        out += ["T", _TENSOR_KINDS[t.kind], len(t.shape)] + [int(d) for d in t.shape]
        out += [float(lo).hex(), float(hi).hex(), t.cache, _s_str(t.name), "\n"]
    out += ["targets", len(prog.targets), "\n"]
    for name, t in prog.targets.items():
        out += ["target", _s_str(name), t.output, _COMPILE_TARGETS.get(t.compile_target, t.compile_target),
                len(t.shapes), len(t.kernels), len(t.tensors)] + list(t.tensors) + ["\n"]
        for sc in t.shapes:
            if sc.kind == "Copy":
                out += ["S", "copy", sc.dest, sc.priority, sc.src]
            elif sc.kind == "Dims":
                out += ["S", "dims", sc.dest, sc.priority, len(sc.dims)]
                for d in sc.dims:
                    _s_li(d, out)
            elif sc.kind == "Rank":
                out += ["S", "rank", sc.dest, sc.priority, sc.rank]
            else:
                out += ["S", "linear", sc.dest, sc.priority, len(sc.reads)]
                for tensor, dims in sc.reads.items():
                    out += [tensor, len(dims)]
                    for dim in dims:
                        out.append(len(dim))
                        for li in dim:
                            _s_li(li, out)
                out.append(len(sc.write))
                for d in sc.write:
                    _s_li(d, out)
            out.append("\n")
        for k in t.kernels:
            _s_kernel(k, out)
    out += ["end", "\n"]
    return " ".join(str(x) for x in out)
