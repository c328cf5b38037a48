"""ORACLE - TEST INFRASTRUCTURE ONLY. Never imported by the product (exprgrad_b200/).

Restatement of the middle-end passes that define the semantics of exprgrad's compiled path
(pipeline order: exprgrad/model.nim:46-77). Only the passes that change WHAT is computed are restated;
CPU/OpenCL scheduling passes (fuse/tile/cache/LICM) cannot change fp results because fast-math is off
(exprgrad/wrappers/llvm.nim:486-491) and are omitted.

  fold_linear_indices      passes.nim:195-253
  dead_code_elim           passes.nim:268-329
  deduplicate_reads        passes.nim:352-381
  infer_shape_constraints  passes.nim:1040-1117
  generate / derive        passes.nim:383-549 (instr rules), 519-549 (kernel), 558-698 (driver)
  dead_kernel_elim         passes.nim:331-350
  infer_loop_bounds        passes.nim:1001-1038
  identify_independent     passes.nim:1774-1781
  collect_tensors          passes.nim:936-967
  sort_shape_constraints   passes.nim:1119-1221
  reorder_loops            passes.nim:700-745
  solve / eval / infer_shapes  passes.nim:1252-1436 (run-time, integer-exact)
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Dict, List

from .ir import (PRIO_CONDITION, PRIO_INFERRED, PRIO_USER, GradientError, Instr, Kernel, LinearIndex, Loop,
                 Program, ShapeConstraint, ShapeError, Target, TensorDef, TensorOp)


# ----------------------------------------------------------------------------- folding / DCE / dedup

def _fold_setup(index: LinearIndex, kernel: Kernel) -> LinearIndex:
    regs: Dict[int, LinearIndex] = {}
    for loop in kernel.loops:
        regs[loop.iter] = LinearIndex.reg(loop.iter)
    for ins in index.setup:
        k = ins.kind
        if k == "Index":
            regs[ins.res] = LinearIndex.const(ins.lit)
        elif k == "Add":
            regs[ins.res] = regs[ins.args[0]] + regs[ins.args[1]]
        elif k == "Sub":
            regs[ins.res] = regs[ins.args[0]] - regs[ins.args[1]]
        elif k == "Mul":
            try:
                regs[ins.res] = regs[ins.args[0]].mul(regs[ins.args[1]])
            except ValueError:
                regs[ins.res] = LinearIndex.reg(ins.res)
        elif k == "Negate":
            regs[ins.res] = -regs[ins.args[0]]
        else:
            regs[ins.res] = LinearIndex.reg(ins.res)
    total = LinearIndex()
    for reg, factor in index.factors.items():
        total = total + regs[reg].scale(factor)
    total.setup = []
    used = set(total.factors.keys())
    kept = []
    for ins in reversed(index.setup):
        if ins.res in used:
            kept.append(ins)
            used.update(ins.args)
    total.setup = list(reversed(kept))
    return total


def fold_linear_indices(kernel: Kernel):
    for loop in kernel.loops:
        loop.start = _fold_setup(loop.start, kernel)
        loop.stop = _fold_setup(loop.stop, kernel)
    for op in list(kernel.reads) + [kernel.write]:
        op.dims = [_fold_setup(d, kernel) for d in op.dims]


def _dce_instrs(instrs: List[Instr], used: set) -> List[Instr]:
    out = []
    for ins in reversed(instrs):
        if ins.res != 0 and ins.res in used:
            used.update(ins.args)
            out.append(ins)
    return list(reversed(out))


def _dce_index(idx: LinearIndex, used: set):
    used.update(idx.factors.keys())
    idx.setup = _dce_instrs(idx.setup, used)


def dead_code_elim(kernel: Kernel):
    if kernel.generator.kind != "None":
        return
    used = set()
    if kernel.write.data:
        used.add(kernel.write.data)
    for d in kernel.write.dims:
        _dce_index(d, used)
    kernel.instrs = _dce_instrs(kernel.instrs, used)
    kept = []
    for r in kernel.reads:
        if r.data in used:
            for d in r.dims:
                _dce_index(d, used)
            kept.append(r)
    kernel.reads = kept
    for loop in reversed(kernel.loops):
        _dce_index(loop.start, used)
        _dce_index(loop.stop, used)


def deduplicate_reads(kernel: Kernel):
    unique: Dict[tuple, int] = {}
    subs: Dict[int, int] = {}
    kept = []
    for r in kernel.reads:
        key = r.key_without_data()
        if key in unique:
            subs[r.data] = unique[key]
        else:
            unique[key] = r.data
            kept.append(r)
    kernel.reads = kept
    for ins in kernel.instrs:
        ins.args = [subs.get(a, a) for a in ins.args]
    kernel.res = subs.get(kernel.res, kernel.res)
    kernel.write.data = subs.get(kernel.write.data, kernel.write.data)


def _all_kernels(kernel: Kernel):
    yield kernel
    if kernel.custom_grad is not None:
        for g in kernel.custom_grad["kernels"]:
            yield g


# ----------------------------------------------------------------------------- shape constraints

def _simplify_max_index(indices: List[LinearIndex]) -> List[LinearIndex]:
    """Per dim keep, for every distinct factor table, the largest constant (passes.nim:1040-1057)."""
    max_const: Dict[tuple, int] = {}
    order = []
    complex_ = []
    for idx in indices:
        if not idx.setup:
            key = tuple(sorted(idx.factors.items()))
            if key not in max_const:
                max_const[key] = idx.constant
                order.append(key)
            else:
                max_const[key] = max(max_const[key], idx.constant)
        else:
            complex_.append(idx)
    return complex_ + [LinearIndex(factors=dict(k), constant=max_const[k]) for k in order]


def kernel_shape_constraints(kernel: Kernel) -> List[ShapeConstraint]:
    out = []
    if kernel.write.is_raw:
        if len(kernel.reads) == 1:
            out.append(ShapeConstraint("Copy", kernel.write.tensor, PRIO_INFERRED, src=kernel.reads[0].tensor))
    else:
        lin = ShapeConstraint("Linear", kernel.write.tensor, PRIO_INFERRED)
        for op in kernel.reads:
            if not op.is_raw:
                if op.tensor not in lin.reads:
                    lin.reads[op.tensor] = [[] for _ in op.dims]
                for i, d in enumerate(op.dims):
                    lin.reads[op.tensor][i].append(d)
        lin.write = list(kernel.write.dims)
        for t in lin.reads:
            lin.reads[t] = [_simplify_max_index(d) for d in lin.reads[t]]
        out.append(lin)
    for _, op in kernel.tensor_ops():
        if not op.is_raw:
            out.append(ShapeConstraint("Rank", op.tensor, PRIO_CONDITION, rank=len(op.dims)))
    return out


def infer_shape_constraints(prog: Program):
    for target in prog.targets.values():
        for tid in prog.caches:
            target.shapes.append(ShapeConstraint("Copy", tid, PRIO_INFERRED, src=prog.tdef(tid).cache))
        for kernel in target.kernels:
            if kernel.generator.kind == "None":
                target.shapes.extend(kernel_shape_constraints(kernel))


# ----------------------------------------------------------------------------- autodiff

def derive_instrs(instrs: List[Instr], kernel: Kernel, grad_regs: Dict[int, int]) -> List[Instr]:
    """Reverse sweep over the expression (passes.nim:383-517). One adjoint register per primal register;
    a register used twice receives the sum of both contributions."""
    out: List[Instr] = []
    new = kernel.alloc_reg

    def emit(kind, args=(), lit=None):
        r = new()
        out.append(Instr(kind, list(args), r, lit=lit))
        return r

    for ins in reversed(instrs):
        if ins.res not in grad_regs:
            continue
        g = grad_regs[ins.res]
        a = ins.args
        k = ins.kind
        ga = None
        if k == "Add":
            ga = [g, g]
        elif k == "Sub":
            ga = [g, emit("Negate", [g])]
        elif k == "Mul":
            ga = [emit("Mul", [g, a[1]]), emit("Mul", [g, a[0]])]
        elif k == "Div":
            grad_a = emit("Div", [g, a[1]])
            sq_y = emit("Mul", [a[1], a[1]])
            div_g = emit("Div", [g, sq_y])
            neg_x = emit("Negate", [a[0]])
            ga = [grad_a, emit("Mul", [neg_x, div_g])]
        elif k == "Negate":
            ga = [emit("Negate", [g])]
        elif k in ("Ln", "Log10", "Log2"):
            base = {"Ln": 1.0, "Log10": math.log(10.0), "Log2": math.log(2.0)}[k]
            den = a[0]
            if base != 1.0:
                f = emit("Scalar", lit=base)
                den = emit("Mul", [a[0], f])
            ga = [emit("Div", [g, den])]
        elif k == "Log":
            log_y = emit("Ln", [a[1]])
            mul = emit("Mul", [a[0], log_y])
            gx = emit("Div", [g, mul])
            log_x = emit("Ln", [a[0]])
            neg_log_x = emit("Negate", [log_x])
            log_y_sq = emit("Mul", [log_y, log_y])
            den = emit("Mul", [a[1], log_y_sq])
            num = emit("Mul", [g, neg_log_x])
            ga = [gx, emit("Div", [num, den])]
        elif k == "Exp":
            ga = [emit("Mul", [g, ins.res])]
        elif k == "Sin":
            c = emit("Cos", [a[0]])
            ga = [emit("Mul", [c, g])]
        elif k == "Cos":
            s = emit("Sin", [a[0]])
            ns = emit("Negate", [s])
            ga = [emit("Mul", [ns, g])]
        elif k == "Select":
            zero = emit("Scalar", lit=0.0)
            ga = [0, emit("Select", [a[0], g, zero]), emit("Select", [a[0], zero, g])]
        elif k == "Sqrt":
            two = emit("Scalar", lit=2.0)
            den = emit("Mul", [two, ins.res])
            ga = [emit("Div", [g, den])]
        elif k == "Pow":
            one = emit("Scalar", lit=1.0)
            new_exp = emit("Sub", [a[1], one])
            p = emit("Pow", [a[0], new_exp])
            pf = emit("Mul", [a[1], p])
            g_base = emit("Mul", [g, pf])
            lg = emit("Ln", [a[0]])
            prod = emit("Mul", [ins.res, lg])
            ga = [g_base, emit("Mul", [g, prod])]
        elif k in ("ToScalar", "ToIndex"):
            ga = [0]
        else:
            ga = []
        if len(ga) != len(a):
            raise GradientError("Unable to derive " + k)
        for arg, garg in zip(a, ga):
            if garg != 0:
                if arg in grad_regs:
                    grad_regs[arg] = emit("Add", [grad_regs[arg], garg])
                else:
                    grad_regs[arg] = garg
    return out


def derive_kernel(kernel: Kernel, grad_tensors: Dict[int, int]) -> List[Kernel]:
    """One adjoint kernel per read that receives gradient (passes.nim:519-549)."""
    base = kernel.clone()
    grad_regs: Dict[int, int] = {}
    write_grad = base.alloc_reg()
    base.reads.append(TensorOp(grad_tensors[kernel.write.tensor], kernel.write.is_raw,
                               [d.clone() for d in kernel.write.dims], write_grad))
    grad_regs[kernel.write.data] = write_grad
    base.instrs = base.instrs + derive_instrs(kernel.instrs, base, grad_regs)
    out = []
    for read in kernel.reads:
        if read.data in grad_regs:
            gk = base.clone()
            gk.res = grad_regs[read.data]
            gk.write = TensorOp(grad_tensors[read.tensor], read.is_raw, [d.clone() for d in read.dims],
                                grad_regs[read.data])
            dead_code_elim(gk)
            out.append(gk)
    return out


def _iota_kernel(src_len_tensor: int, write_tensor: int, read_tensor: int = 0, lit=None) -> Kernel:
    """`dst{i} = 1.0` (seed) or `dst{i} = src{i}` (reshape) over len(src) (passes.nim:574-600, 643-673)."""
    k = Kernel()
    k.nregs = 3
    stop = LinearIndex([Instr("Len", res=3, tensor=src_len_tensor)], {3: 1})
    k.loops = [Loop(2, True, LinearIndex.const(0), stop, 1)]
    if read_tensor:
        k.reads = [TensorOp(read_tensor, True, [LinearIndex.reg(2)], 1)]
    else:
        k.instrs = [Instr("Scalar", res=1, lit=lit)]
    k.res = 1
    k.write = TensorOp(write_tensor, True, [LinearIndex.reg(2)], 1)
    return k


def generate(prog: Program):
    for target in prog.targets.values():
        it = 0
        while it < len(target.kernels):
            kernel = target.kernels[it]
            gk = kernel.generator.kind
            if gk == "Backwards":
                grad_tensors: Dict[int, int] = {}
                grad_kernels: List[Kernel] = []
                loss = kernel.generator.tensor
                grad_loss = prog.alloc_tensor(TensorDef("Result"))
                grad_kernels.append(_iota_kernel(loss, grad_loss, lit=1.0))
                target.shapes.append(ShapeConstraint("Copy", grad_loss, PRIO_INFERRED, src=loss))
                grad_tensors[loss] = grad_loss
                for k2 in target.kernels[it + 1:]:
                    if k2.generator.kind == "Gradient":
                        grad_tensors[k2.generator.tensor] = k2.write.tensor
                        target.shapes.append(ShapeConstraint("Copy", k2.write.tensor, PRIO_INFERRED,
                                                             src=k2.generator.tensor))
                for k2 in reversed(target.kernels[:it]):
                    for read in k2.reads:
                        if read.tensor not in grad_tensors:
                            gt = prog.alloc_tensor(TensorDef("Result"))
                            target.shapes.append(ShapeConstraint("Copy", gt, PRIO_INFERRED, src=read.tensor))
                            grad_tensors[read.tensor] = gt
                    if k2.custom_grad is not None:
                        subs = dict(k2.custom_grad["subs"])
                        for initial, g in k2.custom_grad["tensors"].items():
                            t = k2.custom_grad["subs"].get(initial, initial)
                            subs[g] = grad_tensors[t]
                        for ck in reversed(k2.custom_grad["kernels"]):
                            c = ck.clone()
                            c.substitute_tensors(subs)
                            grad_kernels.append(c)
                    else:
                        if k2.generator.kind != "None":
                            continue
                        if k2.write.tensor not in grad_tensors:
                            continue
                        grad_kernels.extend(derive_kernel(k2, grad_tensors))
                target.kernels[it:it + 1] = grad_kernels
                it += len(grad_kernels)
            elif gk == "Gradient":
                del target.kernels[it]
            elif gk == "Reshape":
                src = kernel.generator.tensor
                dst = kernel.write.tensor
                target.kernels[it] = _iota_kernel(src, dst, read_tensor=src)
                sc = ShapeConstraint("Dims", dst, PRIO_INFERRED)
                prod = 1
                for s in kernel.generator.reshape:
                    if s >= 0:
                        prod *= s
                for s in kernel.generator.reshape:
                    if s >= 0:
                        sc.dims.append(LinearIndex.const(s))
                    else:
                        sc.dims.append(LinearIndex([Instr("Len", res=1, tensor=src), Instr("Index", res=2, lit=prod),
                                                    Instr("IndexDiv", [1, 2], 3)], {3: 1}))
                target.shapes.append(sc)
                it += 1
            else:
                it += 1


def dead_kernel_elim(prog: Program):
    for target in prog.targets.values():
        used = {i + 1 for i, t in enumerate(prog.tensors) if t.kind != "Result"}
        if target.output:
            used.add(target.output)
        kept = []
        for kernel in reversed(target.kernels):
            if kernel.write.tensor in used:
                for r in kernel.reads:
                    used.add(r.tensor)
                kept.append(kernel)
        target.kernels = list(reversed(kept))


# ----------------------------------------------------------------------------- loops

def infer_loop_bounds(prog: Program):
    for target in prog.targets.values():
        for kernel in target.kernels:
            iters = {l.iter: l for l in kernel.loops if not l.has_bounds}
            for _, op in kernel.tensor_ops():
                for dim_idx, dim in enumerate(op.dims):
                    reg = dim.only_register()
                    if reg and reg in iters and not iters[reg].has_bounds:
                        loop = iters[reg]
                        loop.has_bounds = True
                        loop.start = LinearIndex.const(0)
                        size = kernel.alloc_reg()
                        if op.is_raw:
                            setup = [Instr("Len", res=size, tensor=op.tensor)]
                        else:
                            setup = [Instr("Shape", res=size, tensor=op.tensor, dim=dim_idx)]
                        loop.stop = LinearIndex(setup, {size: 1})
                        loop.step = 1


def identify_independent(prog: Program):
    for target in prog.targets.values():
        for kernel in target.kernels:
            indep = {d.only_register() for d in kernel.write.dims if d.only_register()}
            for loop in kernel.loops:
                if loop.iter in indep:
                    loop.mode = 1


def reorder_loops(kernel: Kernel):
    """Greedy topological order on "dim i-1 register -> dim i register" edges, reads weigh 10, the
    write 1 (passes.nim:700-745). Determines the nesting (= fp accumulation) order of reductions."""
    n = len(kernel.loops)
    loop_of = {l.iter: i for i, l in enumerate(kernel.loops)}
    graph = [{"read": [], "write": []} for _ in range(n)]
    for kind, op in kernel.tensor_ops():
        for i in range(1, len(op.dims)):
            for ra in op.dims[i - 1].factors:
                for rb in op.dims[i].factors:
                    if ra in loop_of and rb in loop_of:
                        graph[loop_of[ra]][kind].append(loop_of[rb])
    val = {"read": 10, "write": 1}
    scores = [0] * n
    for edges in graph:
        for kind, tg in edges.items():
            for t in tg:
                scores[t] += val[kind]
    closed = [False] * n
    order = []
    for _ in range(n):
        best, best_score = -1, 0
        for i, s in enumerate(scores):
            if not closed[i] and (best < 0 or s < best_score):
                best, best_score = i, s
        closed[best] = True
        order.append(best)
        for kind, tg in graph[best].items():
            for t in tg:
                scores[t] -= val[kind]
    kernel.loops = [kernel.loops[i] for i in order]


# ----------------------------------------------------------------------------- target bookkeeping

def _instr_tensors(instrs, out):
    for ins in instrs:
        if ins.tensor and ins.tensor not in out:
            out.append(ins.tensor)


def collect_tensors(prog: Program):
    for target in prog.targets.values():
        ts: List[int] = []
        for kernel in target.kernels:
            for _, op in kernel.tensor_ops():
                if op.tensor and op.tensor not in ts:
                    ts.append(op.tensor)
            for loop in kernel.loops:
                _instr_tensors(loop.start.setup, ts)
                _instr_tensors(loop.stop.setup, ts)
            _instr_tensors(kernel.instrs, ts)
        target.tensors = ts


def _underconstrained(sc: ShapeConstraint) -> bool:
    if sc.kind == "Rank":
        return sc.rank > 0
    if sc.kind in ("Dims", "Copy"):
        return False
    defined = set()
    for dims in sc.reads.values():
        for indices in dims:
            for idx in indices:
                defined.update(idx.factors.keys())
    for d in sc.write:
        for reg in d.factors:
            if reg not in defined:
                return True
    return False


def _deps(sc: ShapeConstraint):
    if sc.kind == "Dims":
        for d in sc.dims:
            for ins in d.setup:
                if ins.tensor:
                    yield ins.tensor
    elif sc.kind == "Linear":
        yield from sc.reads.keys()
    elif sc.kind == "Copy":
        yield sc.src


def sort_shape_constraints(prog: Program):
    for target in prog.targets.values():
        chosen: Dict[int, ShapeConstraint] = {}
        conditions = []
        for sc in target.shapes:
            if sc.dest not in chosen or chosen[sc.dest].priority < sc.priority:
                chosen[sc.dest] = sc  # first wins on ties (passes.nim:1182-1185)
            if sc.priority == PRIO_CONDITION:
                conditions.append(sc)
        for cond in conditions:
            if cond.dest not in chosen:
                continue
            sc = chosen[cond.dest]
            while sc.kind == "Copy" and sc.src in chosen and len(prog.tdef(sc.dest).shape) == 0:
                sc = chosen[sc.src]
            if sc.kind == "Copy" and len(prog.tdef(sc.dest).shape) == 0:
                chosen[sc.src] = cond
            else:
                if len(prog.tdef(sc.dest).shape) > 0:
                    rank = len(prog.tdef(sc.dest).shape)
                else:
                    rank = {"Dims": len(sc.dims), "Linear": len(sc.write), "Rank": sc.rank}.get(sc.kind, -1)
                if cond.rank != rank:
                    raise ShapeError(f"A condition requires that tensor{cond.dest - 1} has rank {cond.rank}, "
                                     f"but it has rank {rank}")
        order: List[ShapeConstraint] = []
        closed = set()

        def visit(tid):
            if prog.tdef(tid).kind in ("Result", "Cache", "Random") and tid not in closed:
                closed.add(tid)
                if tid not in chosen:
                    raise ShapeError(f"tensor{tid - 1} ({prog.tdef(tid).name}) requires shape")
                sc = chosen[tid]
                if _underconstrained(sc):
                    raise ShapeError(f"Shape for tensor{tid - 1} is underconstrained")
                for dep in _deps(sc):
                    visit(dep)
                order.append(sc)

        for tid in target.tensors:
            visit(tid)
        target.shapes = order


# ----------------------------------------------------------------------------- run-time shape inference

def solve(equations: List[LinearIndex]) -> Dict[int, Fraction]:
    """Exact solve of `index = 0` equations: first n distinct (normalised) rows, fraction-free
    elimination with partial pivoting, rational back-substitution (passes.nim:1252-1323)."""
    indices: Dict[int, int] = {}
    for eq in equations:
        for reg in eq.factors:
            if reg not in indices:
                indices[reg] = len(indices)
    n = len(indices)
    if n == 0:
        return {}
    if len(equations) < n:
        raise ValueError("Underconstrained linear system")
    rows = []
    known = set()
    for eq in equations:
        if not eq.factors:
            if eq.constant != 0:
                raise ValueError("No solution")
            continue
        row = [0] * (n + 1)
        for reg, f in eq.factors.items():
            row[indices[reg]] = f
        row[n] = -eq.constant
        first = 0
        norm = []
        for v in row:
            if first == 0:
                first = v
            norm.append(Fraction(0) if first == 0 else Fraction(v, first))
        key = tuple(norm)
        if key not in known:
            known.add(key)
            rows.append(row)
            if len(rows) >= n:
                break
    if len(rows) < n:
        raise ValueError("Underconstrained linear system")
    m = rows
    for pivot in range(n):
        max_row = pivot
        for y in range(pivot + 1, n):
            if abs(m[y][pivot]) > abs(m[max_row][pivot]):
                max_row = y
        if max_row != pivot:
            m[max_row], m[pivot] = m[pivot], m[max_row]
        tgt = m[pivot][pivot]
        for y in range(pivot + 1, n):
            cur = m[y][pivot]
            if cur != 0:
                m[y] = [m[y][x] * tgt - m[pivot][x] * cur for x in range(n + 1)]
    sol = [Fraction(0)] * n
    for y in range(n - 1, -1, -1):
        s = Fraction(m[y][n])
        for x in range(y + 1, n):
            s -= sol[x] * m[y][x]
        if m[y][y] == 0:
            raise ZeroDivisionError("singular shape system")
        sol[y] = s / m[y][y]
    return {reg: sol[i] for reg, i in indices.items()}


def _trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _trunc_mod(a: int, b: int) -> int:
    return a - _trunc_div(a, b) * b


def eval_index_instrs(instrs: List[Instr], shapes: Dict[int, List[int]], regs: Dict[int, int]) -> str:
    """Mini evaluator for shape expressions (passes.nim:1328-1374)."""
    for ins in instrs:
        if any(a not in regs for a in ins.args) or (ins.tensor and ins.tensor not in shapes):
            return "dynamic_reg"
        k = ins.kind
        if k == "Shape":
            shape = shapes[ins.tensor]
            if not shape:
                return "dynamic_shape"
            size = shape[len(shape) + ins.dim] if ins.dim < 0 else shape[ins.dim]
            if size < 0:
                return "dynamic_shape"
            regs[ins.res] = size
        elif k == "Len":
            shape = shapes[ins.tensor]
            if not shape or any(s < 0 for s in shape):
                return "dynamic_shape"
            regs[ins.res] = math.prod(shape)
        elif k == "ShapeLen":
            regs[ins.res] = len(shapes[ins.tensor])
        elif k == "Index":
            regs[ins.res] = ins.lit
        elif k == "Add":
            regs[ins.res] = regs[ins.args[0]] + regs[ins.args[1]]
        elif k == "Sub":
            regs[ins.res] = regs[ins.args[0]] - regs[ins.args[1]]
        elif k == "Mul":
            regs[ins.res] = regs[ins.args[0]] * regs[ins.args[1]]
        elif k == "IndexDiv":
            regs[ins.res] = _trunc_div(regs[ins.args[0]], regs[ins.args[1]])
        elif k == "Mod":
            regs[ins.res] = _trunc_mod(regs[ins.args[0]], regs[ins.args[1]])
        elif k == "Wrap":
            v = _trunc_mod(regs[ins.args[0]], regs[ins.args[1]])
            regs[ins.res] = v + regs[ins.args[1]] if v < 0 else v
        elif k == "Negate":
            regs[ins.res] = -regs[ins.args[0]]
        else:
            return "invalid"
    return "ok"


def _matches(static: List[int], shape: List[int]) -> bool:
    if not static:
        return True
    if len(static) != len(shape):
        return False
    return all(s < 0 or s == d for s, d in zip(static, shape))


def infer_shapes(prog: Program, target_name: str, inputs: Dict[int, List[int]]) -> Dict[int, List[int]]:
    """passes.nim:1386-1436. Integer-exact; the product must reproduce these bit for bit."""
    res: Dict[int, List[int]] = {}
    for tid, shape in inputs.items():
        res[tid] = list(shape)
        static = prog.tdef(tid).shape
        if not _matches(static, list(shape)):
            raise ShapeError(f"Given shape for tensor{tid - 1} is {list(shape)}, but its static shape is {static}")
    for tid in prog.params:
        res[tid] = list(prog.tdef(tid).shape)
    for sc in prog.targets[target_name].shapes:
        for dep in _deps(sc):
            if dep not in res:
                raise ShapeError(f"Missing shape for tensor{dep - 1}, maybe you forgot to pass an input to the model?")
        if sc.kind == "Rank":
            res[sc.dest] = [0] * sc.rank
        elif sc.kind == "Dims":
            sizes = []
            for idx in sc.dims:
                regs: Dict[int, int] = {}
                st = eval_index_instrs(idx.setup, res, regs)
                if st != "ok":
                    raise ShapeError({"dynamic_shape": "Not all shapes are known.",
                                      "invalid": "Invalid instruction in tensor shape",
                                      "dynamic_reg": "Unable to evaluate all instructions."}[st])
                sizes.append(idx.eval(regs))
            res[sc.dest] = sizes
        elif sc.kind == "Copy":
            res[sc.dest] = list(res[sc.src])
        elif sc.kind == "Linear":
            eqs = []
            for tid, dims in sc.reads.items():
                if len(dims) != len(res[tid]):
                    raise ShapeError(f"tensor{tid - 1} is read with {len(dims)} indices but has rank {len(res[tid])}")
                for d, indices in enumerate(dims):
                    assert len(indices) == 1
                    eqs.append(indices[0] - (res[tid][d] - 1))
            max_values = {reg: _trunc_div(v.numerator, v.denominator) for reg, v in solve(eqs).items()}
            res[sc.dest] = [idx.eval(max_values) + 1 for idx in sc.write]
    return res


# ----------------------------------------------------------------------------- pipeline

def compile_program(prog: Program):
    """The semantics-defining prefix of exprgrad/model.nim:46-77."""
    for i, t in enumerate(prog.tensors):  # makeTensorLookups (passes.nim:1745-1758)
        tid = i + 1
        if t.kind == "Param":
            prog.params.append(tid)
        elif t.kind == "Input":
            prog.inputs[t.name] = tid
        elif t.kind == "Cache":
            prog.caches.append(tid)
    for target in prog.targets.values():
        for kernel in target.kernels:
            for k in _all_kernels(kernel):
                dead_code_elim(k)
                fold_linear_indices(k)
                deduplicate_reads(k)
    infer_shape_constraints(prog)
    generate(prog)
    dead_kernel_elim(prog)
    infer_loop_bounds(prog)
    identify_independent(prog)
    dead_kernel_elim(prog)
    collect_tensors(prog)
    sort_shape_constraints(prog)
    # static shapes of caches (subset of inferStaticShapes, passes.nim:1444-1514)
    for tid in prog.caches:
        src = prog.tdef(tid).cache
        if not prog.tdef(src).shape or any(s < 0 for s in prog.tdef(src).shape):
            raise ShapeError(f'Shape of cache "{prog.tdef(tid).name}" must be inferred at compile time')
        prog.tdef(tid).shape = list(prog.tdef(src).shape)
    for target in prog.targets.values():
        for kernel in target.kernels:
            reorder_loops(kernel)
