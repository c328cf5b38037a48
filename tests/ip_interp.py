"""An independent, sequential interpreter for the device programs of the generic loop-nest kernel
(exprgrad_b200/csrc/interp.hpp `IpProgram`, dumped by egb_program_lower_dump): test infrastructure that lets the CPU
tier check what csrc/lower.cpp produces - loop partition, flattened affine accesses, per-point index arithmetic, the
register program with its literal pool, constant folding - against the oracle without a GPU. It restates the semantics
of csrc/interp.cu in its bit-exact configuration (one thread per output point, reduction loops in nesting order, fp32
operation by operation) and nothing else: no thread layout, no vector paths."""
import struct

import numpy as np

(NOP, FADD, FSUB, FMUL, FDIV, FNEG, SIN, COS, EXP, LN, SQRT, POW, LOG10, LOG2, LOGB, IADD, ISUB, IMUL, IDIV, IMOD, IWRAP,
 INEG, FEQ, FLT, FLE, IEQ, ILT, ILE, BEQ, AND, OR, SELECT, TOSCALAR, TOINDEX, ARRAY_READ) = range(35)

_F = struct.Struct("<f")
_I = struct.Struct("<I")
_Q = struct.Struct("<q")
_UQ = struct.Struct("<Q")
MASK = (1 << 64) - 1


def _f(u):          # low word of a slot as fp32
    return np.float32(_F.unpack(_I.pack(u & 0xFFFFFFFF))[0])


def _fu(x):         # fp32 -> slot with a zero high word (the kernel clears the slot before it stores .f)
    return _I.unpack(_F.pack(np.float32(x)))[0]


def _i(u):          # slot as int64
    return _Q.unpack(_UQ.pack(u & MASK))[0]


def _iu(v):         # int64 (wrapping) -> slot
    return v & MASK


def _cdiv(a, d):    # C++ integer division truncates towards zero
    q = abs(a) // abs(d)
    return q if (a < 0) == (d < 0) else -q


def run_instrs(instrs, s, array_table):
    with np.errstate(all="ignore"):
        for op, dst, a, b, c, imm in instrs:
            if op == FADD: r = _fu(_f(s[a]) + _f(s[b]))
            elif op == FSUB: r = _fu(_f(s[a]) - _f(s[b]))
            elif op == FMUL: r = _fu(_f(s[a]) * _f(s[b]))
            elif op == FDIV: r = _fu(_f(s[a]) / _f(s[b]))
            elif op == FNEG: r = _fu(np.float32(0.0) - _f(s[a]))
            elif op == SIN: r = _fu(np.sin(_f(s[a])))
            elif op == COS: r = _fu(np.cos(_f(s[a])))
            elif op == EXP: r = _fu(np.exp(_f(s[a])))
            elif op == LN: r = _fu(np.log(_f(s[a])))
            elif op == SQRT: r = _fu(np.sqrt(_f(s[a])))
            elif op == POW: r = _fu(np.power(_f(s[a]), _f(s[b])))
            elif op == LOG10: r = _fu(np.log10(_f(s[a])))
            elif op == LOG2: r = _fu(np.log2(_f(s[a])))
            elif op == LOGB: r = _fu(np.log(_f(s[a])) / np.log(_f(s[b])))
            elif op == IADD: r = _iu(_i(s[a]) + _i(s[b]))
            elif op == ISUB: r = _iu(_i(s[a]) - _i(s[b]))
            elif op == IMUL: r = _iu(_i(s[a]) * _i(s[b]))
            elif op == IDIV:
                d = _i(s[b]); r = _iu(_cdiv(_i(s[a]), d)) if d else 0
            elif op == IMOD:
                d = _i(s[b]); x = _i(s[a]); r = _iu(x - d * _cdiv(x, d)) if d else 0
            elif op == IWRAP:
                d = _i(s[b]); x = _i(s[a])
                if d:
                    m = x - d * _cdiv(x, d)
                    m = m + d
                    r = _iu(m - d * _cdiv(m, d))
                else:
                    r = 0
            elif op == INEG: r = _iu(0 - _i(s[a]))
            elif op == FEQ: r = int(_f(s[a]) == _f(s[b]))
            elif op == FLT: r = int(_f(s[a]) < _f(s[b]))
            elif op == FLE: r = int(_f(s[a]) <= _f(s[b]))
            elif op == IEQ: r = int(_i(s[a]) == _i(s[b]))
            elif op == ILT: r = int(_i(s[a]) < _i(s[b]))
            elif op == ILE: r = int(_i(s[a]) <= _i(s[b]))
            elif op == BEQ: r = int((s[a] != 0) == (s[b] != 0))
            elif op == AND: r = int((s[a] != 0) and (s[b] != 0))
            elif op == OR: r = int((s[a] != 0) or (s[b] != 0))
            elif op == SELECT: r = s[b] if s[a] != 0 else s[c]
            elif op == TOSCALAR: r = _fu(np.float32(_i(s[a])))
            elif op == TOINDEX: r = _iu(int(np.trunc(np.float64(_f(s[a])))))
            elif op == ARRAY_READ: r = s[array_table[imm + _i(s[a])]]
            else: r = 0
            s[dst] = r


def _flat(op, s):
    idx = op["offset"]
    for slot, coef in op["terms"]:
        idx += coef * _i(s[slot])
    return idx


def _decode(loops, lo, hi, lin, s):
    for l in range(hi - 1, lo - 1, -1):
        start, step, count, slot = loops[l]
        q = lin // count
        s[slot] = _iu(start + step * (lin - q * count))
        lin = q


def run_program(p, tensors):
    """Execute one dumped program on `tensors` (tensor id -> flat float32 array, updated in place)."""
    s = [0] * 256
    for slot, bits in p["lits"]:
        s[slot] = int(bits)
    loops, npar = p["loops"], p["npar"]
    out = tensors[p["write"]["tensor"]]
    for point in range(p["npoints"]):
        _decode(loops, 0, npar, point, s)
        acc = np.float32(0.0)
        widx = 0
        if not p["scatter"] and p["nred"] > 0:
            _decode(loops, npar, len(loops), 0, s)
            run_instrs(p["index_instrs"], s, p["array_table"])
            widx = _flat(p["write"], s)
            assert 0 <= widx < out.size, f"kernel {p['kernel']}: write at {widx} of {out.size}"
            if p["accumulate"]:
                acc = out[widx]
        for r in range(p["nred"]):
            _decode(loops, npar, len(loops), r, s)
            run_instrs(p["index_instrs"], s, p["array_table"])
            for rd in p["reads"]:
                src = tensors[rd["tensor"]]
                idx = _flat(rd, s)
                # numpy would wrap a negative index silently; the device would read outside the tensor
                assert 0 <= idx < src.size, f"kernel {p['kernel']}: read of tensor {rd['tensor']} at {idx} of {src.size}"
                s[rd["dst"]] = _fu(src[idx])
            run_instrs(p["instrs"], s, p["array_table"])
            v = _f(s[p["write"]["dst"]])
            if p["scatter"]:
                w = _flat(p["write"], s)
                assert 0 <= w < out.size, f"kernel {p['kernel']}: scatter write at {w} of {out.size}"
                out[w] = out[w] + v if p["accumulate"] else v
            else:
                acc = np.float32(acc + v)
        if not p["scatter"] and p["nred"] > 0:
            out[widx] = acc


def run_target(prog, target, inputs, state, strict=True, epoch=0, cache=None):
    """Run every kernel of `target` in order, the way the reference does (model.nim:275-318, 385-411): results start
    at zero, every kernel accumulates. `inputs`: name -> array, `state`: tensor id -> array of the parameters / caches
    (updated in place). Returns the target's output tensor. `cache`: optional dict that keeps the dumped programs per
    (target, input shapes, epoch) across calls."""
    shapes = {k: list(v.shape) for k, v in inputs.items()}
    key = (target, tuple(sorted((k, tuple(v)) for k, v in shapes.items())), strict, epoch)
    if cache is not None and key in cache:
        dump = cache[key]
    else:
        dump = prog.lower_dump(target, shapes, strict=strict, epoch=epoch)
        if cache is not None:
            cache[key] = dump
    names = {}
    for tid in range(1, prog.tensor_count() + 1):
        info = prog.tensor_info(tid)
        if info["kind"] == "input":
            names[tid] = info["name"]
    tensors, tshape = {}, {}
    used = set()
    for p in dump:
        used.add(p["write"]["tensor"])
        used.update(rd["tensor"] for rd in p["reads"])
    out_id = prog.target_output(target)
    if out_id:
        used.add(out_id)
    for tid in sorted(used):
        info = prog.tensor_info(tid)
        if info["kind"] == "input":
            arr = np.ascontiguousarray(inputs[info["name"]], np.float32)
            tshape[tid] = list(arr.shape)
            tensors[tid] = arr.reshape(-1).copy()
        elif info["kind"] in ("param", "cache"):
            tshape[tid] = list(state[tid].shape)
            tensors[tid] = np.ascontiguousarray(state[tid], np.float32).reshape(-1).copy()
        else:
            tshape[tid] = prog.infer_shapes(target, shapes, tensor_id=tid)
            tensors[tid] = np.zeros(int(np.prod(tshape[tid])) if tshape[tid] else 1, np.float32)
    for p in dump:
        run_program(p, tensors)
    for tid in state:
        if tid in tensors:
            state[tid] = tensors[tid].reshape(tshape[tid]).copy()
    return tensors[out_id].reshape(tshape[out_id]) if out_id else None
