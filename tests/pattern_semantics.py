"""What a matched kernel pattern MEANS, written out in numpy (test infrastructure): the contraction, the three conv2
kernels and every map form of csrc/pattern.cpp, applied to the operands and constants the planner would launch the
specialised device kernel with (the `gemm` / `conv2` / `eltwise` records of egb_program_lower_dump). Compared, kernel by
kernel, with the generic lowering of the same IR kernel run by tests/ip_interp.py: a matcher that binds the wrong
operand, transposes the wrong side, drops a captured literal or mistakes an access pattern cannot pass."""
import numpy as np

f32 = np.float32


def _elt(form, x, y, p, epoch_unused=None):
    """one map form on fp32 arrays, operation by operation like EltFn<KIND> in csrc/eltwise_stream.cu"""
    zero, one = f32(0.0), f32(1.0)
    with np.errstate(all="ignore"):
        if form in ("copy", "bias-row-add"): return x
        if form == "relu": return np.where(zero <= x, x, zero)
        if form == "leakyRelu": return np.where(zero <= x, one, p[0]) * x
        if form == "sigmoid": return one / (one + np.exp(zero - x))
        if form == "tanh":
            e, f = np.exp(x), np.exp(zero - x)
            return (e - f) / (e + f)
        if form == "scale": return x * p[0]
        if form == "sgd-axpy": return (zero - x) * p[0]
        if form == "div-const": return x / p[0]
        if form == "add": return x + y
        if form == "sub": return x - y
        if form == "mul": return x * y
        if form == "relu-adjoint": return np.where(zero <= x, y, zero)
        if form == "leakyRelu-adjoint": return y * np.where(zero <= x, one, p[0])
        if form == "sigmoid-adjoint":
            e = np.exp(zero - x); s = one + e
            return zero - ((f32(-1.0) * (y / (s * s))) * e)
        if form == "tanh-adjoint":
            e, f = np.exp(x), np.exp(zero - x); s = e + f
            t = (zero - (e - f)) * (y / (s * s)); q = y / s
            return (zero - ((t + (zero - q)) * f)) + ((t + q) * e)
        if form == "adam-m": return (x * p[0]) + (p[1] * y)
        if form == "adam-v": return (x * p[0]) + (p[1] * (y * y))
        if form == "adam-step": return (p[0] * (x / p[1])) / (np.sqrt(y / p[2]) + p[3])
        if form == "square-adjoint": return (y * x) + (y * x)
    raise KeyError(form)


def launch_constants(form, lit, epoch):
    """csrc/runtime.cpp build_eltwise_node: constant sub-expressions are folded in float64 and rounded to fp32
    (passes.nim:1629-1650, llvmgen.nim:213-218); pow(b, epoch) is a run-time fp32 value"""
    if form in ("adam-m", "adam-v"):
        return [f32(lit[0] - 1.0), f32(1.0 - lit[0]), f32(0), f32(0)]
    if form == "adam-step":
        return [f32(0.0 - lit[0]), f32(1.0) - np.power(f32(lit[1]), f32(epoch)), f32(1.0) - np.power(f32(lit[2]), f32(epoch)),
                f32(lit[3])]
    return [f32(v) for v in lit]


def apply_eltwise(rec, write_tensor, tensors, epoch):
    """out += form(in0, in1) on flat fp32 tensors (every kernel of the dump accumulates)"""
    n = rec["n"]
    out = tensors[write_tensor]
    ops = []
    for rd in rec["reads"]:
        t = tensors[rd["tensor"]]
        if rd["scalar"]:
            ops.append(np.full(n, t[rd["offset"]], f32))
        elif rd["row"]:
            ops.append(np.tile(t[:rec["row"]], n // rec["row"]))
        else:
            ops.append(t[:n].copy())
    x = ops[0]
    y = ops[1] if len(ops) > 1 else None
    p = launch_constants(rec["form"], rec["lit"], epoch)
    out[:n] = out[:n] + _elt(rec["form"], x, y, p).astype(f32)


def apply_gemm(g, tensors):
    """C[M,N] += op(A) op(B), row-major with leading dimensions (float64 products: compared with a tolerance)"""
    def mat(tid, rows, cols, ld):
        return tensors[tid][:rows * ld].reshape(rows, ld)[:, :cols].astype(np.float64)
    a = mat(g["a"], g["K"], g["M"], g["lda"]).T if g["ta"] else mat(g["a"], g["M"], g["K"], g["lda"])
    b = mat(g["b"], g["N"], g["K"], g["ldb"]).T if g["tb"] else mat(g["b"], g["K"], g["N"], g["ldb"])
    c = tensors[g["c"]][:g["M"] * g["ldc"]].reshape(g["M"], g["ldc"])
    c[:, :g["N"]] = (c[:, :g["N"]].astype(np.float64) + a @ b).astype(f32)


def apply_conv2(c, tensors):
    """dnn.nim:45-49 and its two adjoints on NHWC / [F,KH,KW,C] tensors (valid convolution), float64"""
    N, H, W, C, F, KH, KW = (c[k] for k in ("N", "H", "W", "C", "F", "KH", "KW"))
    OH, OW = H - KH + 1, W - KW + 1
    img = tensors[c["img"]].reshape(N, H, W, C)
    fil = tensors[c["fil"]].reshape(F, KH, KW, C)
    out = tensors[c["out"]].reshape(N, OH, OW, F)
    acc = {0: out, 1: fil, 2: img}[c["kind"]].astype(np.float64)
    i64, f64, o64 = img.astype(np.float64), fil.astype(np.float64), out.astype(np.float64)
    for dy in range(KH):
        for dx in range(KW):
            patch = i64[:, dy:dy + OH, dx:dx + OW, :]                      # [N, OH, OW, C]
            if c["kind"] == 0:
                acc += np.einsum("nyxc,fc->nyxf", patch, f64[:, dy, dx, :])
            elif c["kind"] == 1:
                acc[:, dy, dx, :] += np.einsum("nyxf,nyxc->fc", o64, patch)
            else:
                acc[:, dy:dy + OH, dx:dx + OW, :] += np.einsum("nyxf,fc->nyxc", o64, f64[:, dy, dx, :])
    {0: out, 1: fil, 2: img}[c["kind"]][...] = acc.astype(f32)
