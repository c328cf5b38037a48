"""TEST INFRASTRUCTURE ONLY - the oracle's contraction kernel through an LLVM JIT (llvmlite / MCJIT), shaped like
the reference's own code generator, to show that the gcc-compiled restatement (oracle/cgen.py) is representative of a
JIT-compiled path and not of gcc.

What is mirrored (exprgrad/llvmgen.nim):
  * loops as cond / body / end / incr basic blocks with a phi for the iterator and an `icmp eq iter, stop` exit test
    (llvmgen.nim:320-360);
  * tensor accesses as in-bounds GEPs with 4-byte aligned loads and stores, `++=` as load - fadd - store
    (llvmgen.nim:277-301); index arithmetic as nsw adds / muls (llvmgen.nim:219-221);
  * no fast-math flags anywhere (llvm.nim:486-491), so fmul + fadd are never contracted into an fma and the result is
    bit-identical to the gcc build with -ffp-contract=off (checked by tests/test_oracle_jit.py);
  * the optimisation pipeline: `default<O3>` on a target machine for the host CPU with the host's feature string
    (llvmgen.nim:616-647);
  * the loop order the reference's scheduling passes produce for c[y, x] ++= a[y, it] * b[it, x]: y (parallel), it, x
    (passes.nim:700-745, 1794-1823), rows split contiguously over the threads (model.nim:110-132).
Only bench.py's cpu_baseline leg and tests/ may import this module."""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_ENGINE = None
_FN = None


def available():
    try:
        import llvmlite.binding  # noqa: F401
        import llvmlite.ir  # noqa: F401
        return True
    except Exception:
        return False


def _loop(builder, fn, start, stop, name):
    """cond / body / end / incr blocks of llvmgen.nim's InstrLoop; returns (iter phi, body builder hook, closer)."""
    from llvmlite import ir
    header = builder.block
    cond = fn.append_basic_block(name + "_cond")
    body = fn.append_basic_block(name + "_body")
    end = fn.append_basic_block(name + "_end")
    incr = fn.append_basic_block(name + "_incr")
    builder.branch(cond)
    builder.position_at_end(cond)
    it = builder.phi(ir.IntType(64), name="iter_" + name)
    builder.cbranch(builder.icmp_signed("==", it, stop, name="exitcond"), end, body)
    builder.position_at_end(body)

    def close():
        builder.branch(incr)
        builder.position_at_end(incr)
        nxt = builder.add(it, ir.Constant(ir.IntType(64), 1), name="incr_iter")
        builder.branch(cond)
        it.add_incoming(start, header)
        it.add_incoming(nxt, incr)
        builder.position_at_end(end)

    return it, close


def matmul_ir():
    """LLVM IR of  c[y, x] ++= a[y, it] * b[it, x]  for rows y0 <= y < y1 (loops y, it, x)."""
    from llvmlite import ir
    i64, f32 = ir.IntType(64), ir.FloatType()
    fp = f32.as_pointer()
    mod = ir.Module(name="exprgrad_oracle_jit")
    fn = ir.Function(mod, ir.FunctionType(ir.VoidType(), [fp, fp, fp, i64, i64, i64, i64]), name="matmul_rows")
    a, b, c, y0, y1, K, N = fn.args
    for p in (a, b, c):
        p.add_attribute("noalias")   # distinct tensors (the reference passes them as separate buffers of the model)
    builder = ir.IRBuilder(fn.append_basic_block("entry"))
    zero = ir.Constant(i64, 0)
    y, close_y = _loop(builder, fn, y0, y1, "y")
    it, close_it = _loop(builder, fn, zero, K, "it")
    x, close_x = _loop(builder, fn, zero, N, "x")
    a_idx = builder.add(builder.mul(y, K, flags=["nsw"]), it, flags=["nsw"])
    b_idx = builder.add(builder.mul(it, N, flags=["nsw"]), x, flags=["nsw"])
    c_idx = builder.add(builder.mul(y, N, flags=["nsw"]), x, flags=["nsw"])
    av = builder.load(builder.gep(a, [a_idx], inbounds=True, name="value_ptr"), align=4)
    bv = builder.load(builder.gep(b, [b_idx], inbounds=True, name="value_ptr"), align=4)
    prod = builder.fmul(av, bv)
    cp = builder.gep(c, [c_idx], inbounds=True, name="value_ptr")
    builder.store(builder.fadd(builder.load(cp, align=4), prod, name="new_value"), cp, align=4)
    close_x()
    close_it()
    close_y()
    builder.ret_void()
    return str(mod)


def _compile():
    global _ENGINE, _FN
    if _FN is not None:
        return _FN
    import llvmlite.binding as llvm
    try:
        llvm.initialize()
    except Exception:
        pass   # newer llvmlite initialises on import
    llvm.initialize_native_target()
    llvm.initialize_native_asmprinter()
    target = llvm.Target.from_default_triple()
    tm = target.create_target_machine(cpu=llvm.get_host_cpu_name(), features=llvm.get_host_cpu_features().flatten(), opt=3)
    mod = llvm.parse_assembly(matmul_ir())
    mod.verify()
    # default<O3> through the new pass manager (what run_passes("default<O3>") does in llvmgen.nim:636-645)
    pto = llvm.create_pipeline_tuning_options(speed_level=3, size_level=0)
    pto.loop_vectorization = True
    pto.slp_vectorization = True
    pto.loop_unrolling = True
    pb = llvm.create_pass_builder(tm, pto)
    pb.getModulePassManager().run(mod, pb)
    _ENGINE = llvm.create_mcjit_compiler(mod, tm)
    _ENGINE.finalize_object()
    addr = _ENGINE.get_function_address("matmul_rows")
    fp = ctypes.POINTER(ctypes.c_float)
    _FN = ctypes.CFUNCTYPE(None, fp, fp, fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64)(addr)
    return _FN


def matmul(a, b, threads=None, out=None):
    """c = a @ b with the JIT-compiled reference loop nest; rows split contiguously over `threads` workers."""
    fn = _compile()
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    c = np.zeros((M, N), np.float32) if out is None else out   # zero-filled result, as model.nim:295-300
    if out is not None:
        c[...] = 0
    fp = ctypes.POINTER(ctypes.c_float)
    pa, pb_, pc = (v.ctypes.data_as(fp) for v in (a, b, c))
    threads = max(1, min(threads or (os.cpu_count() or 1), M))
    bounds = [M * t // threads for t in range(threads + 1)]
    if threads == 1:
        fn(pa, pb_, pc, 0, M, K, N)
    else:
        with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL for the duration of the call
            list(ex.map(lambda t: fn(pa, pb_, pc, bounds[t], bounds[t + 1], K, N), range(threads)))
    return c
