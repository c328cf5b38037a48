"""CPU-only checks of the drop-in boundary: libegb200.so loads and exports every symbol that
include/egb200.h declares (no compute calls - there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "egb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(egb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "exprgrad_b200", "libegb200.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_declares_every_symbol():
    from exprgrad_b200 import _ffi
    declared = set(declared_symbols())
    bound = {n for n in dir(_ffi.lib) if n.startswith("egb_")} | set(_ffi._SIGS)
    assert declared <= bound, sorted(declared - bound)


def test_version_and_error_string():
    from exprgrad_b200 import _ffi
    assert b"sm_100a" in _ffi.lib.egb_version()
    assert _ffi.lib.egb_last_error() is not None


def test_no_device_is_reported_not_crashed():
    """cl.nim:95-99: newGpuContext raises GpuError("Unable to find device") when there is none."""
    import pytest
    import exprgrad_b200 as eg
    if len(eg.list_devices()) == 0:
        with pytest.raises(eg.GpuError):
            eg.new_gpu_context()


def test_contraction_planner_invariants():
    """egb_gemm_plan is pure host arithmetic: tile width and cluster split-K factor of the contraction kernel
    (csrc/gemm_tcgen05.cu) must always describe a launchable configuration."""
    import ctypes
    from exprgrad_b200._ffi import check, lib
    shapes = [(1024, 512, 784), (1024, 512, 512), (1024, 10, 512), (512, 10, 1024), (1024, 512, 10), (512, 512, 1024),
              (784, 512, 1024), (128, 4096, 4096), (1, 1, 1), (130, 70, 9), (300, 200, 129), (2048, 2048, 64), (64, 4000, 7000)]
    for (m, n, k) in shapes:
        for b_mn in (0, 1):
            for sms in (148, 72, 16):
                bn, ck = ctypes.c_int(0), ctypes.c_int(0)
                check(lib.egb_gemm_plan(m, n, k, b_mn, sms, ctypes.byref(bn), ctypes.byref(ck)))
                bn, ck = bn.value, ck.value
                assert 32 <= bn <= 256 and bn % (64 if b_mn else 32) == 0, (m, n, k, b_mn, sms, bn)
                assert ck in (1, 2, 4, 8)
                tiles = ((m + 127) // 128) * ((n + bn - 1) // bn)
                kb = (k + 63) // 64
                if ck > 1:
                    assert tiles * ck <= sms, "clusters of a split contraction must be co-resident in one wave"
                    assert (ck - 1) * ((kb + ck - 1) // ck) < kb, "every CTA of a cluster needs k-blocks"
                    assert 128 * (bn + 4) * 4 <= 192 * 1024, "the partial tile must fit the operand stages"
    # the dense-net adjoints: long reductions with few tiles are split, wide short ones are not
    bn, ck = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.egb_gemm_plan(512, 10, 1024, 1, 148, ctypes.byref(bn), ctypes.byref(ck)))
    assert ck.value >= 4
    check(lib.egb_gemm_plan(1024, 512, 10, 0, 148, ctypes.byref(bn), ctypes.byref(ck)))
    assert ck.value == 1


def test_latency_kernel_planner_invariants():
    """egb_gemm_lat_plan (csrc/gemm_lat.cu): tile width, cluster split-K factor and CTA count of the latency kernel
    must describe a launchable configuration; the dense-net shapes get the configurations the device timeline was
    measured with."""
    import ctypes
    from exprgrad_b200._ffi import check, lib
    shapes = [(1024, 512, 784), (1024, 512, 512), (784, 512, 1024), (512, 512, 1024), (128, 4096, 4096), (1, 4, 1), (130, 72, 9),
              (300, 200, 129), (37, 48, 64)]
    for (m, n, k) in shapes:
        for b_mn in (0, 1):
            for sms in (148, 72, 32, 16):
                bn, ck, ctas = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
                check(lib.egb_gemm_lat_plan(m, n, k, b_mn, sms, ctypes.byref(bn), ctypes.byref(ck), ctypes.byref(ctas)))
                bn, ck, ctas = bn.value, ck.value, ctas.value
                assert 32 <= bn <= 256 and bn % (64 if b_mn else 32) == 0, (m, n, k, b_mn, sms, bn)
                assert ck in (1, 2, 4)
                kb = (k + 63) // 64
                if ck > 1:
                    assert bn <= 64, "a split tile is pushed through distributed shared memory: at most 64 columns"
                    assert ctas <= sms, "clusters of a split contraction must be co-resident in one wave"
                    assert (ck - 1) * ((kb + ck - 1) // ck) < kb, "every CTA of a cluster needs k-blocks"
                assert 1 <= ctas <= max(sms, ((m + 127) // 128) * ((n + bn - 1) // bn) * ck)
    def plan(m, n, k, sms=148):
        bn, ck, ctas = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        check(lib.egb_gemm_lat_plan(m, n, k, 0, sms, ctypes.byref(bn), ctypes.byref(ck), ctypes.byref(ctas)))
        return bn.value, ck.value, ctas.value
    assert plan(1024, 512, 784) == (64, 2, 128)      # layer 1 forward
    assert plan(784, 512, 1024) == (64, 2, 112)      # layer 1 weight gradient
    assert plan(512, 512, 1024, sms=32) == (64, 1, 32)   # a side contraction inside the SMs the chain leaves free
