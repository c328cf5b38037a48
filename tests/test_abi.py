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
