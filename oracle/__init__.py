"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's compiled tensor-execution path (exprgrad, Nim + LLVM-JIT), used
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
and baseline. The product package (exprgrad_b200/) never imports anything from here.

Parity pinning: the reference itself cannot be built in this image (no Nim compiler, no LLVM 13), so
the oracle is pinned against the reference's own known-answer tests (tests/test_model.nim,
test_talks.nim, test_dnn.nim, test_errors.nim, test_tensors.nim) re-stated in
tests/test_reference_vectors.py.
"""
from .ir import *  # noqa: F401,F403
from .ir import input, param, cond, cache, rand  # noqa: F401
from .model import Model, compile  # noqa: F401
from . import layers  # noqa: F401
