"""ctypes binding of libegb200.so (include/egb200.h). The library is the product; this file only
declares signatures and converts status codes into the reference's exception types
(exprgrad/ir.nim:18-28, exprgrad/runtimes/cl.nim:18, 41-43). There is no fallback: if the shared
library is missing the import fails loudly."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libegb200.so")


class GpuError(Exception): pass
class RuntimeError_(Exception): pass
class ShapeError(Exception): pass
class ParserError(Exception): pass
class GradientError(Exception): pass
class GeneratorError(Exception): pass
class ValueError_(Exception): pass

_ERRORS = {1: GpuError, 2: RuntimeError_, 3: ShapeError, 4: ParserError, 5: GradientError, 6: GeneratorError,
           7: ValueError_}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(exprgrad_b200 has no CPU fallback)")

lib = ctypes.CDLL(LIB_PATH)

P = ctypes.c_void_p
I = ctypes.c_int
I64 = ctypes.c_int64
SZ = ctypes.c_size_t
F = ctypes.c_float
D = ctypes.c_double
S = ctypes.c_char_p
PP = ctypes.POINTER(P)
PI = ctypes.POINTER(I)
PI64 = ctypes.POINTER(I64)
U64 = ctypes.c_uint64

_SIGS = {
    "egb_last_error": (S, []),
    "egb_version": (S, []),
    "egb_device_count": (I, [PI]),
    "egb_device_name": (I, [I, S, SZ]),
    "egb_device_vendor": (I, [I, S, SZ]),
    "egb_device_version": (I, [I, S, SZ]),
    "egb_device_is_gpu": (I, [I, PI]),
    "egb_context_create": (I, [I, PP]),
    "egb_context_destroy": (I, [P]),
    "egb_context_synchronize": (I, [P]),
    "egb_context_stream": (P, [P]),
    "egb_context_launch_count": (I64, [P]),
    "egb_context_set_option": (I, [P, S, I64]),
    "egb_context_set_timing": (I, [P, I]),
    "egb_context_kernel_time": (I, [P, I, ctypes.POINTER(D), PI64]),
    "egb_event_create": (I, [P, PP]),
    "egb_event_record": (I, [P, P]),
    "egb_event_elapsed_ms": (I, [P, P, ctypes.POINTER(D)]),
    "egb_event_destroy": (I, [P]),
    "egb_host_alloc": (I, [SZ, PP]),
    "egb_host_free": (I, [P]),
    "egb_alloc_buffer": (I, [P, SZ, PP]),
    "egb_buffer_free": (I, [P]),
    "egb_buffer_size": (SZ, [P]),
    "egb_buffer_device_ptr": (P, [P]),
    "egb_buffer_write": (I, [P, P, SZ]),
    "egb_buffer_fill": (I, [P, P, SZ]),
    "egb_buffer_read_into": (I, [P, P, SZ]),
    "egb_compile": (I, [P, S, S, PP]),
    "egb_kernel_free": (I, [P]),
    "egb_kernel_arg_count": (I, [P, PI]),
    "egb_kernel_arg_buffer": (I, [P, I, P]),
    "egb_kernel_arg_shape": (I, [P, I, I, PI64]),
    "egb_kernel_arg_index": (I, [P, I, I64]),
    "egb_kernel_run": (I, [P, I, PI64, PI64]),
    "egb_gemm_f32": (I, [P, I, I, I64, I64, I64, P, I64, P, I64, P, I64, I, P, F]),
    "egb_gemm_plan": (I, [I64, I64, I64, I, I, PI, PI]),
    "egb_gemm_lat_plan": (I, [I64, I64, I64, I, I, PI, PI, PI]),
    "egb_gemm_planes": (I, [P, I64, I64, I64, P, P, I64, P, P, I64, P, I64, I, P, F, I]),
    "egb_split_bf16": (I, [P, P, I64, I64, I64, I, P, P, I64, I]),
    "egb_program_parse": (I, [S, SZ, PP]),
    "egb_program_compile": (I, [P]),
    "egb_program_serialize": (I, [P, P, SZ, ctypes.POINTER(SZ)]),
    "egb_program_describe": (I, [P, S, P, SZ, ctypes.POINTER(SZ)]),
    "egb_program_classify": (I, [P, S, I, ctypes.POINTER(S), PI, PI64, P, SZ, ctypes.POINTER(SZ)]),
    "egb_program_lower_dump": (I, [P, S, I, ctypes.POINTER(S), PI, PI64, I, I64, P, SZ, ctypes.POINTER(SZ)]),
    "egb_program_free": (I, [P]),
    "egb_program_tensor_count": (I, [P, PI]),
    "egb_program_tensor_info": (I, [P, I, PI, PI, PI64, P, SZ]),
    "egb_program_target_output": (I, [P, S, PI]),
    "egb_program_infer_shapes": (I, [P, S, I, ctypes.POINTER(S), PI, PI64, I, PI, PI64]),
    "egb_model_create": (I, [P, P, U64, PP]),
    "egb_model_free": (I, [P]),
    "egb_model_set_option": (I, [P, S, I64]),
    "egb_model_epoch": (I, [P, PI64]),
    "egb_model_write_tensor": (I, [P, I, P, SZ]),
    "egb_model_read_tensor": (I, [P, I, P, SZ]),
    "egb_model_tensor_shape": (I, [P, I, PI, PI64]),
    "egb_model_tensor_device_ptr": (I, [P, I, PP]),
    "egb_model_call": (I, [P, S, I, ctypes.POINTER(S), PP, PI, PI64, PI, PI, PI64]),
    "egb_model_call_read": (I, [P, S, I, ctypes.POINTER(S), PP, PI, PI64, PI, P, SZ, PI, PI64]),
    "egb_model_read_output": (I, [P, P, SZ]),
    "egb_model_fit": (I, [P, S, I, ctypes.POINTER(S), PP, PI, PI64, I64, PI64]),
    "egb_model_describe_plan": (I, [P, P, SZ, ctypes.POINTER(SZ)]),
    "egb_model_plan_count": (I, [P, PI]),
    "egb_comm_unique_id": (I, [P, SZ]),
    "egb_comm_create": (I, [P, P, I, I, PP]),
    "egb_comm_destroy": (I, [P]),
    "egb_comm_info": (I, [P, PI, PI, PI]),
    "egb_comm_allreduce_avg_f32": (I, [P, P, SZ]),
    "egb_model_set_data_parallel": (I, [P, P]),
}


def declare(sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args


declare(_SIGS)


def check(status: int):
    if status != 0:
        msg = lib.egb_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(status, GpuError)(msg)
