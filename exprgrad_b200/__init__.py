"""exprgrad_b200: B200 (sm_100a) execution backend for exprgrad's compiled hot path.

The product is libegb200.so (CUDA kernels + host runtime behind the C ABI of include/egb200.h);
this package is the thin host-side mirror of the reference's own interfaces on top of it.
Importing it requires the built library - there is no CPU fallback."""
from ._ffi import (GpuError, RuntimeError_, ShapeError, ParserError, GradientError, GeneratorError, ValueError_,
                   LIB_PATH)
from .gpu import (GpuBuffer, GpuContext, GpuDevice, GpuTensor, alloc_tensor, list_devices, new_gpu_context)
from .model import Model, Program, compile, load_model
from . import frontend, layers, dist
