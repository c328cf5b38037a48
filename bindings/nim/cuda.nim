## runtimes/cuda.nim - exprgrad's backend-neutral device API (exprgrad/runtimes/gpu.nim:25-52) bound to
## libegb200.so (include/egb200.h). Selected by `-d:cuda` the way `-d:opencl` selects runtimes/cl.nim
## (exprgrad/runtimes/gpu.nim:20-22): same exported names, same argument meaning, same error type.
##
## Drop this file into exprgrad/runtimes/. It only declares `importc` procs and thin wrappers; there is no Nim
## toolchain in the build image of this repository, so it has been checked against the reference's signatures
## and against include/egb200.h (tests/test_bindings.py compares every imported symbol and its arity with the
## header) but not compiled there.
{.passL: "-legb200".}
type
  GpuError* = ref object of CatchableError          # cl.nim:18
  GpuDevice* = object
    index: cint
  EgbContext {.importc: "egb_context", header: "egb200.h", incompleteStruct.} = object
  EgbBuffer {.importc: "egb_buffer", header: "egb200.h", incompleteStruct.} = object
  EgbKernel {.importc: "egb_kernel", header: "egb200.h", incompleteStruct.} = object
  GpuContext* = ref object
    handle*: ptr EgbContext
  GpuBuffer* = object                                 # cl.nim:27-30
    ctx: GpuContext
    size: int
    handle: ptr EgbBuffer
  GpuKernelSource* = object                           # cl.nim:33-35: `source` = compiled-program text
    name*, source*: string
  GpuKernel* = ref object
    handle: ptr EgbKernel

proc egb_last_error(): cstring {.importc, cdecl.}
proc egb_device_count(count: ptr cint): cint {.importc, cdecl.}
proc egb_device_name(device: cint, buf: cstring, cap: csize_t): cint {.importc, cdecl.}
proc egb_device_vendor(device: cint, buf: cstring, cap: csize_t): cint {.importc, cdecl.}
proc egb_device_version(device: cint, buf: cstring, cap: csize_t): cint {.importc, cdecl.}
proc egb_device_is_gpu(device: cint, isGpu: ptr cint): cint {.importc, cdecl.}
proc egb_context_create(device: cint, res: ptr ptr EgbContext): cint {.importc, cdecl.}
proc egb_alloc_buffer(ctx: ptr EgbContext, bytes: csize_t, res: ptr ptr EgbBuffer): cint {.importc, cdecl.}
proc egb_buffer_free(buf: ptr EgbBuffer): cint {.importc, cdecl.}
proc egb_buffer_write(buf: ptr EgbBuffer, data: pointer, bytes: csize_t): cint {.importc, cdecl.}
proc egb_buffer_fill(buf: ptr EgbBuffer, value: pointer, elemSize: csize_t): cint {.importc, cdecl.}
proc egb_buffer_read_into(buf: ptr EgbBuffer, data: pointer, bytes: csize_t): cint {.importc, cdecl.}
proc egb_compile(ctx: ptr EgbContext, name, source: cstring, res: ptr ptr EgbKernel): cint {.importc, cdecl.}
proc egb_kernel_arg_buffer(k: ptr EgbKernel, index: cint, buf: ptr EgbBuffer): cint {.importc, cdecl.}
proc egb_kernel_arg_shape(k: ptr EgbKernel, index, rank: cint, dims: ptr int64): cint {.importc, cdecl.}
proc egb_kernel_arg_index(k: ptr EgbKernel, index: cint, value: int64): cint {.importc, cdecl.}
proc egb_kernel_run(k: ptr EgbKernel, dims: cint, group, local: ptr int64): cint {.importc, cdecl.}

template check(status: cint) =                        # cl.nim:41-43
  if status != 0: raise GpuError(msg: $egb_last_error())

proc listDevices*(): seq[GpuDevice] =                 # cl.nim:63-65
  var n: cint
  check egb_device_count(n.addr)
  for it in 0..<n: result.add(GpuDevice(index: it))
proc name*(device: GpuDevice): string =               # cl.nim:74
  result = newString(256); check egb_device_name(device.index, result.cstring, 256)
  result.setLen(result.cstring.len)
proc queryString(device: GpuDevice, query: proc (device: cint, buf: cstring, cap: csize_t): cint {.cdecl.}): string =
  result = newString(256); check query(device.index, result.cstring, 256)
  result.setLen(result.cstring.len)
proc vendor*(device: GpuDevice): string = device.queryString(egb_device_vendor)    # cl.nim:76
proc version*(device: GpuDevice): string = device.queryString(egb_device_version)  # cl.nim:77
proc isGpu*(device: GpuDevice): bool =                # cl.nim:78-81
  var flag: cint
  check egb_device_is_gpu(device.index, flag.addr)
  result = flag != 0
proc newGpuContext*(device: GpuDevice): GpuContext =  # cl.nim:83-93
  result = GpuContext(); check egb_context_create(device.index, result.handle.addr)
proc newGpuContext*(): GpuContext =                   # cl.nim:95-99 ("Unable to find device" comes from the library)
  result = GpuContext(); check egb_context_create(-1, result.handle.addr)
proc allocBuffer*(ctx: GpuContext, size: int): GpuBuffer =       # cl.nim:101-106
  result = GpuBuffer(ctx: ctx, size: size); check egb_alloc_buffer(ctx.handle, csize_t(size), result.handle.addr)
proc dealloc*(buffer: GpuBuffer) = check egb_buffer_free(buffer.handle)   # cl.nim:108-109
proc write*(buffer: GpuBuffer, data: pointer, size: int) =       # cl.nim:111-116 (size check is in the library)
  check egb_buffer_write(buffer.handle, data, csize_t(size))
proc write*[T](buffer: GpuBuffer, data: openArray[T]) =
  buffer.write(data[0].unsafeAddr, data.len * sizeof(T))
proc fill*[T](buffer: GpuBuffer, value: T) =                     # cl.nim:122-126
  var v = value; check egb_buffer_fill(buffer.handle, v.addr, csize_t(sizeof(T)))
proc readInto*[T](buffer: GpuBuffer, data: ptr UncheckedArray[T]) =   # cl.nim:128-131
  check egb_buffer_read_into(buffer.handle, data, csize_t(buffer.size))
proc readInto*[T](buffer: GpuBuffer, data: var seq[T]) =              # cl.nim:133-138
  if data.len * sizeof(T) != buffer.size:
    raise GpuError(msg: "Buffer size is not equal to target size")
  if data.len > 0:
    buffer.readInto(cast[ptr UncheckedArray[T]](data[0].addr))
proc read*[T](buffer: GpuBuffer): seq[T] =                            # cl.nim:140-147
  if buffer.size mod sizeof(T) != 0:
    raise GpuError(msg: "Buffer size is not divisible by item size")
  result = newSeq[T](buffer.size div sizeof(T))
  buffer.readInto(result)
proc compile*(ctx: GpuContext, name, source: string): GpuKernel =     # cl.nim:149-176
  result = GpuKernel(); check egb_compile(ctx.handle, name.cstring, source.cstring, result.handle.addr)
proc compile*(ctx: GpuContext, source: GpuKernelSource): GpuKernel = ctx.compile(source.name, source.source)
proc arg*(kernel: GpuKernel, index: int, buffer: GpuBuffer): GpuKernel =  # cl.nim:186-188
  result = kernel; check egb_kernel_arg_buffer(kernel.handle, cint(index), buffer.handle)
proc arg*[T](kernel: GpuKernel, index: int, value: T): GpuKernel =        # cl.nim:181-184
  result = kernel; check egb_kernel_arg_index(kernel.handle, cint(index), int64(value))
proc run*(kernel: GpuKernel, groupSize, localSize: openArray[int]) =      # cl.nim:190-207
  var g = @groupSize; var l = @localSize
  check egb_kernel_run(kernel.handle, cint(g.len), cast[ptr int64](g[0].addr), cast[ptr int64](l[0].addr))
