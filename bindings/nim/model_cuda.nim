## model_cuda.nim - the `when defined(cuda)` branch of exprgrad/model.nim for CompileGpu targets: `newModel`
## hands the compiled program to libegb200.so once (egb_program_parse + egb_model_create), `call/apply/fit`
## (exprgrad/model.nim:392-454) become one C-ABI call each, and `flushStateTensors` (model.nim:326-345) moves
## model.params / model.caches through egb_model_read_tensor / egb_model_write_tensor.
## `include` this file from model.nim next to the existing OpenCL branch; it uses model.nim's own types
## (Model[T], ModelObj.gpuHandle is the one field to add) and runtimes/cuda.nim. Not compiled here (no Nim).
import std/[tables, sequtils]
import ir, tensors, runtimes/gpu, io/egbtext

when defined(cuda):
  proc egb_program_parse(text: cstring, len: csize_t, res: ptr pointer): cint {.importc, cdecl.}
  proc egb_model_create(ctx, program: pointer, seed: uint64, res: ptr pointer): cint {.importc, cdecl.}
  proc egb_model_write_tensor(model: pointer, tensor: cint, host: pointer, bytes: csize_t): cint {.importc, cdecl.}
  proc egb_model_read_tensor(model: pointer, tensor: cint, host: pointer, bytes: csize_t): cint {.importc, cdecl.}
  proc egb_model_call(model: pointer, target: cstring, nArgs: cint, names: cstringArray, data: ptr pointer,
                      ranks: ptr cint, dims: ptr int64, onDevice: ptr cint,
                      outRank: ptr cint, outDims: ptr int64): cint {.importc, cdecl.}
  proc egb_model_read_output(model: pointer, dst: pointer, bytes: csize_t): cint {.importc, cdecl.}
  proc egb_model_fit(model: pointer, target: cstring, nArgs: cint, names: cstringArray, data: ptr pointer,
                     ranks: ptr cint, dims: ptr int64, batchSize: int64, batches: ptr int64): cint {.importc, cdecl.}

  # newModel (model.nim:232-251): the program is handed over once; parameters keep the host copy the
  # reference initialises with newRandTensor so that `model.params` stays meaningful.
  proc newGpuModel[T](program: Program, ctx: GpuContext, params: Table[TensorId, Tensor[T]]): pointer =
    let text = program.toEgbText()                 # io/egbtext.nim, stage 1 (already compiled by passes.nim)
    var prog: pointer
    check egb_program_parse(text.cstring, csize_t(text.len), prog.addr)
    check egb_model_create(ctx.handle, prog, 0'u64, result.addr)
    for id, tensor in params:                      # flushStateTensors(to = CompileGpu), model.nim:337-344
      check egb_model_write_tensor(result, cint(id), tensor.dataPtr, csize_t(tensor.len * sizeof(T)))

  # call (model.nim:392-406), CompileGpu case
  proc callGpu[T](model: Model[T], targetName: string, args: openArray[(string, Tensor[T])]): Tensor[T] =
    var names = allocCStringArray(args.mapIt(it[0])); defer: deallocCStringArray(names)
    var data: seq[pointer]; var ranks: seq[cint]; var dims: seq[int64]
    for (name, tensor) in args:
      data.add(tensor.dataPtr); ranks.add(cint(tensor.shape.len))
      for d in tensor.shape: dims.add(int64(d))
    var outRank: cint; var outDims: array[8, int64]
    let status = egb_model_call(model.gpuHandle, targetName.cstring, cint(args.len), names, data[0].addr,
                                ranks[0].addr, dims[0].addr, nil, outRank.addr, outDims[0].addr)
    if status == 2: raise RuntimeError(msg: $egb_last_error())      # model.nim:358-359, 395-396
    if status == 3: raise ShapeError(msg: $egb_last_error())        # passes.nim:1393-1403
    check status
    if outRank >= 0:
      result = newTensor[T](outDims[0..<outRank].mapIt(int(it)))    # readOutput, model.nim:370-376
      check egb_model_read_output(model.gpuHandle, result.dataPtr, csize_t(result.len * sizeof(T)))

  # One-shot form of call: run the target and fill a host tensor in one blocking call (streams row blocks of a
  # single-contraction target - benchmarks/matmul/matmul_gpu.nim:28-75 - so that copies and tensor cores overlap).
  proc egb_model_call_read(model: pointer, target: cstring, nArgs: cint, names: cstringArray, data: ptr pointer,
                           ranks: ptr cint, dims: ptr int64, onDevice: ptr cint, outHost: pointer, outBytes: csize_t,
                           outRank: ptr cint, outDims: ptr int64): cint {.importc, cdecl.}

  # fit (model.nim:413-454): one call; the library uploads the data set once and slices batches on the device.
  proc fitGpu[T](model: Model[T], targetName: string, args: openArray[(string, Tensor[T])], batchSize: int) =
    var names = allocCStringArray(args.mapIt(it[0])); defer: deallocCStringArray(names)
    var data: seq[pointer]; var ranks: seq[cint]; var dims: seq[int64]
    for (name, tensor) in args:
      data.add(tensor.dataPtr); ranks.add(cint(tensor.shape.len))
      for d in tensor.shape: dims.add(int64(d))
    var batches: int64
    let status = egb_model_fit(model.gpuHandle, targetName.cstring, cint(args.len), names, data[0].addr,
                               ranks[0].addr, dims[0].addr, int64(batchSize), batches.addr)
    if status == 2: raise RuntimeError(msg: $egb_last_error())
    if status == 3: raise ShapeError(msg: $egb_last_error())
    check status
    model.epoch += 1

  # flushStateTensors(to = CompileCpu) (model.nim:330-336): device state back into model.params / model.caches
  proc readStateGpu[T](model: Model[T]) =
    for id, tensor in model.params:
      check egb_model_read_tensor(model.gpuHandle, cint(id), tensor.dataPtr, csize_t(tensor.len * sizeof(T)))
    for id, tensor in model.caches:
      check egb_model_read_tensor(model.gpuHandle, cint(id), tensor.dataPtr, csize_t(tensor.len * sizeof(T)))
