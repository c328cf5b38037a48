## io/egbtext.nim - prints an exprgrad `Program` (exprgrad/ir.nim:263-270) in the text form that
## `egb_program_parse` (include/egb200.h, csrc/program.cpp) reads: a field-wise walk over the same types
## io/serialize.nim:323-342 stores in binary. Used by the `-d:cuda` branch of model.nim (model_cuda.nim) to hand
## a compiled program to libegb200.so once per `compile[T]`.
##
## Grammar (whitespace-separated tokens): INTEGRATION.md section 4. The writer below and
## exprgrad_b200/frontend.py `serialize` (the Python stand-in for this file, exercised by every test of this
## repository) emit token-identical text for the same program. Not compiled in this repository's image (no Nim).
import std/[tables, sets, strutils, algorithm]
import ../ir

proc tok(res: var string, value: string) = res.add(value); res.add(' ')
proc tok(res: var string, value: int) = res.add($value); res.add(' ')
proc tok(res: var string, value: bool) = res.add(if value: "1 " else: "0 ")

proc tokFloat(res: var string, value: float64) =
  ## C99 hex float (what strtod reads back exactly); inf / nan by name
  if value != value: res.tok("nan")
  elif value == Inf: res.tok("inf")
  elif value == NegInf: res.tok("-inf")
  else:
    var buf: array[64, char]
    proc snprintf(buf: cstring, cap: csize_t, fmt: cstring): cint {.importc, header: "<stdio.h>", varargs.}
    let len = snprintf(cast[cstring](buf[0].addr), 64, "%a", value)
    var text = newString(len)
    for it in 0..<len: text[it] = buf[it]
    res.tok(text)

proc tokStr(res: var string, value: string) =
  ## percent-encoded, "-" = empty (program.cpp Reader::str)
  if value.len == 0:
    res.tok("-")
    return
  var text = ""
  for chr in value:
    if chr in {'a'..'z', 'A'..'Z', '0'..'9', '_', '.'}: text.add(chr)
    else: text.add("%" & toHex(ord(chr), 2))
  res.tok(text)

proc opName(kind: InstrKind): string = ($kind)[len("Instr")..^1]   # InstrAdd -> Add (program.cpp op_from_name)

proc emit(res: var string, instr: Instr) =
  res.tok("I"); res.tok(instr.kind.opName)
  res.tok(int(instr.res)); res.tok(int(instr.tensor))
  res.tok(if instr.kind == InstrShape: instr.dim else: 0)
  res.tok(instr.args.len)
  for arg in instr.args: res.tok(int(arg))
  res.tokFloat(if instr.kind == InstrScalar: instr.scalarLit else: 0.0)
  case instr.kind:
    of InstrIndex: res.tok(instr.indexLit)
    of InstrBoolean: res.tok(ord(instr.booleanLit))
    else: res.tok(0)

proc emit(res: var string, index: LinearIndex) =
  res.tok("LI"); res.tok(index.setup.len); res.tok(index.factors.len); res.tok(index.constant)
  for instr in index.setup: res.emit(instr)
  var regs: seq[int] = @[]
  for reg in index.factors.keys: regs.add(int(reg))
  regs.sort()                                        # deterministic order (the parser keeps a sorted map)
  for reg in regs:
    res.tok(reg); res.tok(index.factors[RegId(reg)])

proc emit(res: var string, op: TensorOp, tag: string) =
  res.tok(tag); res.tok(int(op.tensor)); res.tok(op.isRaw); res.tok(int(op.data)); res.tok(op.dims.len)
  for dim in op.dims: res.emit(dim)

proc emit(res: var string, kernel: Kernel) =
  res.tok("K"); res.tok(ord(kernel.generator.kind)); res.tok(int(kernel.generator.tensor))
  if kernel.generator.kind == GenReshape:
    res.tok(kernel.generator.reshape.len)
    for size in kernel.generator.reshape: res.tok(size)
  else:
    res.tok(0)
  res.tok(kernel.regs.len); res.tok(kernel.loops.len); res.tok(kernel.reads.len)
  res.tok(kernel.expr.instrs.len); res.tok(int(kernel.expr.res)); res.tok(kernel.grad.isCustom)
  res.add('\n')
  for loop in kernel.loops:
    res.tok("L"); res.tok(int(loop.iter)); res.tok(loop.hasBounds); res.tok(loop.step)
    res.tok(if loop.mode >= LoopIndependent: 1 else: 0)
    res.emit(loop.start); res.emit(loop.stop); res.add('\n')
  for read in kernel.reads:
    res.emit(read, "R"); res.add('\n')
  for instr in kernel.expr.instrs: res.emit(instr)
  res.add('\n')
  res.emit(kernel.write, "W"); res.add('\n')
  if kernel.grad.isCustom:
    proc sortedPairs(tab: Table[TensorId, TensorId]): seq[(int, int)] =
      for a, b in tab: result.add((int(a), int(b)))
      result.sort()
    res.tok("C")
    let tensors = kernel.grad.tensors.sortedPairs()
    res.tok(tensors.len)
    for (a, b) in tensors: res.tok(a); res.tok(b)
    let subs = kernel.grad.subs.sortedPairs()
    res.tok(subs.len)
    for (a, b) in subs: res.tok(a); res.tok(b)
    res.tok(kernel.grad.kernels.len); res.add('\n')
    for child in kernel.grad.kernels: res.emit(child)

proc emit(res: var string, constr: ShapeConstraint) =
  res.tok("S")
  case constr.kind:
    of ShapeCopy:
      res.tok("copy"); res.tok(int(constr.dest)); res.tok(ord(constr.priority)); res.tok(int(constr.src))
    of ShapeDims:
      res.tok("dims"); res.tok(int(constr.dest)); res.tok(ord(constr.priority)); res.tok(constr.dims.len)
      for dim in constr.dims: res.emit(dim)
    of ShapeRank:
      res.tok("rank"); res.tok(int(constr.dest)); res.tok(ord(constr.priority)); res.tok(constr.rank)
    of ShapeLinear:
      res.tok("linear"); res.tok(int(constr.dest)); res.tok(ord(constr.priority)); res.tok(constr.reads.len)
      for tensor, dims in constr.reads:
        res.tok(int(tensor)); res.tok(dims.len)
        for dim in dims:
          res.tok(dim.len)
          for index in dim: res.emit(index)
      res.tok(constr.write.len)
      for index in constr.write: res.emit(index)
    of ShapeNone:
      raise newException(ValueError, "ShapeNone constraint cannot be serialised")
  res.add('\n')

proc toEgbText*(program: Program): string =
  ## `compiled` (stage 1) = the program went through passes.nim up to sortShapeConstraints (model.nim:46-58);
  ## a source program (stage 0) is compiled by the library's own restatement of those passes.
  let compiled = StageSortedShapes in program.stages
  result.tok("egbprog"); result.tok(1)
  result.tok(if program.scalarType == Scalar32: "f32" else: "f64")
  result.tok(compiled); result.add('\n')
  result.tok("tensors"); result.tok(program.tensors.len); result.add('\n')
  for def in program.tensors:
    result.tok("T"); result.tok(ord(def.kind)); result.tok(def.shape.len)
    for size in def.shape: result.tok(size)
    case def.kind:
      of TensorParam:
        result.tokFloat(def.initRange.a); result.tokFloat(def.initRange.b); result.tok(0)
      of TensorRandom:
        result.tokFloat(def.randomRange.a); result.tokFloat(def.randomRange.b); result.tok(0)
      of TensorCache:
        result.tokFloat(0.0); result.tokFloat(0.0); result.tok(int(def.cache))
      else:
        result.tokFloat(0.0); result.tokFloat(0.0); result.tok(0)
    result.tokStr(def.name); result.add('\n')
  result.tok("targets"); result.tok(program.targets.len); result.add('\n')
  for name, target in program.targets:
    result.tok("target"); result.tokStr(name); result.tok(int(target.output)); result.tok(ord(target.compileTarget))
    result.tok(target.shapes.len); result.tok(target.kernels.len); result.tok(target.tensors.len)
    var ids: seq[int] = @[]
    for tensor in target.tensors: ids.add(int(tensor))
    ids.sort()
    for id in ids: result.tok(id)
    result.add('\n')
    for constr in target.shapes: result.emit(constr)
    for kernel in target.kernels: result.emit(kernel)
  result.tok("end"); result.add('\n')
