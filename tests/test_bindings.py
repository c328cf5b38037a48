"""CPU-only checks of the boundary artifacts a maintainer of the reference would use:
  * include/egb200.h compiles as strict C99 and a C program links against libegb200.so and runs;
  * bindings/nim/*.nim import only symbols the header declares, with the header's parameter counts
    (there is no Nim compiler in this image, so the bindings are checked textually)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    """name -> number of parameters, from include/egb200.h"""
    text = open(os.path.join(ROOT, "include", "egb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(egb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def nim_imports(path):
    """(name, number of parameters) of every `proc egb_...(...) ... {.importc ...}` declaration"""
    text = open(path).read()
    out = []
    for m in re.finditer(r"proc\s+(egb_[a-z0-9_]+)\s*\((.*?)\)\s*(?::\s*[\w ]+)?\s*\{\.importc", text, flags=re.S):
        # "a, b: T, c: ptr U" declares one parameter per comma-separated piece
        n = sum(1 for piece in m.group(2).split(",") if piece.strip())
        out.append((m.group(1), n))
    return out


def test_header_compiles_as_c99_and_links(tmp_path):
    exe = tmp_path / "abi_c99"
    lib_dir = os.path.join(ROOT, "exprgrad_b200")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_c99.c"), "-o", str(exe), "-L", lib_dir, "-l:libegb200.so",
           "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok ")


def test_nim_bindings_match_the_header():
    funcs = header_functions()
    assert len(funcs) >= 50
    nim_dir = os.path.join(ROOT, "bindings", "nim")
    seen = set()
    for fn in ("cuda.nim", "model_cuda.nim", "egbtext.nim"):
        path = os.path.join(nim_dir, fn)
        assert os.path.exists(path), path
        for name, nparams in nim_imports(path):
            assert name in funcs, f"{fn} imports {name}, which include/egb200.h does not declare"
            assert funcs[name] == nparams, f"{fn}: {name} has {nparams} parameters, the header declares {funcs[name]}"
            seen.add(name)
    # the device API of runtimes/gpu.nim:25-52 and the model entry points must all be bound
    for needed in ("egb_device_count", "egb_context_create", "egb_alloc_buffer", "egb_buffer_write", "egb_buffer_fill",
                   "egb_buffer_read_into", "egb_compile", "egb_kernel_arg_buffer", "egb_kernel_run", "egb_program_parse",
                   "egb_model_create", "egb_model_call", "egb_model_call_read", "egb_model_fit", "egb_model_read_tensor",
                   "egb_model_write_tensor", "egb_model_read_output"):
        assert needed in seen, needed


def test_nim_text_writer_and_python_writer_share_the_grammar():
    """io/egbtext.nim and exprgrad_b200/frontend.py `serialize` must emit the same record tags in the same
    order; the C++ parser (csrc/program.cpp) is the arbiter for the Python one (tests/test_passes_parity.py)."""
    nim = open(os.path.join(ROOT, "bindings", "nim", "egbtext.nim")).read()
    tags = re.findall(r'res(?:ult)?\.tok\("([A-Za-z]+)"\)', nim) + re.findall(r'\.emit\(\w+(?:\.\w+)?, "([RW])"\)', nim)
    for tag in ("egbprog", "tensors", "T", "targets", "target", "S", "copy", "dims", "rank", "linear", "K", "L", "R",
                "I", "W", "C", "LI", "end"):
        assert tag in tags, tag
    assert tags.index("egbprog") < tags.index("tensors") < tags.index("targets") < tags.index("end")
