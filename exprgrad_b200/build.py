"""In-tree build of libegb200.so (sm_100a only).

Every translation unit under csrc/ is compiled with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` and linked into
exprgrad_b200/libegb200.so, next to this file, so that the library travels with the source tree.
Objects are cached under csrc/_obj keyed on the source + header contents.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libegb200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "--expt-relaxed-constexpr", "-I", os.path.join(os.path.dirname(HERE), "include")]
# Kernels that restate the reference's strict fp32 arithmetic (separate fmul + fadd, llvmgen.nim:219-221)
# must not be contracted into FMAs.
STRICT = {"interp.cu", "eltwise.cu", "eltwise_stream.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers_digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".hpp", ".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(os.path.dirname(HERE), "include", "egb200.h"), "rb").read())
    return h.hexdigest()


def _flags(src):
    fl = ARCH + COMMON
    if src in STRICT:
        fl = fl + ["-fmad=false"]
    if src.endswith(".cpp"):
        fl = fl + ["-x", "cu"]
    return fl


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()
    jobs, objs = [], []
    for src in _sources():
        path = os.path.join(CSRC, src)
        key = hashlib.sha256(open(path, "rb").read() + hd.encode() + " ".join(_flags(src)).encode()).hexdigest()[:16]
        obj = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{key}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            for old in os.listdir(OBJ):
                if old.startswith(os.path.splitext(src)[0] + "."):
                    os.remove(os.path.join(OBJ, old))
            jobs.append([NVCC] + _flags(src) + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libegb200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
