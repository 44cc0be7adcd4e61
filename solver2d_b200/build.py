"""Build recipe for the product library ``solver2d_b200/libsolver2d.so`` (host C + CUDA for sm_100a).

* host C   : gcc -std=gnu17 -O2 (the reference's own flags: no -march, no -ffast-math) for csrc/host/*.c
* device   : nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false for csrc/device/*.cu
             (-fmad=false keeps every float op un-contracted so the kernels evaluate the same IEEE operations as the
             reference CPU solver; the path is bandwidth/latency bound, the extra FMUL/FADD issue slots are free)
* link     : nvcc -shared, static cudart, in-tree so the .so travels to the GPU box with the snapshot

Run ``python -m solver2d_b200.build`` or call :func:`build`.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsolver2d.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I" + INCLUDE, "-I" + os.path.join(HERE, "csrc", "device"), "-I" + os.path.join(HERE, "csrc"),
]
C_FLAGS = ["-std=gnu17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", "-fvisibility=hidden",
           "-I" + INCLUDE, "-I" + os.path.join(HERE, "csrc", "host"), "-I" + os.path.join(HERE, "csrc")]


def _deps_hash(src: str, flags: list[str]) -> str:
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    # conservative: any header change rebuilds everything
    files = [src] + sorted(glob.glob(os.path.join(INCLUDE, "**", "*.h"), recursive=True)) \
        + sorted(glob.glob(os.path.join(HERE, "csrc", "**", "*.cuh"), recursive=True)) \
        + sorted(glob.glob(os.path.join(HERE, "csrc", "**", "*.h"), recursive=True))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    is_cu = src.endswith(".cu")
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    flags = NVCC_FLAGS if is_cu else C_FLAGS
    stamp = obj + ".sha1"
    digest = _deps_hash(src, flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = ([NVCC] + flags + ["-c", src, "-o", obj]) if is_cu else ([CC] + flags + ["-c", src, "-o", obj])
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"compile failed: {src}")
    if verbose and res.stderr.strip():
        sys.stderr.write(res.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return obj


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ, "*")):
            os.remove(f)
    srcs = sorted(glob.glob(os.path.join(HERE, "csrc", "device", "*.cu"))) + \
        sorted(glob.glob(os.path.join(HERE, "csrc", "host", "*.c")))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                                                     "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined",
                                                     "-Xlinker", "-Bsymbolic", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
