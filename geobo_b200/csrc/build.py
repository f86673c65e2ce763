#!/usr/bin/env python
"""Build libgeobo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python geobo_b200/csrc/build.py [--force] [--verbose]

Objects are compiled in parallel and linked with `nvcc -shared`; the only link-time
dependency is the CUDA runtime (libnccl is dlopen'ed lazily by comm.cu).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["api.cu", "gemm_f64.cu", "cov.cu", "sens.cu", "chol.cu", "comm.cu", "ozaki.cu", "ozaki_gemm.cu", "refine.cu", "acq.cu", "drill.cu", "kron.cu", "stencil.cu", "fftconv.cu"]
HEADERS = ["common.cuh", "comm.h", "umma.cuh", "ozaki.cuh", "drill.cuh", "kron.cuh", "stencil.cuh", "fftconv.cuh", "formulas.cuh", os.path.join("..", "..", "include", "geobo_b200.h")]
LIB = os.path.join(PKG, "libgeobo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    if not force and _newer(LIB, srcs + hdrs + [os.path.abspath(__file__)]):
        return LIB

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if not force and _newer(obj, [src] + hdrs):
            return obj, ""
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(compile_one, srcs))
    if verbose:
        for obj, log in results:
            print(log)
    objs = [o for o, _ in results]
    r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
