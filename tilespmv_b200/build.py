"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

  libtilespmv_b200.so  -- CUDA kernels + the C-ABI of include/tilespmv.h (nvcc, sm_100a only)
  libtilespmv_gen.so   -- host-side synthetic matrix generators (gcc + OpenMP)
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_CUDA = os.path.join(PKG, "libtilespmv_b200.so")
LIB_GEN = os.path.join(PKG, "libtilespmv_gen.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GCC = "/usr/bin/gcc"  # the image's $CC (/opt/gcc) has no libgomp.spec
GPP = "/usr/bin/g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false" if False else "-fmad=true",  # plain IEEE: no fast-math anywhere
    "-Xcompiler", "-fPIC,-fopenmp,-O3", "-shared", "-ccbin", GPP,
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def cuda_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "tilespmv.h"))
    return deps


def build_gen(force=False, verbose=False):
    src = os.path.join(CSRC, "gen.c")
    if force or _stale(LIB_GEN, [src]):
        cmd = [GCC, "-O3", "-fopenmp", "-fPIC", "-shared", src, "-o", LIB_GEN]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB_GEN


def build_cuda(force=False, verbose=False, extra=()):
    if force or _stale(LIB_CUDA, cuda_deps()):
        cmd = [NVCC] + NVCC_FLAGS + list(extra) + cuda_sources() + ["-o", LIB_CUDA, "-lcudart", "-lgomp"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB_CUDA


def build_all(force=False, verbose=False):
    build_gen(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
