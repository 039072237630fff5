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


# NCCL: the communicator of the multi-GPU repeated-SpMV loop lives inside the library (csrc/comm.cu)
LINK_LIBS = ["-lcudart", "-lnccl", "-lgomp", "-lrt", "-lpthread"]


def _digest(sources, flags):
    """Content hash of the sources + the command line (mtimes do not survive a snapshot copy to the GPU box)."""
    import hashlib
    # absolute paths inside the flags (-I<repo>/include ...) must not enter the digest: the GPU box unpacks the repo
    # somewhere else, and a digest that changes with the location makes EVERY process there rebuild the library
    h = hashlib.sha256(" ".join(f.replace(ROOT, "<repo>") for f in flags).encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, flags=()):
    """True when `target` is missing or was built from other sources / flags than the current ones (the digest of what
    it was built from sits next to it in <target>.srchash, git-ignored like the .so itself)."""
    if not os.path.exists(target):
        return True
    try:
        with open(target + ".srchash") as f:
            return f.read().strip() != _digest(sources, flags)
    except OSError:
        return True


def _stamp(target, sources, flags=()):
    with open(target + ".srchash", "w") as f:
        f.write(_digest(sources, flags))


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def cuda_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "tilespmv.h"))
    return deps


def build_gen(force=False, verbose=False):
    src = os.path.join(CSRC, "gen.c")
    flags = ["-O3", "-fopenmp", "-fPIC", "-shared"]
    if force or _stale(LIB_GEN, [src], flags):
        cmd = [GCC] + flags + [src, "-o", LIB_GEN]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        _stamp(LIB_GEN, [src], flags)
    return LIB_GEN


def build_cuda(force=False, verbose=False, extra=()):
    """Every .cu is compiled to its own object (in parallel, only when its sources or flags changed) and the objects are
    linked into libtilespmv_b200.so; nothing device-side crosses translation units, so no -rdc."""
    flags = NVCC_FLAGS + list(extra) + LINK_LIBS
    if not (force or _stale(LIB_CUDA, cuda_deps(), flags)):
        return LIB_CUDA
    # one builder at a time (the ranks of a torchrun job all come through here): take the lock, then look again
    import fcntl
    with open(os.path.join(PKG, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not (force or _stale(LIB_CUDA, cuda_deps(), flags)):
                return LIB_CUDA
            return _build_cuda_locked(force, verbose, extra, flags)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_cuda_locked(force, verbose, extra, flags):
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [d for d in cuda_deps() if not d.endswith(".cu")]
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + headers, cflags):
            cmd = [NVCC] + cflags + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            _stamp(obj, [src] + headers, cflags)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, cuda_sources()))
    tmp = LIB_CUDA + ".tmp%d" % os.getpid()
    # the arch on the link line too: without it nvcc adds an empty default-arch (sm_52) image and warns about it
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", GPP] + objs + ["-o", tmp] + LINK_LIBS
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_CUDA)  # atomic: a concurrent reader never maps a half-written library
    _stamp(LIB_CUDA, cuda_deps(), flags)
    return LIB_CUDA


def build_all(force=False, verbose=False):
    build_gen(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
