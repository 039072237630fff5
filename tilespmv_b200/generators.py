"""Synthetic CSR inputs of the BASELINE.json configs (ctypes over csrc/gen.c).

Every function returns ``(m, n, rowptr int32[m+1], colidx int32[nnz], val float64[nnz])`` with
ascending columns inside each row.  ``val_mode=1`` gives the reference driver's integer data
``val[j] = j % 10`` (/root/reference/src/main.cu:68-69) for which fp64 sums are exact.
"""
import ctypes as C

import numpy as np

from . import build

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build.build_gen())
        for name in ("tsgen_lap2d", "tsgen_lap3d27", "tsgen_banded", "tsgen_band_contig",
                     "tsgen_uniform_rows", "tsgen_banded_rows", "tsgen_rmat", "tsgen_seven_formats", "tsgen_lap3d27_slab"):
            getattr(_lib, name).restype = C.c_int64
    return _lib


_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)
_NULLS = (_IP(), _IP(), _DP())


def _alloc(m, nnz):
    if nnz >= 2 ** 31:
        raise ValueError("nnz does not fit the reference's int indexing (MAT_PTR_TYPE int)")
    return np.zeros(m + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)


def _ptrs(rp, ci, v):
    return rp.ctypes.data_as(_IP), ci.ctypes.data_as(_IP), v.ctypes.data_as(_DP)


def lap2d(G, val_mode=0):
    """2-D 5-point Laplacian on a G x G grid, Dirichlet (BASELINE config 1 at G=1024)."""
    f = _L().tsgen_lap2d
    nnz = f(C.c_int(G), C.c_int(val_mode), *_NULLS)
    rp, ci, v = _alloc(G * G, nnz)
    f(C.c_int(G), C.c_int(val_mode), *_ptrs(rp, ci, v))
    return G * G, G * G, rp, ci, v


def lap3d27(G, val_mode=0):
    """3-D 27-point Laplacian on a G^3 grid, Dirichlet (BASELINE config 2 at G=160)."""
    f = _L().tsgen_lap3d27
    nnz = f(C.c_int(G), C.c_int(val_mode), *_NULLS)
    rp, ci, v = _alloc(G ** 3, nnz)
    f(C.c_int(G), C.c_int(val_mode), *_ptrs(rp, ci, v))
    return G ** 3, G ** 3, rp, ci, v


def lap3d27_slab(nx, ny, nz, i0, i1, val_mode=0):
    """Planes [i0,i1) of the 27-point Laplacian on an nx x ny x nz grid: local rows, global columns
    (the row block one GPU owns in multi-GPU runs).  Returns (m_local, n_global, rowptr, colidx, val)."""
    f = _L().tsgen_lap3d27_slab
    a = (C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(i0), C.c_int(i1), C.c_int(val_mode))
    nnz = f(*a, *_NULLS)
    m = (i1 - i0) * ny * nz
    rp, ci, v = _alloc(m, nnz)
    f(*a, *_ptrs(rp, ci, v))
    return m, nx * ny * nz, rp, ci, v


def banded(N, hb=64, per_row=36, seed=3, val_mode=0):
    """Banded FEM-like matrix: diagonal + per_row random offsets within +-hb (config 3)."""
    f = _L().tsgen_banded
    a = (C.c_int64(N), C.c_int(hb), C.c_int(per_row), C.c_uint64(seed), C.c_int(val_mode))
    nnz = f(*a, *_NULLS)
    rp, ci, v = _alloc(N, nnz)
    f(*a, *_ptrs(rp, ci, v))
    return N, N, rp, ci, v


def banded_rows(N, row0, nrows, hb=64, per_row=36, seed=3, val_mode=0):
    """Rows [row0, row0+nrows) of banded(N, ...): local rowptr, global columns, the same values as the full matrix."""
    f = _L().tsgen_banded_rows
    a = (C.c_int64(N), C.c_int64(row0), C.c_int64(nrows), C.c_int(hb), C.c_int(per_row), C.c_uint64(seed), C.c_int(val_mode))
    nnz = f(*a, *_NULLS)
    if nnz < 0:
        raise ValueError("row range outside the matrix")
    rp, ci, v = _alloc(nrows, nnz)
    f(*a, *_ptrs(rp, ci, v))
    return nrows, N, rp, ci, v


def band_contig(N, hb=18, seed=3, val_mode=0):
    """Contiguous band |i-j| <= hb: Dense + CSR + COO tile mix (config 3b)."""
    f = _L().tsgen_band_contig
    a = (C.c_int64(N), C.c_int(hb), C.c_uint64(seed), C.c_int(val_mode))
    nnz = f(*a, *_NULLS)
    rp, ci, v = _alloc(N, nnz)
    f(*a, *_ptrs(rp, ci, v))
    return N, N, rp, ci, v


def uniform_rows(ncols, row0, nrows, per_row=20, seed=5, val_mode=0):
    """Rows [row0,row0+nrows) of the uniform random matrix with per_row distinct columns per row
    (config 5); shards generate only their own row block."""
    f = _L().tsgen_uniform_rows
    a = (C.c_int64(ncols), C.c_int64(row0), C.c_int64(nrows), C.c_int(per_row), C.c_uint64(seed),
         C.c_int(val_mode))
    nnz = f(*a, *_NULLS)
    rp, ci, v = _alloc(nrows, nnz)
    f(*a, *_ptrs(rp, ci, v))
    return nrows, ncols, rp, ci, v


def uniform(N, per_row=20, seed=5, val_mode=0):
    return uniform_rows(N, 0, N, per_row, seed, val_mode)


def rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=4, val_mode=0):
    """R-MAT power-law graph, duplicates merged, no permutation (config 4 at scale 24)."""
    f = _L().tsgen_rmat
    n = 1 << scale
    args = (C.c_int(scale), C.c_int(edge_factor), C.c_double(a), C.c_double(b), C.c_double(c),
            C.c_uint64(seed), C.c_int(val_mode))
    nnz = f(*args, *_NULLS, C.c_int64(0))
    if nnz < 0:
        raise MemoryError("rmat generator failed")
    rp, ci, v = _alloc(n, nnz)
    got = f(*args, *_ptrs(rp, ci, v), C.c_int64(nnz))
    assert got == nnz
    return n, n, rp, ci, v


def seven_formats():
    """The 32 x 40, 385-nnz fixture whose tiles are DenseRow, DenseCol, COO, Dense, CSR, ELL."""
    f = _L().tsgen_seven_formats
    nnz = f(*_NULLS)
    rp, ci, v = _alloc(32, nnz)
    f(*_ptrs(rp, ci, v))
    return 32, 40, rp, ci, v


def row_slice(rowptr, colidx, val, r0, r1):
    """CSR rows [r0, r1) as a self-contained CSR (row-block shard; global columns kept)."""
    lo, hi = int(rowptr[r0]), int(rowptr[r1])
    return (rowptr[r0:r1 + 1] - rowptr[r0]).astype(np.int32), colidx[lo:hi], val[lo:hi]


def write_mtx(path, m, n, rowptr, colidx, val, column_major=True):
    """Matrix Market 'coordinate real general', entries sorted column-major like SuiteSparse files."""
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(rowptr))
    cols = colidx.astype(np.int64)
    order = np.lexsort((rows, cols)) if column_major else np.arange(len(cols))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{m} {n} {len(cols)}\n")
        for k in order:
            f.write(f"{rows[k] + 1} {cols[k] + 1} {float(val[k])!r}\n")


def write_mtx_fast(path, m, n, rowptr, colidx, val):
    """C writer (row-major order) for large inputs of the reference's own driver."""
    rp = np.ascontiguousarray(rowptr, np.int32)
    ci = np.ascontiguousarray(colidx, np.int32)
    v = np.ascontiguousarray(val, np.float64)
    L = _L()
    L.tsgen_write_mtx.restype = C.c_int
    rc = L.tsgen_write_mtx(C.c_char_p(path.encode()), C.c_int(m), C.c_int(n), *_ptrs(rp, ci, v))
    if rc != 0:
        raise OSError(f"cannot write {path}")
