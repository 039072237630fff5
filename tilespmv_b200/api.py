"""Host-side mirror of the reference's operator surface over the C-ABI.

Same names and argument meaning as the reference driver uses them (src/main.cu:87-180):

    M  = Tile_create(rowA, colA, csrRowPtrA, csrColIdxA, csrValA)       # csr2tile.h:629
    p1, p2, rbb, rowidx, cstart, cstop = tilespmv_prepare(M, rowA)      # tilespmv_cpu.h:68-118
    y  = call_tilespmv_cuda(filename, M, rowA, colA, nnzA, x)           # tilespmv_cuda.h:794

plus the handle API (DeviceTileMatrix / Plan) for device-resident repeated SpMV.  Everything here
calls libtilespmv_b200.so; nothing computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import F32, F64, TileSpMVError, check  # noqa: F401


def _precision_of(val):
    dt = np.asarray(val).dtype
    if dt == np.float64:
        return F64
    if dt == np.float32:
        return F32
    raise TypeError(f"values must be float64 or float32, not {dt}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostTileMatrix:
    """A host Tile_matrix (format.h:3-56) owned by this object; freed with Tile_destroy."""

    def __init__(self, precision, rowA, colA):
        self.precision, self.rowA, self.colA = precision, rowA, colA
        self.struct = (_capi.TileMatrixF64 if precision == F64 else _capi.TileMatrixF32)()
        self.val_dtype = np.float64 if precision == F64 else np.float32
        self._owned = False

    @property
    def tilenum(self):
        return self.struct.tilenum

    def arrays(self):
        return _capi.tile_matrix_arrays(self.struct, self.rowA, self.val_dtype)

    def destroy(self):
        if self._owned:
            L = _capi.load()
            (L.Tile_destroy_f64 if self.precision == F64 else L.Tile_destroy_f32)(C.byref(self.struct))
            self._owned = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def Tile_create(rowA, colA, csrRowPtrA, csrColIdxA, csrValA):
    """Tile_create (csr2tile.h:629-635): host CSR -> host Tile_matrix, converted on the GPU."""
    L = _capi.load()
    precision = _precision_of(csrValA)
    rp = np.ascontiguousarray(csrRowPtrA, np.int32)
    ci = np.ascontiguousarray(csrColIdxA, np.int32)
    v = np.ascontiguousarray(csrValA)
    M = HostTileMatrix(precision, rowA, colA)
    fn = L.Tile_create_f64 if precision == F64 else L.Tile_create_f32
    fn(C.byref(M.struct), C.c_int(rowA), C.c_int(colA), C.c_int(len(ci)), _ptr(rp), _ptr(ci), _ptr(v))
    if M.struct.tilenum < 0:
        raise TileSpMVError("Tile_create failed: " + L.tilespmv_last_error().decode())
    M._owned = True
    return M


def tilespmv_prepare(M, rowA):
    """ptroffset1/2 and the warp-chunk schedule the reference gets from tilespmv_cpu
    (tilespmv_cpu.h:68-118, :142-257) -- without any CPU SpMV."""
    L = _capi.load()
    T = max(M.struct.tilenum, 1)
    p1, p2 = np.zeros(T, np.int32), np.zeros(T, np.int32)
    rbb = C.c_int(0)
    a, b, c = C.POINTER(C.c_uint)(), C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    fn = L.tilespmv_prepare_f64 if M.precision == F64 else L.tilespmv_prepare_f32
    check(fn(C.byref(M.struct), _ptr(p1), _ptr(p2), C.byref(rbb), C.byref(a), C.byref(b), C.byref(c),
             C.c_int(rowA)), "tilespmv_prepare")
    n = rbb.value
    out = (p1[:M.struct.tilenum], p2[:M.struct.tilenum], n, _capi._np_from(a, n, np.uint32),
           _capi._np_from(b, n, np.int32), _capi._np_from(c, n, np.int32))
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for p in (a, b, c):
        libc.free(C.cast(p, C.c_void_p))
    return out


def call_tilespmv_cuda(filename, M, rowA, colA, nnzA, x, alpha=1.0):
    """call_tilespmv_cuda (tilespmv_cuda.h:794-809): host x -> host y = A*x; prints the reference's
    runtime line and appends to ./results.csv."""
    L = _capi.load()
    x = np.ascontiguousarray(x, M.val_dtype)
    y = np.zeros(rowA, M.val_dtype)
    null = C.c_void_p(None)
    vt = C.c_double if M.precision == F64 else C.c_float
    fn = L.call_tilespmv_cuda_f64 if M.precision == F64 else L.call_tilespmv_cuda_f32
    fn(C.c_char_p(filename.encode()), C.byref(M.struct), null, null, C.c_int(0), null, null, null,
       C.c_int(rowA), C.c_int(colA), C.c_int(nnzA), null, null, null, vt(alpha), _ptr(x), _ptr(y), null)
    err = L.tilespmv_last_error().decode()  # the entry clears the error string first: anything left is this call's
    if err:
        raise TileSpMVError(err)
    return y


class DeviceTileMatrix:
    """tilespmv_dmat: a Tile_matrix resident in device memory."""

    def __init__(self, handle, precision, rowA, colA):
        self.handle, self.precision, self.rowA, self.colA = handle, precision, rowA, colA
        self.val_dtype = np.float64 if precision == F64 else np.float32

    @classmethod
    def from_csr(cls, rowA, colA, rowptr, colidx, val, on_device=False, precision=None, enable_hyb=False):
        """GPU csr2tile.  Host numpy arrays, or raw device pointers (ints) with on_device=True.
        enable_hyb switches on the reference's dormant HYB rule (csr2tile.h:279-316); default off."""
        L = _capi.load()
        h = C.c_void_p()
        hyb = _capi.ENABLE_HYB if enable_hyb else 0
        if on_device:
            check(L.tilespmv_convert(precision, rowA, colA, C.c_void_p(rowptr), C.c_void_p(colidx),
                                     C.c_void_p(val), _capi.CSR_ON_DEVICE | hyb, C.byref(h)), "tilespmv_convert")
        else:
            precision = _precision_of(val)
            rp = np.ascontiguousarray(rowptr, np.int32)
            ci = np.ascontiguousarray(colidx, np.int32)
            v = np.ascontiguousarray(val)
            check(L.tilespmv_convert(precision, rowA, colA, _ptr(rp), _ptr(ci), _ptr(v), hyb, C.byref(h)),
                  "tilespmv_convert")
        return cls(h, precision, rowA, colA)

    @classmethod
    def upload(cls, M):
        L = _capi.load()
        h = C.c_void_p()
        fn = L.tilespmv_dmat_upload_f64 if M.precision == F64 else L.tilespmv_dmat_upload_f32
        check(fn(C.byref(M.struct), M.rowA, M.colA, C.byref(h)), "tilespmv_dmat_upload")
        return cls(h, M.precision, M.rowA, M.colA)

    def export(self):
        L = _capi.load()
        M = HostTileMatrix(self.precision, self.rowA, self.colA)
        fn = L.tilespmv_dmat_export_f64 if self.precision == F64 else L.tilespmv_dmat_export_f32
        check(fn(self.handle, C.byref(M.struct)), "tilespmv_dmat_export")
        M._owned = True
        return M

    def info(self):
        i = _capi.DmatInfo()
        check(_capi.load().tilespmv_dmat_get_info(self.handle, C.byref(i)), "tilespmv_dmat_get_info")
        return i

    def destroy(self):
        if self.handle:
            _capi.load().tilespmv_dmat_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Plan:
    """tilespmv_plan: packed stream + persistent chunk schedule for one DeviceTileMatrix."""

    def __init__(self, dmat, chunk_bytes=0, xstage_bytes=0, ctas_per_sm=0, stages=0, max_warps=0, csr_groups=True,
                 xpanel_bytes=0, format_mask=0, flat_side=True):
        L = _capi.load()
        flags = (0 if csr_groups else _capi.PLAN_NO_CSR_GROUPS) | (0 if flat_side else _capi.PLAN_NO_FLAT_SIDE)
        opts = _capi.PlanOptions(chunk_bytes, xstage_bytes, ctas_per_sm, stages, max_warps, flags, xpanel_bytes, format_mask)
        h = C.c_void_p()
        check(L.tilespmv_plan_create(dmat.handle, C.byref(opts), C.byref(h)), "tilespmv_plan_create")
        self.handle, self.precision = h, dmat.precision
        self.rowA, self.colA = dmat.rowA, dmat.colA
        self.val_dtype = dmat.val_dtype

    def save(self, path):
        """tilespmv_plan_save: binary cache of the packed plan (and its x-panel sub-plans)."""
        check(_capi.load().tilespmv_plan_save(self.handle, path.encode()), "tilespmv_plan_save")

    @classmethod
    def load(cls, path, precision, rowA, colA):
        """tilespmv_plan_load: a plan saved on a GPU with the same SM count / shared memory; raises TileSpMVError otherwise."""
        h = C.c_void_p()
        check(_capi.load().tilespmv_plan_load(path.encode(), C.byref(h)), "tilespmv_plan_load")
        self = cls.__new__(cls)
        self.handle, self.precision, self.rowA, self.colA = h, precision, rowA, colA
        self.val_dtype = np.float64 if precision == F64 else np.float32
        return self

    def info(self):
        i = _capi.PlanInfo()
        check(_capi.load().tilespmv_plan_get_info(self.handle, C.byref(i)), "tilespmv_plan_get_info")
        return i

    def spmv(self, d_x, d_y, stream=0):
        """y = A*x on raw device pointers (e.g. torch.Tensor.data_ptr()), async on `stream`."""
        check(_capi.load().tilespmv_plan_spmv(self.handle, C.c_void_p(d_x), C.c_void_p(d_y),
                                               C.c_void_p(stream)), "tilespmv_plan_spmv")

    def spmv_host(self, x, y=None):
        """y = A*x on host numpy arrays: H2D, SpMV, D2H (the end-to-end path)."""
        x = np.ascontiguousarray(x, self.val_dtype)
        if y is None:
            y = np.empty(self.rowA, self.val_dtype)
        check(_capi.load().tilespmv_plan_spmv_host(self.handle, _ptr(x), _ptr(y)), "tilespmv_plan_spmv_host")
        return y

    def spmv_host_batch(self, x_ptrs, y_ptrs):
        """y[i] = A*x[i] for lists of raw HOST pointers (ints; pinned memory for full overlap): H2D of vector
        i+1, SpMV of vector i and D2H of vector i-1 run concurrently (tilespmv_plan_spmv_host_batch)."""
        n = len(x_ptrs)
        assert len(y_ptrs) == n
        xa = (C.c_void_p * max(n, 1))(*[C.c_void_p(p) for p in x_ptrs])
        ya = (C.c_void_p * max(n, 1))(*[C.c_void_p(p) for p in y_ptrs])
        check(_capi.load().tilespmv_plan_spmv_host_batch(self.handle, n, xa, ya), "tilespmv_plan_spmv_host_batch")

    def iterate(self, d_xa, d_xb, niters, stream=0):
        """x <- A*x niters times, ping-pong between two device buffers, replayed as one CUDA graph; the result is in
        d_xa when niters is even, else in d_xb (tilespmv_plan_iterate)."""
        check(_capi.load().tilespmv_plan_iterate(self.handle, C.c_void_p(d_xa), C.c_void_p(d_xb), niters,
                                                  C.c_void_p(stream)), "tilespmv_plan_iterate")

    def time(self, d_x, d_y, warmup=3, iters=20, stream=0):
        ms = C.c_double(0)
        check(_capi.load().tilespmv_plan_time(self.handle, C.c_void_p(d_x), C.c_void_p(d_y), warmup, iters,
                                               C.c_void_p(stream), C.byref(ms)), "tilespmv_plan_time")
        return ms.value

    def set_peers(self, peer_ptrs, row_offset):
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(p) for p in peer_ptrs])
        check(_capi.load().tilespmv_plan_set_peers(self.handle, len(peer_ptrs), arr, C.c_int64(row_offset)),
              "tilespmv_plan_set_peers")

    def destroy(self):
        if self.handle:
            _capi.load().tilespmv_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def format_profile(dmat, d_x, d_y, warmup=3, iters=20):
    """tilespmv_format_profile: ms per SpMV restricted to each tile format (0..6), to no format (7) and for the whole
    matrix (8), and the nonzeros each of those plans multiplies."""
    ms = (C.c_double * 9)()
    nnz = (C.c_int64 * 9)()
    check(_capi.load().tilespmv_format_profile(dmat.handle, None, C.c_void_p(d_x), C.c_void_p(d_y), warmup, iters, ms, nnz),
          "tilespmv_format_profile")
    return list(ms), list(nnz)


def mmio_allinone(filename, precision=F64):
    """mmio_allinone (mmio_highlevel.h:593-759): returns (rc, m, n, isSymmetric, rowptr, colidx, val)."""
    L = _capi.load()
    m, n, nnz, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rp, ci = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    vt = C.c_double if precision == F64 else C.c_float
    cv = C.POINTER(vt)()
    fn = L.tilespmv_mmio_allinone_f64 if precision == F64 else L.tilespmv_mmio_allinone_f32
    rc = fn(C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym), C.byref(rp), C.byref(ci), C.byref(cv),
            C.c_char_p(filename.encode()))
    if rc != 0:
        return rc, None
    dt = np.float64 if precision == F64 else np.float32
    out = (m.value, n.value, sym.value, _capi._np_from(rp, m.value + 1, np.int32),
           _capi._np_from(ci, nnz.value, np.int32), _capi._np_from(cv, nnz.value, dt))
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for p in (rp, ci, cv):
        libc.free(C.cast(p, C.c_void_p))
    return 0, out
