"""ctypes access to the parity checkers -- TEST INFRASTRUCTURE ONLY.

Loads (building on demand) the C restatement ``oracle/libtilespmv_oracle_{f64,f32}.so`` and, when
it exists, the unmodified reference CPU path ``oracle/_ref/libtilespmv_ref_{f64,f32}.so``
(compiled from /root/reference/src by oracle/Makefile; see oracle/ref_shim.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package (tilespmv_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_INT_P = C.POINTER(C.c_int)
_UCHAR_P = C.POINTER(C.c_ubyte)
_CHAR_P = C.POINTER(C.c_byte)


def tile_matrix_struct(val_ctype):
    """Field-for-field mirror of Tile_matrix, /root/reference/src/format.h:3-56."""
    VP = C.POINTER(val_ctype)

    class TileMatrix(C.Structure):
        _fields_ = [
            ("tilem", C.c_int), ("tilen", C.c_int), ("tilenum", C.c_int),
            ("tile_ptr", _INT_P), ("tile_columnidx", _INT_P), ("tile_nnz", _INT_P),
            ("Format", _CHAR_P), ("blknnz", _INT_P), ("blknnznnz", _UCHAR_P),
            ("dnsrowptr", _INT_P), ("dnscolptr", _INT_P), ("tilewidth", _CHAR_P),
            ("csr_offset", _INT_P), ("csrptr_offset", _INT_P), ("coo_offset", _INT_P),
            ("ell_offset", _INT_P), ("hyb_offset", _INT_P), ("hyb_coocount", _INT_P),
            ("dns_offset", _INT_P), ("dnsrow_offset", _INT_P), ("dnscol_offset", _INT_P),
            ("new_coocount", _INT_P),
            ("Blockcsr_Val", VP), ("Blockcsr_Ptr", _UCHAR_P), ("csr_compressedIdx", _UCHAR_P),
            ("csrsize", C.c_int), ("csrptrlen", C.c_int),
            ("Blockcoo_Val", VP), ("coo_compressed_Idx", _UCHAR_P), ("coosize", C.c_int),
            ("Blockell_Val", VP), ("ell_compressedIdx", _UCHAR_P), ("ellsize", C.c_int),
            ("Blockhyb_Val", VP), ("hybIdx", _UCHAR_P), ("hybsize", C.c_int),
            ("hybellsize", C.c_int), ("hybcoosize", C.c_int),
            ("Blockdense_Val", VP), ("dnssize", C.c_int),
            ("Blockdenserow_Val", VP), ("denserowid", _CHAR_P), ("dnsrowsize", C.c_int),
            ("Blockdensecol_Val", VP), ("densecolid", _CHAR_P), ("dnscolsize", C.c_int),
            ("coototal", C.c_int),
            ("deferredcoo_val", VP), ("deferredcoo_colidx", _INT_P), ("deferredcoo_ptr", _INT_P),
        ]

    return TileMatrix


TileMatrixF64 = tile_matrix_struct(C.c_double)
TileMatrixF32 = tile_matrix_struct(C.c_float)


def _np_from(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)),
                                 shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


def tile_matrix_arrays(M, rowA, val_dtype):
    """Every array of a Tile_matrix as numpy copies, with the lengths of SURVEY.md A.1."""
    T = M.tilenum
    i4, u1, i1 = np.int32, np.uint8, np.int8
    out = {
        "scalars": np.array([M.tilem, M.tilen, M.tilenum, M.csrsize, M.csrptrlen, M.coosize,
                             M.ellsize, M.hybsize, M.hybellsize, M.hybcoosize, M.dnssize,
                             M.dnsrowsize, M.dnscolsize, M.coototal], dtype=np.int64),
        "tile_ptr": _np_from(M.tile_ptr, M.tilem + 1, i4),
        "tile_columnidx": _np_from(M.tile_columnidx, T, i4),
        "tile_nnz": _np_from(M.tile_nnz, T + 1, i4),
        "Format": _np_from(M.Format, T, i1),
        "blknnz": _np_from(M.blknnz, T + 1, i4),
        "blknnznnz": _np_from(M.blknnznnz, T + 1, u1),
        "dnsrowptr": _np_from(M.dnsrowptr, T + 1, i4),
        "dnscolptr": _np_from(M.dnscolptr, T + 1, i4),
        "tilewidth": _np_from(M.tilewidth, T, i1),
    }
    for name in ("csr_offset", "csrptr_offset", "coo_offset", "ell_offset", "hyb_offset",
                 "hyb_coocount", "dns_offset", "dnsrow_offset", "dnscol_offset", "new_coocount"):
        out[name] = _np_from(getattr(M, name), T + 1, i4)
    out["Blockcsr_Val"] = _np_from(M.Blockcsr_Val, M.csrsize, val_dtype)
    out["Blockcsr_Ptr"] = _np_from(M.Blockcsr_Ptr, M.csrptrlen, u1)
    out["csr_compressedIdx"] = _np_from(M.csr_compressedIdx, (M.csrsize + 1) // 2, u1)
    out["Blockcoo_Val"] = _np_from(M.Blockcoo_Val, M.coosize, val_dtype)
    out["coo_compressed_Idx"] = _np_from(M.coo_compressed_Idx, M.coosize, u1)
    out["Blockell_Val"] = _np_from(M.Blockell_Val, M.ellsize, val_dtype)
    out["ell_compressedIdx"] = _np_from(M.ell_compressedIdx, (M.ellsize + 1) // 2, u1)
    out["Blockhyb_Val"] = _np_from(M.Blockhyb_Val, M.hybellsize + M.hybcoosize, val_dtype)
    out["hybIdx"] = _np_from(M.hybIdx, (M.hybellsize + 1) // 2 + M.hybcoosize, u1)
    out["Blockdense_Val"] = _np_from(M.Blockdense_Val, M.dnssize, val_dtype)
    out["Blockdenserow_Val"] = _np_from(M.Blockdenserow_Val, M.dnsrowsize, val_dtype)
    ndr = int(out["dnsrowptr"][T]) if T >= 0 and len(out["dnsrowptr"]) else 0
    ndc = int(out["dnscolptr"][T]) if T >= 0 and len(out["dnscolptr"]) else 0
    out["denserowid"] = _np_from(M.denserowid, ndr, i1)
    out["Blockdensecol_Val"] = _np_from(M.Blockdensecol_Val, M.dnscolsize, val_dtype)
    out["densecolid"] = _np_from(M.densecolid, ndc, i1)
    out["deferredcoo_ptr"] = _np_from(M.deferredcoo_ptr, rowA + 1, i4)
    out["deferredcoo_colidx"] = _np_from(M.deferredcoo_colidx, M.coototal, i4)
    out["deferredcoo_val"] = _np_from(M.deferredcoo_val, M.coototal, val_dtype)
    return out


def _ensure_built():
    want = [os.path.join(HERE, f"libtilespmv_oracle_{p}.so") for p in ("f64", "f32")]
    src = os.path.join(HERE, "tilespmv_oracle.c")
    if all(os.path.exists(w) and os.path.getmtime(w) >= os.path.getmtime(src) for w in want):
        return
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)


def build_all():
    """Compile the restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)


class _Lib:
    def __init__(self, path, prefix, precision):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.precision = precision
        self.val_ctype = C.c_double if precision == "f64" else C.c_float
        self.val_dtype = np.float64 if precision == "f64" else np.float32
        self.TileMatrix = TileMatrixF64 if precision == "f64" else TileMatrixF32
        assert getattr(self.lib, prefix + "_sizeof_val")() == np.dtype(self.val_dtype).itemsize
        assert getattr(self.lib, prefix + "_sizeof_tile_matrix")() == C.sizeof(self.TileMatrix)

    def fn(self, name, restype=None):
        f = getattr(self.lib, f"{self.prefix}_{name}")
        f.restype = restype
        return f


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Oracle:
    """The C restatement (oracle/tilespmv_oracle.c)."""

    kind = "port"

    def __init__(self, precision="f64", enable_hyb=False):
        """enable_hyb: the dormant HYB rule of csr2tile.h:279-316 switched on (pinned against the
        "refhyb" variant of Reference); default = the reference as shipped."""
        _ensure_built()
        self.L = _Lib(os.path.join(HERE, f"libtilespmv_oracle_{precision}.so"), "oracle", precision)
        self.val_dtype = self.L.val_dtype
        self.enable_hyb = bool(enable_hyb)

    def threads(self):
        return self.L.fn("omp_max_threads", C.c_int)()

    def tile_create(self, rowA, colA, rowptr, colidx, val):
        M = self.L.TileMatrix()
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        val = np.ascontiguousarray(val, self.val_dtype)
        self.L.fn("set_enable_hyb")(C.c_int(1 if self.enable_hyb else 0))
        self.L.fn("tile_create")(C.byref(M), C.c_int(rowA), C.c_int(colA), C.c_int(len(colidx)),
                                 _p(rowptr, C.c_int), _p(colidx, C.c_int), _p(val, self.L.val_ctype))
        return M

    def tile_destroy(self, M):
        self.L.fn("tile_destroy")(C.byref(M))

    def arrays(self, M, rowA):
        return tile_matrix_arrays(M, rowA, self.val_dtype)

    def schedule(self, M):
        rbb = C.c_int(0)
        a = C.POINTER(C.c_uint)()
        b = _INT_P()
        c = _INT_P()
        self.L.fn("build_schedule")(C.byref(M), C.byref(rbb), C.byref(a), C.byref(b), C.byref(c))
        n = rbb.value
        out = (n, _np_from(a, n, np.uint32), _np_from(b, n, np.int32), _np_from(c, n, np.int32))
        free = self.L.fn("free")
        for ptr in (a, b, c):
            free(ptr)
        return out

    def tilespmv_cpu(self, M, rowA, colA, x):
        x = np.ascontiguousarray(x, self.val_dtype)
        y = np.zeros(rowA, self.val_dtype)
        p1 = np.zeros(max(M.tilenum, 1), np.int32)
        p2 = np.zeros(max(M.tilenum, 1), np.int32)
        self.L.fn("tilespmv_cpu")(C.byref(M), _p(p1, C.c_int), _p(p2, C.c_int), C.c_int(rowA),
                                  C.c_int(colA), _p(x, self.L.val_ctype), _p(y, self.L.val_ctype))
        return y, p1[:M.tilenum], p2[:M.tilenum]

    def time_tilespmv_cpu(self, M, rowA, colA, x):
        x = np.ascontiguousarray(x, self.val_dtype)
        y = np.zeros(rowA, self.val_dtype)
        ms = self.L.fn("time_tilespmv_cpu", C.c_double)(C.byref(M), C.c_int(rowA), C.c_int(colA),
                                                        _p(x, self.L.val_ctype), _p(y, self.L.val_ctype))
        return ms, y

    def _csr(self, name, rowA, rowptr, colidx, val, x):
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        val = np.ascontiguousarray(val, self.val_dtype)
        x = np.ascontiguousarray(x, self.val_dtype)
        y = np.zeros(rowA, self.val_dtype)
        self.L.fn(name)(C.c_int(rowA), _p(rowptr, C.c_int), _p(colidx, C.c_int),
                        _p(val, self.L.val_ctype), _p(x, self.L.val_ctype), _p(y, self.L.val_ctype))
        return y

    def csr_spmv(self, rowA, rowptr, colidx, val, x, parallel=False):
        return self._csr("csr_spmv_omp" if parallel else "csr_spmv", rowA, rowptr, colidx, val, x)

    def csr_abs_spmv(self, rowA, rowptr, colidx, val, x):
        return self._csr("csr_abs_spmv", rowA, rowptr, colidx, val, x)

    def mtx_read(self, path):
        m, n, nnz, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rp, ci = _INT_P(), _INT_P()
        cv = C.POINTER(self.L.val_ctype)()
        rc = self.L.fn("mtx_read", C.c_int)(path.encode(), C.byref(m), C.byref(n), C.byref(nnz),
                                             C.byref(sym), C.byref(rp), C.byref(ci), C.byref(cv))
        if rc != 0:
            return rc, None
        out = (m.value, n.value, sym.value, _np_from(rp, m.value + 1, np.int32),
               _np_from(ci, nnz.value, np.int32), _np_from(cv, nnz.value, self.val_dtype))
        free = self.L.fn("free")
        for ptr in (rp, ci, cv):
            free(ptr)
        return 0, out


def ref_available(precision="f64", variant="ref"):
    return os.path.exists(os.path.join(HERE, "_ref", f"libtilespmv_{variant}_{precision}.so"))


class Reference:
    """The UNMODIFIED reference CPU path (oracle/_ref, built from /root/reference/src)."""

    kind = "reference"

    def __init__(self, precision="f64", variant="ref"):
        """variant "ref": the unmodified reference; "refhyb": the same sources with the dormant HYB
        selection rule of csr2tile.h:308-317 un-commented (oracle/Makefile, target ref_hyb)."""
        self.L = _Lib(os.path.join(HERE, "_ref", f"libtilespmv_{variant}_{precision}.so"), "ref", precision)
        self.val_dtype = self.L.val_dtype

    def threads(self):
        return self.L.fn("omp_max_threads", C.c_int)()

    def tile_create(self, rowA, colA, rowptr, colidx, val):
        """Tile_create, /root/reference/src/csr2tile.h:629 (prints the tile count to stdout)."""
        M = self.L.TileMatrix()
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        val = np.ascontiguousarray(val, self.val_dtype)
        self.L.fn("Tile_create")(C.byref(M), C.c_int(rowA), C.c_int(colA), C.c_int(len(colidx)),
                                 _p(rowptr, C.c_int), _p(colidx, C.c_int), _p(val, self.L.val_ctype))
        return M

    def arrays(self, M, rowA):
        return tile_matrix_arrays(M, rowA, self.val_dtype)

    def tilespmv_cpu(self, M, rowA, colA, rowptr, colidx, val, x, y_golden=None):
        """tilespmv_cpu, /root/reference/src/tilespmv_cpu.h:3: returns y, ptroffset1/2, schedule."""
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        val = np.ascontiguousarray(val, self.val_dtype)
        x = np.ascontiguousarray(x, self.val_dtype)
        # the reference zeroes 16 entries per block row even past rowA: give it room
        y = np.zeros(((rowA + 15) // 16) * 16, self.val_dtype)
        yg = np.zeros(rowA, self.val_dtype) if y_golden is None else np.ascontiguousarray(y_golden, self.val_dtype)
        p1 = np.zeros(max(M.tilenum, 1), np.int32)
        p2 = np.zeros(max(M.tilenum, 1), np.int32)
        rbb = C.c_int(0)
        a = C.POINTER(C.c_uint)()
        b = _INT_P()
        c = _INT_P()
        vt = self.L.val_ctype
        self.L.fn("tilespmv_cpu")(C.byref(M), _p(p1, C.c_int), _p(p2, C.c_int), C.byref(rbb),
                                  C.byref(a), C.byref(b), C.byref(c), C.c_int(rowA), C.c_int(colA),
                                  C.c_int(len(colidx)), _p(rowptr, C.c_int), _p(colidx, C.c_int),
                                  _p(val, vt), _p(x, vt), _p(y, vt), _p(yg, vt))
        n = rbb.value
        sched = (n, _np_from(a, n, np.uint32), _np_from(b, n, np.int32), _np_from(c, n, np.int32))
        free = self.L.fn("free")
        for ptr in (a, b, c):
            free(ptr)
        return y[:rowA].copy(), p1[:M.tilenum], p2[:M.tilenum], sched

    def time_tilespmv_cpu(self, M, rowA, colA, rowptr, colidx, val, x):
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        colidx = np.ascontiguousarray(colidx, np.int32)
        val = np.ascontiguousarray(val, self.val_dtype)
        x = np.ascontiguousarray(x, self.val_dtype)
        y = np.zeros(((rowA + 15) // 16) * 16, self.val_dtype)
        yg = np.zeros(rowA, self.val_dtype)
        vt = self.L.val_ctype
        ms = self.L.fn("time_tilespmv_cpu", C.c_double)(
            C.byref(M), C.c_int(rowA), C.c_int(colA), C.c_int(len(colidx)), _p(rowptr, C.c_int),
            _p(colidx, C.c_int), _p(val, vt), _p(x, vt), _p(y, vt), _p(yg, vt))
        return ms, y[:rowA].copy()

    def mtx_read(self, path):
        m, n, nnz, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rp, ci = _INT_P(), _INT_P()
        cv = C.POINTER(self.L.val_ctype)()
        rc = self.L.fn("mmio_allinone", C.c_int)(C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym),
                                                  C.byref(rp), C.byref(ci), C.byref(cv), path.encode())
        if rc != 0:
            return rc, None
        out = (m.value, n.value, sym.value, _np_from(rp, m.value + 1, np.int32),
               _np_from(ci, nnz.value, np.int32), _np_from(cv, nnz.value, self.val_dtype))
        free = self.L.fn("free")
        for ptr in (rp, ci, cv):
            free(ptr)
        return 0, out
