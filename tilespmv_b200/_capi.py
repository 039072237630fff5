"""ctypes mirror of include/tilespmv.h (the C-ABI of libtilespmv_b200.so).

The reference is compiled C code without any binding layer, so the host side above the C-ABI is
kept as thin as possible: this module only declares the exported symbols and wraps raw pointers.
Loading fails loudly if the CUDA library is missing -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import build

PKG = os.path.dirname(os.path.abspath(__file__))

F64, F32 = 8, 4
CSR_ON_DEVICE = 1
ENABLE_HYB = 2
PLAN_NO_CSR_GROUPS = 1
PLAN_NO_FLAT_SIDE = 2
OK = 0
ERR_NODEVICE = -5

_INT_P = C.POINTER(C.c_int)
_UCHAR_P = C.POINTER(C.c_ubyte)
_CHAR_P = C.POINTER(C.c_byte)


def _tile_matrix(val_ctype, name):
    VP = C.POINTER(val_ctype)
    fields = [
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
    return type(name, (C.Structure,), {"_fields_": fields})


TileMatrixF64 = _tile_matrix(C.c_double, "Tile_matrix_f64")
TileMatrixF32 = _tile_matrix(C.c_float, "Tile_matrix_f32")


class DmatInfo(C.Structure):
    _fields_ = [("precision", C.c_int), ("rowA", C.c_int), ("colA", C.c_int), ("tilem", C.c_int),
                ("tilen", C.c_int), ("tilenum", C.c_int), ("nnz", C.c_int64), ("nnz_side", C.c_int64),
                ("tiles_by_format", C.c_int64 * 7), ("device_bytes", C.c_int64), ("slots_by_format", C.c_int64 * 7)]


class PlanOptions(C.Structure):
    _fields_ = [("chunk_bytes", C.c_int), ("xstage_bytes", C.c_int), ("ctas_per_sm", C.c_int),
                ("stages", C.c_int), ("max_warps", C.c_int), ("flags", C.c_int), ("xpanel_bytes", C.c_int), ("format_mask", C.c_int)]


class DistInfo(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("row0", C.c_int64), ("rows", C.c_int64),
                ("slice_bytes", C.c_int64), ("device_bytes", C.c_int64), ("launch_units", C.c_int),
                ("equal_slices", C.c_int), ("unit_deps", C.c_uint32 * 64), ("halo_eligible", C.c_int),
                ("need_lo", C.c_int64), ("need_hi", C.c_int64)]


COMM_NCCL = 1
DIST_UNIFORM_PANELS = 1
EXCHANGE_NCCL, EXCHANGE_FUSED, EXCHANGE_PIPELINED, EXCHANGE_HALO = 0, 1, 2, 3


class PlanInfo(C.Structure):
    _fields_ = [("precision", C.c_int), ("nchunks", C.c_int64), ("stream_bytes", C.c_int64),
                ("algorithmic_bytes", C.c_int64), ("csr_bytes", C.c_int64), ("split_rows", C.c_int64),
                ("launches_per_spmv", C.c_int), ("grid", C.c_int), ("block", C.c_int),
                ("smem_bytes", C.c_int), ("chunk_bytes", C.c_int), ("xstage_bytes", C.c_int),
                ("device_bytes", C.c_int64), ("csr_groups", C.c_int64), ("xpanels", C.c_int64)]


# every symbol include/tilespmv.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "Tile_create_f64", "Tile_create_f32", "Tile_destroy_f64", "Tile_destroy_f32",
    "tilespmv_prepare_f64", "tilespmv_prepare_f32", "call_tilespmv_cuda_f64", "call_tilespmv_cuda_f32",
    "tilespmv_convert", "tilespmv_dmat_upload_f64", "tilespmv_dmat_upload_f32",
    "tilespmv_dmat_export_f64", "tilespmv_dmat_export_f32", "tilespmv_dmat_destroy",
    "tilespmv_dmat_get_info", "tilespmv_plan_create", "tilespmv_plan_destroy", "tilespmv_plan_save", "tilespmv_plan_load", "tilespmv_plan_spmv",
    "tilespmv_plan_spmv_host", "tilespmv_plan_spmv_host_batch", "tilespmv_plan_iterate", "tilespmv_partition_rows", "tilespmv_plan_set_peers", "tilespmv_plan_get_info", "tilespmv_plan_time", "tilespmv_format_profile",
    "tilespmv_mmio_allinone_f64", "tilespmv_mmio_allinone_f32", "tilespmv_last_error",
    "tilespmv_version", "tilespmv_kernel_launch_count",
    "tilespmv_comm_create", "tilespmv_comm_destroy", "tilespmv_comm_barrier", "tilespmv_dist_create",
    "tilespmv_dist_destroy", "tilespmv_dist_iterate", "tilespmv_dist_x", "tilespmv_dist_plan", "tilespmv_dist_sync",
    "tilespmv_dist_get_info",
]

_lib = None


def lib_path():
    return os.environ.get("TILESPMV_LIB_PATH") or build.LIB_CUDA


def load(rebuild=False):
    """Loads libtilespmv_b200.so (building it in-tree if the sources are newer)."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    path = os.environ.get("TILESPMV_LIB_PATH") or build.LIB_CUDA  # the override is for A/B runs of two builds on one box
    # build_cuda() is a no-op unless a source differs from what the .so was built from: a stale binary whose struct
    # layouts differ from the ctypes mirrors below must never be loaded silently
    if path == build.LIB_CUDA:
        build.build_cuda(force=rebuild)
    if not os.path.exists(path):
        raise RuntimeError("libtilespmv_b200.so is missing and could not be built; there is no CPU fallback")
    _preload_nccl()
    L = C.CDLL(path)
    L.tilespmv_last_error.restype = C.c_char_p
    L.tilespmv_version.restype = C.c_char_p
    L.tilespmv_kernel_launch_count.restype = C.c_int64
    for name in ("Tile_create_f64", "Tile_create_f32", "Tile_destroy_f64", "Tile_destroy_f32",
                 "call_tilespmv_cuda_f64", "call_tilespmv_cuda_f32", "tilespmv_dmat_destroy",
                 "tilespmv_plan_destroy"):
        getattr(L, name).restype = None
    L.tilespmv_plan_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tilespmv_plan_spmv_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.tilespmv_plan_spmv_host_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.tilespmv_partition_rows.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.tilespmv_plan_iterate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.tilespmv_plan_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.POINTER(C.c_double)]
    L.tilespmv_format_profile.argtypes = [C.c_void_p, C.POINTER(PlanOptions), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.tilespmv_plan_set_peers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int64]
    L.tilespmv_plan_create.argtypes = [C.c_void_p, C.POINTER(PlanOptions), C.POINTER(C.c_void_p)]
    L.tilespmv_plan_destroy.argtypes = [C.c_void_p]
    L.tilespmv_plan_save.argtypes = [C.c_void_p, C.c_char_p]
    L.tilespmv_plan_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.tilespmv_plan_get_info.argtypes = [C.c_void_p, C.POINTER(PlanInfo)]
    L.tilespmv_dmat_destroy.argtypes = [C.c_void_p]
    L.tilespmv_dmat_get_info.argtypes = [C.c_void_p, C.POINTER(DmatInfo)]
    L.tilespmv_convert.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                   C.POINTER(C.c_void_p)]
    L.tilespmv_dmat_export_f64.argtypes = [C.c_void_p, C.POINTER(TileMatrixF64)]
    L.tilespmv_dmat_export_f32.argtypes = [C.c_void_p, C.POINTER(TileMatrixF32)]
    L.tilespmv_dmat_upload_f64.argtypes = [C.POINTER(TileMatrixF64), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.tilespmv_dmat_upload_f32.argtypes = [C.POINTER(TileMatrixF32), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.tilespmv_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_void_p)]
    L.tilespmv_comm_destroy.argtypes = [C.c_void_p]
    L.tilespmv_comm_destroy.restype = None
    L.tilespmv_comm_barrier.argtypes = [C.c_void_p]
    L.tilespmv_dist_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(PlanOptions), C.c_uint,
                                       C.POINTER(C.c_void_p)]
    L.tilespmv_dist_destroy.argtypes = [C.c_void_p]
    L.tilespmv_dist_destroy.restype = None
    L.tilespmv_dist_iterate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tilespmv_dist_x.argtypes = [C.c_void_p]
    L.tilespmv_dist_x.restype = C.c_void_p
    L.tilespmv_dist_plan.argtypes = [C.c_void_p]
    L.tilespmv_dist_plan.restype = C.c_void_p
    L.tilespmv_dist_sync.argtypes = [C.c_void_p, C.c_void_p]
    L.tilespmv_dist_get_info.argtypes = [C.c_void_p, C.POINTER(DistInfo)]
    _lib = L
    return L


def _preload_nccl():
    """libtilespmv_b200.so links libnccl.so.2.  In a process that also runs torch the SAME NCCL must serve both, and
    torch ships its own (newer) copy: load that one first so the library's DT_NEEDED resolves to it whatever the
    import order.  A plain C host simply gets the system libnccl."""
    import glob
    import sysconfig
    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        for so in glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")):
            try:
                C.CDLL(so, mode=C.RTLD_GLOBAL)
                return so
            except OSError:
                pass
    return None


class TileSpMVError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != OK:
        msg = load().tilespmv_last_error().decode()
        raise TileSpMVError(f"{what} failed with status {rc}: {msg}")


def _np_from(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)),
                                 shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


def tile_matrix_arrays(M, rowA, val_dtype):
    """Every array of a host Tile_matrix as numpy copies (lengths of SURVEY.md A.1)."""
    T = M.tilenum
    i4, u1, i1 = np.int32, np.uint8, np.int8
    out = {
        "scalars": np.array([M.tilem, M.tilen, M.tilenum, M.csrsize, M.csrptrlen, M.coosize, M.ellsize,
                             M.hybsize, M.hybellsize, M.hybcoosize, M.dnssize, M.dnsrowsize, M.dnscolsize,
                             M.coototal], dtype=np.int64),
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
    for name in ("csr_offset", "csrptr_offset", "coo_offset", "ell_offset", "hyb_offset", "hyb_coocount",
                 "dns_offset", "dnsrow_offset", "dnscol_offset", "new_coocount"):
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
    ndr = int(out["dnsrowptr"][T]) if len(out["dnsrowptr"]) else 0
    ndc = int(out["dnscolptr"][T]) if len(out["dnscolptr"]) else 0
    out["denserowid"] = _np_from(M.denserowid, ndr, i1)
    out["Blockdensecol_Val"] = _np_from(M.Blockdensecol_Val, M.dnscolsize, val_dtype)
    out["densecolid"] = _np_from(M.densecolid, ndc, i1)
    out["deferredcoo_ptr"] = _np_from(M.deferredcoo_ptr, rowA + 1, i4)
    out["deferredcoo_colidx"] = _np_from(M.deferredcoo_colidx, M.coototal, i4)
    out["deferredcoo_val"] = _np_from(M.deferredcoo_val, M.coototal, val_dtype)
    return out
