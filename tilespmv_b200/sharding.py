"""Row-block sharding of the TileSpMV path across the GPUs of one box.

Block rows of tiles are independent units (tilespmv_cpu.h:125-272: every y row depends on one block
row of tiles plus the read-only x), so the matrix is cut into contiguous ranges of block rows, one
per rank, at multiples of 16 rows; x stays replicated.  A single SpMV then needs no communication;
for repeated SpMV (x <- A*x, square A) every rank's y slice is its slice of the next x and the
slices -- of UNEQUAL length, because the cuts balance streamed bytes, not rows -- are all-gathered.

Host logic only (numpy + torch.distributed): the same code runs under gloo on CPU (tests) and
under NCCL on GPUs (bench.py).  The compute itself is always libtilespmv_b200.so.
"""
import numpy as np

TS = 16  # BLOCK_SIZE (common.h:37-39)


def block_row_weights(rowptr, rowA, value_bytes):
    """Streamed bytes per block row with the weights of B_alg: value + 4-bit index
    per nonzero, 16 y values written, one RowRec.  Tile headers are not known before conversion;
    they are second order (5 B per tile) and proportional to the nonzeros for a given matrix."""
    rowptr = np.asarray(rowptr, np.int64)
    tilem = (rowA + TS - 1) // TS
    edges = np.minimum(np.arange(tilem + 1, dtype=np.int64) * TS, rowA)
    nnz_br = rowptr[edges[1:]] - rowptr[edges[:-1]]
    return nnz_br * (value_bytes + 0.5) + TS * value_bytes + 16.0


def partition(weights, nranks):
    """Contiguous block-row ranges [b0, b1) per rank, cut where the prefix of `weights` crosses
    g/G of the total.  Every rank gets a (possibly empty) range; ranges tile [0, len(weights))."""
    w = np.asarray(weights, np.float64)
    n = len(w)
    pre = np.concatenate([[0.0], np.cumsum(w)])
    total = pre[-1]
    cuts = [0]
    for g in range(1, nranks):
        target = total * g / nranks
        b = int(np.searchsorted(pre, target, side="left"))
        # pick the nearer of the two neighbouring cut points
        if b > 0 and abs(pre[b - 1] - target) <= abs(pre[min(b, n)] - target):
            b -= 1
        b = min(max(b, cuts[-1]), n)
        cuts.append(b)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nranks)]


def row_ranges(parts, rowA):
    """Row ranges [r0, r1) of the block-row ranges, clipped to rowA."""
    return [(min(b0 * TS, rowA), min(b1 * TS, rowA)) for b0, b1 in parts]


def partition_rows(rowptr, rowA, nranks, value_bytes=8):
    """The same cuts computed by the library (tilespmv_partition_rows, host-only C): row ranges per rank."""
    import ctypes as C

    from . import _capi
    rp = np.ascontiguousarray(rowptr, np.int32)
    cuts = np.zeros(nranks + 1, np.int32)
    _capi.check(_capi.load().tilespmv_partition_rows(value_bytes, rowA, rp.ctypes.data_as(C.c_void_p), nranks,
                                                     cuts.ctypes.data_as(C.c_void_p)), "tilespmv_partition_rows")
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(nranks)]


def imbalance(weights, parts):
    """max over ranks of shard weight / mean shard weight (1.0 = perfect)."""
    w = np.asarray(weights, np.float64)
    s = np.array([w[b0:b1].sum() for b0, b1 in parts])
    return float(s.max() / s.mean()) if s.mean() > 0 else 1.0


def shard_csr(rowptr, colidx, val, r0, r1):
    """CSR of rows [r0, r1) with global column indices (views where possible)."""
    rowptr = np.asarray(rowptr)
    a, b = int(rowptr[r0]), int(rowptr[r1])
    rp = (rowptr[r0:r1 + 1] - rowptr[r0]).astype(np.int32)
    return rp, np.ascontiguousarray(colidx[a:b]), np.ascontiguousarray(val[a:b])


def allgather_rows(dist, y_local, rows, x_out, group=None):
    """All-gather of unequal slices: rank r contributes rows[r][1]-rows[r][0] values of y_local,
    which land at x_out[rows[r][0]:rows[r][1]] on every rank.  Uses one broadcast per rank inside
    a single coalesced group (what SURVEY.md 8e names as the baseline collective); works for any
    backend (gloo on CPU tensors, NCCL on CUDA tensors)."""
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    r0, r1 = rows[me]
    x_out[r0:r1].copy_(y_local[: r1 - r0])
    works = []
    for r in range(world):
        a, b = rows[r]
        if b > a:
            works.append(dist.broadcast(x_out[a:b], src=r, group=group, async_op=True))
    for w in works:
        w.wait()
    return x_out
