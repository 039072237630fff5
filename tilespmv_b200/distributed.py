"""Thin ctypes caller of the library's multi-GPU entry points (include/tilespmv.h: tilespmv_comm_* / tilespmv_dist_*).

One process per GPU.  Every rank converts its contiguous range of block rows (tilespmv_b200/sharding.py or
tilespmv_partition_rows pick the cuts), the library owns everything else: the rendezvous (POSIX shared memory), the
CUDA-IPC-mapped x buffers, the NCCL communicator and the three exchanges of the repeated-SpMV loop x <- A*x

  "nccl"       SpMV, then ncclAllGather / grouped ncclBroadcast            (baseline)
  "fused"      the kernel's epilogue stores y into every peer's next x      (NVLink P2P stores)
  "pipelined"  copy-engine pushes ordered by need + per-launch waits        (exchange hidden under the next iteration)
  "halo"       fused stores of the rows a peer reads next + copy-engine replication of the rest (bands, stencils)

Nothing here needs torch: device pointers are plain ints.  torch shows up only in callers that want tensors.
"""
import ctypes as C
import os

import numpy as np

from . import _capi, api, sharding

EXCHANGES = {"nccl": _capi.EXCHANGE_NCCL, "fused": _capi.EXCHANGE_FUSED, "pipelined": _capi.EXCHANGE_PIPELINED,
             "halo": _capi.EXCHANGE_HALO}


class Comm:
    """tilespmv_comm: rendezvous of `nranks` processes of one box under a job-unique `name`."""

    def __init__(self, name, rank, nranks, nccl=True):
        L = _capi.load()
        h = C.c_void_p()
        _capi.check(L.tilespmv_comm_create(name.encode(), rank, nranks, _capi.COMM_NCCL if nccl else 0, C.byref(h)),
                    "tilespmv_comm_create")
        self.handle, self.rank, self.nranks, self.has_nccl = h, rank, nranks, nccl

    def barrier(self):
        _capi.check(_capi.load().tilespmv_comm_barrier(self.handle), "tilespmv_comm_barrier")

    def destroy(self):
        if self.handle:
            _capi.load().tilespmv_comm_destroy(self.handle)
            self.handle = None


def job_name(default="job"):
    """A rendezvous name all ranks of a torchrun / mp.spawn job agree on and no other job shares."""
    return f"{default}_{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', os.getppid())}"


class ShardedSpMV:
    """tilespmv_dist: this rank's row block of a square matrix + the replicated x."""

    def __init__(self, comm, rows, colA, local_rowptr, local_colidx, local_val, plan_kwargs=None, uniform_panels=False):
        """rows: list of (r0, r1) per rank (sharding.row_ranges); local_*: CSR of rows[comm.rank], global columns."""
        L = _capi.load()
        self.comm, self.rows, self.rank, self.colA = comm, rows, comm.rank, colA
        self.r0, self.r1 = rows[comm.rank]
        self.m_local = self.r1 - self.r0
        self.dm = api.DeviceTileMatrix.from_csr(self.m_local, colA, local_rowptr, local_colidx, local_val)
        self.val_dtype = self.dm.val_dtype
        kw = dict(plan_kwargs or {})
        opts = _capi.PlanOptions(kw.get("chunk_bytes", 0), kw.get("xstage_bytes", 0), kw.get("ctas_per_sm", 0), kw.get("stages", 0),
                                 kw.get("max_warps", 0),
                                 (0 if kw.get("csr_groups", True) else _capi.PLAN_NO_CSR_GROUPS) |
                                 (0 if kw.get("flat_side", True) else _capi.PLAN_NO_FLAT_SIDE), kw.get("xpanel_bytes", 0))
        cuts = (C.c_int64 * (comm.nranks + 1))(*([r[0] for r in rows] + [rows[-1][1]]))
        h = C.c_void_p()
        _capi.check(L.tilespmv_dist_create(comm.handle, self.dm.handle, cuts, C.byref(opts),
                                           _capi.DIST_UNIFORM_PANELS if uniform_panels else 0, C.byref(h)), "tilespmv_dist_create")
        self.handle = h
        # the rank's plan, borrowed from the dist object (never destroyed from here)
        self.plan = api.Plan.__new__(api.Plan)
        self.plan.handle = C.c_void_p(L.tilespmv_dist_plan(h))
        self.plan.precision, self.plan.rowA, self.plan.colA, self.plan.val_dtype = self.dm.precision, self.m_local, colA, self.val_dtype
        self.plan.destroy = lambda: None

    def info(self):
        i = _capi.DistInfo()
        _capi.check(_capi.load().tilespmv_dist_get_info(self.handle, C.byref(i)), "tilespmv_dist_get_info")
        return i

    # ---- single SpMV: y_local = A[r0:r1, :] @ x, no communication (raw device pointers) ----
    def spmv(self, d_x, d_y_local, stream=0):
        self.plan.spmv(d_x, d_y_local, stream)

    # ---- repeated SpMV ----
    def iterate(self, d_x0, iters, mode="pipelined", stream=0):
        """x_{k+1} = A @ x_k for `iters` steps, asynchronous on `stream`; d_x0 = device pointer of the replicated start
        vector (or 0 / None to continue).  Returns the device pointer of the replicated result (library-owned)."""
        L = _capi.load()
        _capi.check(L.tilespmv_dist_iterate(self.handle, C.c_void_p(d_x0 or None), iters, EXCHANGES[mode], C.c_void_p(stream)),
                    "tilespmv_dist_iterate")
        return L.tilespmv_dist_x(self.handle)

    def sync(self, stream=0):
        _capi.check(_capi.load().tilespmv_dist_sync(self.handle, C.c_void_p(stream)), "tilespmv_dist_sync")

    def destroy(self):
        if self.handle:
            _capi.load().tilespmv_dist_destroy(self.handle)
            self.handle = None
            self.plan.handle = None
            self.dm.destroy()


def build_sharded(comm, rowA, colA, rowptr, colidx, val, plan_kwargs=None, uniform_panels=False):
    """Convenience for matrices that fit on the host of every rank: partition by streamed bytes and build this rank's shard."""
    vs = 8 if np.asarray(val).dtype == np.float64 else 4
    rows = sharding.partition_rows(rowptr, rowA, comm.nranks, vs)
    r0, r1 = rows[comm.rank]
    lrp, lci, lv = sharding.shard_csr(rowptr, colidx, val, r0, r1)
    return ShardedSpMV(comm, rows, colA, lrp, lci, lv, plan_kwargs, uniform_panels)
