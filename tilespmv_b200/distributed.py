"""One-process-per-GPU driver of the row-block sharded SpMV.

Every rank holds a self-contained Tile_matrix + plan for its contiguous range of block rows
(tilespmv_b200/sharding.py picks the cuts) and a replicated x.  A single SpMV is communication
free.  For the repeated-SpMV loop x <- A*x two exchanges are offered:

  "nccl"   baseline: the y slices (unequal lengths) are all-gathered into the next x with one
           NCCL broadcast per rank in a coalesced group;
  "fused"  the SpMV kernel's own epilogue stores every y value into the next-x buffer of every peer
           through NVLink-mapped pointers (torch symmetric memory = CUDA VMM peer mappings handed to
           tilespmv_plan_set_peers), so the all-gather IS the kernel's store stream and overlaps the
           HBM-bound compute; one device-side barrier per iteration orders the double-buffered x.

PyTorch is plumbing here (device memory, streams, process groups); the arithmetic is the C-ABI library.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api, sharding


class ShardedSpMV:
    def __init__(self, rows, rank, colA, local_rowptr, local_colidx, local_val, plan_kwargs=None, group=None):
        """rows: list of (r0, r1) per rank (sharding.row_ranges); local_*: CSR of rows[rank]."""
        self.rows, self.rank, self.colA, self.group = rows, rank, colA, group
        self.r0, self.r1 = rows[rank]
        self.m_local = self.r1 - self.r0
        self.dm = api.DeviceTileMatrix.from_csr(self.m_local, colA, local_rowptr, local_colidx, local_val)
        self.plan = api.Plan(self.dm, **(plan_kwargs or {}))
        self.dtype = torch.float64 if self.dm.precision == api.F64 else torch.float32
        self._symm = None

    # ---- single SpMV: y_local = A[r0:r1, :] @ x, no communication ----
    def spmv(self, x, y_local, stream=None):
        s = (stream or torch.cuda.current_stream()).cuda_stream
        self.plan.spmv(x.data_ptr(), y_local.data_ptr(), s)

    # ---- repeated SpMV ----
    def _ensure_symmetric(self, n):
        if self._symm is None:
            import torch.distributed._symmetric_memory as symm_mem
            g = self.group or dist.group.WORLD
            bufs = [symm_mem.empty(n, dtype=self.dtype, device="cuda") for _ in range(2)]
            hdls = [symm_mem.rendezvous(b, g) for b in bufs]
            self._symm = (bufs, hdls)
        return self._symm

    def iterate(self, x0, iters, mode="nccl", scale=1.0):
        """x_{k+1} = A @ x_k for `iters` steps (square A); returns the final replicated x.
        `scale` is unused by the kernels (plain x <- A*x like SURVEY.md 8d config 5)."""
        n = self.colA
        world = dist.get_world_size(self.group)
        stream = torch.cuda.current_stream().cuda_stream
        if mode == "nccl":
            xs = [x0.clone(), torch.empty_like(x0)]
            y = torch.empty(max(self.m_local, 1), dtype=self.dtype, device="cuda")
            for i in range(iters):
                src, dst = xs[i & 1], xs[(i + 1) & 1]
                self.plan.spmv(src.data_ptr(), y.data_ptr(), stream)
                sharding.allgather_rows(dist, y, self.rows, dst, self.group)
            self.plan.set_peers([], 0)
            return xs[iters & 1]
        if mode != "fused":
            raise ValueError(mode)
        bufs, hdls = self._ensure_symmetric(n)
        bufs[0].copy_(x0)
        hdls[0].barrier(channel=0)
        esz = bufs[0].element_size()
        for i in range(iters):
            src, dst = i & 1, (i + 1) & 1
            peers = [int(hdls[dst].buffer_ptrs[r]) for r in range(world) if r != self.rank]
            self.plan.set_peers(peers, self.r0)
            # the local slice of the next x is this rank's y: the kernel writes it in place
            self.plan.spmv(bufs[src].data_ptr(), bufs[dst].data_ptr() + self.r0 * esz, stream)
            hdls[dst].barrier(channel=0)  # everybody's stores into everybody's next x are done
        self.plan.set_peers([], 0)
        return bufs[iters & 1]


def build_sharded(rowA, colA, rowptr, colidx, val, plan_kwargs=None, group=None):
    """Convenience for matrices that fit on the host of every rank: partition by streamed bytes and
    build this rank's shard."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    vs = 8 if np.asarray(val).dtype == np.float64 else 4
    w = sharding.block_row_weights(rowptr, rowA, vs)
    parts = sharding.partition(w, world)
    rows = sharding.row_ranges(parts, rowA)
    r0, r1 = rows[rank]
    lrp, lci, lv = sharding.shard_csr(rowptr, colidx, val, r0, r1)
    return ShardedSpMV(rows, rank, colA, lrp, lci, lv, plan_kwargs, group)
