"""Multi-rank host logic on CPU: byte-balanced block-row partition (tilespmv_b200/sharding.py) and the
repeated-SpMV loop with the all-gather of unequal y slices, world_size 2 over gloo.  The per-shard
arithmetic is done by the oracle here (no GPU in this suite); what is under test is that sharding +
exchange reproduce the single-process result of the reference CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_py as O
from tilespmv_b200 import generators as g, sharding as sh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_tiles_the_block_rows_and_balances_bytes():
    m, n, rp, ci, v = g.rmat(12, val_mode=1)  # skewed rows
    w = sh.block_row_weights(rp, m, 8)
    for nranks in (1, 2, 3, 4, 8):
        parts = sh.partition(w, nranks)
        assert len(parts) == nranks and parts[0][0] == 0 and parts[-1][1] == len(w)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(nranks - 1))
        assert all(b1 >= b0 for b0, b1 in parts)
        if nranks <= 4:
            # no shard is heavier than the mean by more than the heaviest single block row
            s = np.array([w[b0:b1].sum() for b0, b1 in parts])
            assert s.max() <= s.mean() + w.max() + 1e-9
    rows = sh.row_ranges(sh.partition(w, 3), m)
    assert rows[0][0] == 0 and rows[-1][1] == m and all(r0 % 16 == 0 for r0, _ in rows)


def test_native_partition_equals_the_python_one():
    """tilespmv_partition_rows (C-ABI, host-only) returns the cuts of sharding.partition + row_ranges."""
    cases = [g.rmat(12, val_mode=1), g.banded(4096 + 16 * 3, val_mode=0), g.lap3d27(24, val_mode=1), g.uniform(2048, val_mode=1),
             (20, 20, np.array([0, 3, 3, 10] + [10] * 17, np.int32), None, None),
             (0, 0, np.zeros(1, np.int32), None, None)]
    for m, n, rp, ci, v in cases:
        for vs in (8, 4):
            for nranks in (1, 2, 3, 4, 8):
                want = sh.row_ranges(sh.partition(sh.block_row_weights(rp, m, vs), nranks), m) if m else [(0, 0)] * nranks
                assert sh.partition_rows(rp, m, nranks, vs) == want, (m, vs, nranks)


def test_partition_more_ranks_than_block_rows_and_ragged_tail():
    w = sh.block_row_weights(np.array([0, 3, 3, 10] + [10] * 17, np.int64), 20, 8)  # 20 rows -> 2 block rows
    parts = sh.partition(w, 4)
    assert parts[0][0] == 0 and parts[-1][1] == 2
    rows = sh.row_ranges(parts, 20)
    assert rows[-1][1] == 20 and sum(r1 - r0 for r0, r1 in rows) == 20


def _worker(rank, world, port, K, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, n, rp, ci, v = g.banded(4096 + 16 * 3, val_mode=0)  # square, 259 block rows
        w = sh.block_row_weights(rp, m, 8)
        parts = sh.partition(w, world)
        rows = sh.row_ranges(parts, m)
        r0, r1 = rows[rank]
        lrp, lci, lv = sh.shard_csr(rp, ci, v, r0, r1)
        ora = O.Oracle("f64")
        M = ora.tile_create(r1 - r0, n, lrp, lci, lv)  # a self-contained Tile_matrix per shard
        x = torch.from_numpy(np.random.default_rng(11).uniform(-1, 1, n))
        x_next = torch.zeros_like(x)
        for _ in range(K):
            y_local, _, _ = ora.tilespmv_cpu(M, r1 - r0, n, x.numpy())
            y_local = torch.from_numpy(y_local / 16.0)  # keep the iterates bounded
            sh.allgather_rows(dist, y_local, rows, x_next)
            x, x_next = x_next, x
        ora.tile_destroy(M)
        if rank == 0:
            out_q.put((x.numpy().copy(), [r1 - r0 for r0, r1 in rows]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_repeated_spmv_over_gloo_matches_single_process():
    world, K = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    x_dist, counts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(counts) == 4096 + 48 and counts[0] != 0 and counts[1] != 0

    # single-process reference: the same loop through the oracle on the whole matrix
    m, n, rp, ci, v = g.banded(4096 + 16 * 3, val_mode=0)
    ora = O.Oracle("f64")
    M = ora.tile_create(m, n, rp, ci, v)
    x = np.random.default_rng(11).uniform(-1, 1, n)
    for _ in range(K):
        y, _, _ = ora.tilespmv_cpu(M, m, n, x)
        x = y / 16.0
    ora.tile_destroy(M)
    # shards cut at block-row boundaries keep every tile intact => the per-row sums are identical
    assert np.array_equal(x_dist, x)
