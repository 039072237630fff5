"""Row-block sharded SpMV on >= 2 GPUs (NCCL + the fused NVLink epilogue), checked against the oracle.
Skipped on boxes with a single GPU; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, K, q):
    import torch.distributed as dist
    from tilespmv_b200 import distributed as D, generators as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m, n, rp, ci, v = getattr(g, case[0])(*case[1], val_mode=0)
        # case[2]: plan options, e.g. x panels (column-panel sub-plans; the fused peer stores ride on the last launch)
        sp = D.build_sharded(m, n, rp, ci, v, plan_kwargs=case[2] if len(case) > 2 else None)
        x0 = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, n)).cuda() / 64.0
        # single SpMV, no communication
        y = torch.empty(max(sp.m_local, 1), dtype=torch.float64, device="cuda")
        sp.spmv(x0, y)
        torch.cuda.synchronize()
        out = {"rows": sp.rows, "y": y[: sp.m_local].cpu().numpy()}
        for mode in ("nccl", "fused"):
            xk = sp.iterate(x0, K, mode=mode)
            torch.cuda.synchronize()
            out[mode] = xk.cpu().numpy().copy()
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case", [("banded", (65536,)), ("rmat", (13,)), ("lap3d27", (32,)),
                                  ("rmat", (13,), dict(xpanel_bytes=8192)), ("uniform", (16384,), dict(xpanel_bytes=16384))],
                         ids=lambda c: c[0] + ("_xpanels" if len(c) > 2 else ""))
def test_sharded_spmv_and_repeated_spmv(case):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle_py as O
    from tilespmv_b200 import generators as g
    world, K = min(torch.cuda.device_count(), 4), 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    m, n, rp, ci, v = getattr(g, case[0])(*case[1], val_mode=0)
    ora = O.Oracle("f64")
    x0 = np.random.default_rng(5).uniform(-1, 1, n) / 64.0
    y_ref = ora.csr_spmv(m, rp, ci, v, x0)
    scale = ora.csr_abs_spmv(m, rp, ci, v, np.abs(x0))
    rows = res[0]["rows"]
    for r in range(world):
        r0, r1 = rows[r]
        assert np.all(np.abs(res[r]["y"] - y_ref[r0:r1]) <= 1e-12 * np.maximum(scale[r0:r1], 1e-300)), f"rank {r} y"
    # repeated SpMV: every rank ends with the same replicated x, equal to the CPU loop
    x = x0.copy()
    bound = np.abs(x0)
    for _ in range(K):
        bound = ora.csr_abs_spmv(m, rp, ci, v, bound)
        x = ora.csr_spmv(m, rp, ci, v, x)
    for mode in ("nccl", "fused"):
        for r in range(world):
            assert np.all(np.abs(res[r][mode] - x) <= 1e-11 * np.maximum(bound, 1e-300)), f"{mode} rank {r}"
        assert all(np.array_equal(res[r][mode], res[0][mode]) for r in range(world)), f"{mode}: x differs between ranks"
