"""BASELINE configs 3 / 5 (and the weak-scaling stencil) row-block sharded over the GPUs of one box.
Launch with torchrun (one rank per GPU) or plain python for 1 GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/scale_configs.py --config c5 [--size 50000000] [--iters 20]

Prints one JSON line (rank 0): single-SpMV GFLOP/s (no communication), and the repeated-SpMV loop
x <- A*x with the per-iteration all-gather of x, NCCL baseline and fused NVLink epilogue.
Times are CUDA events on the launching stream, max over ranks."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tilespmv_b200 import distributed as D, generators as g, sharding as sh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=["c3", "c5"])
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--sim-world", type=int, default=0, help="on ONE GPU: time only the shard rank 0 of this many ranks would own (c5)")
    ap.add_argument("--chunk-bytes", type=int, default=0)
    ap.add_argument("--xpanel-bytes", type=int, default=0)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if "MASTER_ADDR" not in os.environ:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29541"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dt = np.float64 if a.precision == "f64" else np.float32
    tdt = torch.float64 if a.precision == "f64" else torch.float32
    vs = 8 if a.precision == "f64" else 4
    t0 = time.time()
    if a.config == "c5":  # uniform random, 20 distinct columns per row: every rank generates its own rows
        N = a.size or 50_000_000
        W = a.sim_world if (a.sim_world and world == 1) else world
        per = (N // 16 // W) * 16
        rows = [(r * per, (r + 1) * per if r < W - 1 else N) for r in range(W)]
        r0, r1 = rows[rank]
        if W != world:
            rows = [rows[0]]
        m, n, rp, ci, v = g.uniform_rows(N, r0, r1 - r0)
        name = f"uniform random {N}x{N}, 20 nnz/row (config 5)"
    else:  # banded: generated whole on every rank (counter-based RNG), cut by streamed bytes
        N = a.size or 8_000_000
        M, n, grp, gci, gv = g.banded(N)
        w = sh.block_row_weights(grp, M, vs)
        rows = sh.row_ranges(sh.partition(w, world), M)
        r0, r1 = rows[rank]
        rp, ci, v = sh.shard_csr(grp, gci, gv, r0, r1)
        del grp, gci, gv
        name = f"banded {N}x{N}, half-bandwidth 64, 37 nnz/row (config 3)"
    t_gen = time.time() - t0
    nnz_local = int(rp[r1 - r0])
    t0 = time.time()
    sp = D.ShardedSpMV(rows, rank, n, rp, ci, v.astype(dt), plan_kwargs=dict(chunk_bytes=a.chunk_bytes, xpanel_bytes=a.xpanel_bytes))
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    del rp, ci, v
    nnz_t = torch.tensor([nnz_local], device="cuda", dtype=torch.int64)
    dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    gen = torch.Generator(device="cuda").manual_seed(99)
    x0 = (torch.rand(n, dtype=tdt, device="cuda", generator=gen) * 2 - 1) / 64.0
    y = torch.empty(max(sp.m_local, 1), dtype=tdt, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_spmv = timed(lambda: sp.spmv(x0, y), a.iters)
    pi = sp.plan.info()
    out = {"config": name, "n_gpus": world, "precision": a.precision, "nnz": nnz, "rows": n,
           "spmv_ms": ms_spmv, "spmv_gflops": 2.0 * nnz / ms_spmv / 1e6,
           "rank0": {"rows": sp.m_local, "nnz": nnz_local, "chunks": pi.nchunks, "split_rows": pi.split_rows,
                     "stream_bytes": pi.stream_bytes, "B_alg": pi.algorithmic_bytes, "block": pi.block,
                     "xpanels": pi.xpanels, "launches_per_spmv": pi.launches_per_spmv, "chunk_bytes": pi.chunk_bytes,
                     "GBps_alg": pi.algorithmic_bytes / ms_spmv / 1e6},
           "gen_s": round(t_gen, 1), "convert_plan_s": round(t_setup, 1)}
    if world > 1:
        for mode in ("nccl", "fused"):
            try:
                ms = timed(lambda: sp.iterate(x0, 4, mode=mode), max(1, a.iters // 4)) / 4.0
                out[f"iterate_{mode}_ms"] = ms
                out[f"iterate_{mode}_gflops"] = 2.0 * nnz / ms / 1e6
            except Exception as e:
                out[f"iterate_{mode}_error"] = str(e)[:200]
        out["allgather_bytes_in_per_gpu"] = (n - sp.m_local) * vs
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
