"""Small driver for profiling: builds one synthetic matrix, converts + plans on the GPU and runs a
few SpMVs.  python tools/spmv_run.py --workload lap3d27 --grid 160 --iters 5"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tilespmv_b200 import api, generators as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lap3d27")
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--chunk-bytes", type=int, default=0)
    ap.add_argument("--xstage-bytes", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--max-warps", type=int, default=0)
    ap.add_argument("--no-csr-groups", action="store_true")
    ap.add_argument("--hb", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--per-row", type=int, default=20)
    ap.add_argument("--xpanel-bytes", type=int, default=0)
    ap.add_argument("--check", action="store_true", help="compare y with torch's CSR SpMV (cuSPARSE) on the same device data")
    a = ap.parse_args()
    t0 = time.time()
    if a.workload == "lap3d27":
        m, n, rp, ci, v = g.lap3d27(a.grid)
    elif a.workload == "lap2d":
        m, n, rp, ci, v = g.lap2d(a.grid)
    elif a.workload == "banded":
        m, n, rp, ci, v = g.banded(a.n)
    elif a.workload == "band_contig":
        m, n, rp, ci, v = g.band_contig(a.n)
    elif a.workload == "rmat":
        m, n, rp, ci, v = g.rmat(a.scale)
    elif a.workload == "uniform":
        m, n, rp, ci, v = g.uniform_rows(a.n, 0, a.rows or a.n, a.per_row)
    else:
        raise SystemExit("unknown workload")
    dt = np.float64 if a.precision == "f64" else np.float32
    v = v.astype(dt)
    t_gen = time.time() - t0
    t0 = time.time()
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    torch.cuda.synchronize()
    t_conv = time.time() - t0
    t0 = time.time()
    plan = api.Plan(dm, a.chunk_bytes, a.xstage_bytes, a.ctas_per_sm, a.stages, a.max_warps, csr_groups=not a.no_csr_groups, xpanel_bytes=a.xpanel_bytes)
    torch.cuda.synchronize()
    t_plan = time.time() - t0
    pi, di = plan.info(), dm.info()
    tdt = torch.float64 if a.precision == "f64" else torch.float32
    x = torch.rand(n, dtype=tdt, device="cuda") * 2 - 1
    y = torch.empty(m, dtype=tdt, device="cuda")
    ms = plan.time(x.data_ptr(), y.data_ptr(), a.warmup, a.iters)
    nnz = int(rp[m])
    if a.check:
        A = torch.sparse_csr_tensor(torch.from_numpy(rp).cuda().long(), torch.from_numpy(ci).cuda().long(),
                                    torch.from_numpy(v).cuda(), size=(m, n))
        y_ref = A @ x
        scale = torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), A.values().abs(), size=(m, n)) @ x.abs()
        tol = 1e-12 if a.precision == "f64" else 1e-5
        bad = int(((y - y_ref).abs() > tol * scale.clamp_min(1e-30)).sum())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            A @ x
        e0.record()
        for _ in range(a.iters):
            A @ x
        e1.record()
        torch.cuda.synchronize()
        print(f"check vs torch CSR SpMV: {bad} rows out of tolerance {tol} (relative to sum |a||x|); "
              f"torch/cuSPARSE CSR SpMV on the same data: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us", flush=True)
    print(f"{a.workload} m={m} nnz={nnz} tiles={di.tilenum} fmt={list(di.tiles_by_format)} side={di.nnz_side} "
          f"chunks={pi.nchunks} split={pi.split_rows} groups={pi.csr_groups} stream={pi.stream_bytes} B_alg={pi.algorithmic_bytes} "
          f"grid={pi.grid} smem={pi.smem_bytes} | gen {t_gen:.2f}s conv {t_conv:.3f}s plan {t_plan:.3f}s | "
          f"{ms * 1e3:.1f} us/SpMV {2 * nnz / ms / 1e6:.1f} GFLOP/s {pi.algorithmic_bytes / ms / 1e6:.0f} GB/s(alg) "
          f"{pi.stream_bytes / ms / 1e6:.0f} GB/s(stream)", flush=True)


if __name__ == "__main__":
    main()
