"""A/B timing of plan configurations on ONE box, interleaved over several rounds.
python tools/ab.py --grid 128 --configs "2:0,2:20,2:18,3:0" [--rounds 3] [--iters 100]
config = stages:max_warps[:chunk_bytes[:xstage[:nogroups]]] (nogroups = 1 keeps CSR tiles individual)"""
import argparse
import os
import sys
import subprocess

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tilespmv_b200 import api, generators as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lap3d27")
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--configs", default="2:0")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--iters", type=int, default=100)
    a = ap.parse_args()
    gen = {"lap3d27": lambda: g.lap3d27(a.grid), "lap2d": lambda: g.lap2d(a.grid), "banded": lambda: g.banded(a.n),
           "band_contig": lambda: g.band_contig(a.n), "rmat": lambda: g.rmat(a.scale), "uniform": lambda: g.uniform(a.n)}
    m, n, rp, ci, v = gen[a.workload]()
    dt = np.float64 if a.precision == "f64" else np.float32
    tdt = torch.float64 if a.precision == "f64" else torch.float32
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v.astype(dt))
    x = torch.rand(n, dtype=tdt, device="cuda") * 2 - 1
    y = torch.empty(m, dtype=tdt, device="cuda")
    plans = []
    for c in a.configs.split(","):
        f = [int(t) for t in c.split(":")] + [0, 0, 0, 0]
        p = api.Plan(dm, chunk_bytes=f[2], xstage_bytes=f[3], stages=f[0], max_warps=f[1], csr_groups=not f[4])
        plans.append((c, p))
    try:
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
    except Exception:
        clk = "?"
    res = {c: [] for c, _ in plans}
    for r in range(a.rounds):
        for c, p in plans:
            res[c].append(p.time(x.data_ptr(), y.data_ptr(), 5, a.iters) * 1e3)
    nnz = int(rp[m])
    print(f"{a.workload} grid={a.grid} m={m} nnz={nnz} idle clocks: {clk}")
    for c, p in plans:
        i = p.info()
        t = min(res[c])
        print(f"  cfg {c:12s} block={i.block:4d} smem={i.smem_bytes:6d} chunks={i.nchunks} groups={i.csr_groups} stream={i.stream_bytes} "
              f"us={' '.join(f'{u:.1f}' for u in res[c])}  best {t:.1f} us  {i.algorithmic_bytes / t / 1e3:.0f} GB/s(alg) "
              f"{2 * nnz / t / 1e3:.0f} GFLOP/s", flush=True)


if __name__ == "__main__":
    main()
