"""x <- A*x on one GPU: a Python loop of tilespmv_plan_spmv calls vs tilespmv_plan_iterate (one CUDA-graph launch).
python tools/iterate_probe.py [--grid 1024] [--iters 200]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tilespmv_b200 import api, generators as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1024)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
for name, case in (("lap2d %d^2" % a.grid, g.lap2d(a.grid)), ("lap2d 256^2", g.lap2d(256)), ("lap3d27 64^3", g.lap3d27(64))):
    m, n, rp, ci, v = case
    plan = api.Plan(api.DeviceTileMatrix.from_csr(m, n, rp, ci, v * 0.1))
    xa = torch.rand(n, dtype=torch.float64, device="cuda")
    xb = torch.empty_like(xa)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for mode in ("loop", "graph", "loop", "graph"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        if mode == "loop":
            for i in range(a.iters):
                src, dst = (xa, xb) if i % 2 == 0 else (xb, xa)
                plan.spmv(src.data_ptr(), dst.data_ptr())
        else:
            plan.iterate(xa.data_ptr(), xb.data_ptr(), a.iters)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) * 1e3 / a.iters, (time.perf_counter() - t0) * 1e6 / a.iters)
    print(f"{name}: per iteration, device / wall us: loop {res['loop'][0]:.1f} / {res['loop'][1]:.1f}   "
          f"graph {res['graph'][0]:.1f} / {res['graph'][1]:.1f}", flush=True)
