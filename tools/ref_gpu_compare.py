"""Reference GPU kernel (stir_spmv_cuda_kernel_v6, built unmodified for sm_100 by `make -C oracle ref_gpu`)
vs this library on the SAME box and the SAME .mtx inputs.
python tools/ref_gpu_compare.py [--cases lap2d:1024,lap3d27:96]   -> one JSON line per case"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tilespmv_b200 import api, generators as g  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_test_sm100")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="lap2d:1024,lap3d27:96")
    ap.add_argument("--tmp", default="/tmp")
    a = ap.parse_args()
    for case in a.cases.split(","):
        kind, size = case.split(":")
        size = int(size)
        m, n, rp, ci, v = getattr(g, kind)(size, val_mode=1)
        path = os.path.join(a.tmp, f"{kind}_{size}.mtx")
        t0 = time.time()
        g.write_mtx_fast(path, m, n, rp, ci, v)
        t_write = time.time() - t0
        out = {"case": case, "m": m, "nnz": int(rp[m]), "mtx_write_s": round(t_write, 2)}
        # ---- reference: ./test -d 0 file.mtx (its own protocol: 200 warm-up + 4x1000 untimed + 1000 timed) ----
        if os.path.exists(REF_BIN):
            t0 = time.time()
            r = subprocess.run(["stdbuf", "-o0", "-e0", REF_BIN, "-d", "0", path], capture_output=True, text=True, cwd=a.tmp, timeout=1500)
            out["ref_rc"] = r.returncode
            out["ref_wall_s"] = round(time.time() - t0, 1)
            mt = re.search(r"CUDA SpMV runtime\s+([0-9.]+) ms,\s+([0-9.]+) GFlops", r.stdout)
            out["ref_check"] = "PASS" if "Check... PASS" in r.stdout else ("NO PASS" if "NO PASS" in r.stdout else "?")
            if mt:
                out["ref_ms"], out["ref_gflops"] = float(mt.group(1)), float(mt.group(2))
            else:
                out["ref_stdout_tail"] = r.stdout[-300:] + r.stderr[-300:]
        else:
            out["ref"] = "oracle/_ref/ref_test_sm100 not built"
        # ---- this library on the same CSR (values i%10, x = i%10 like main.cu:68-69, 93-97) ----
        m16 = (m // 16) * 16
        dm = api.DeviceTileMatrix.from_csr(m16, n, rp[: m16 + 1], ci, v)
        plan = api.Plan(dm)
        x = torch.from_numpy((np.arange(n) % 10).astype(np.float64)).cuda()
        y = torch.empty(m16, dtype=torch.float64, device="cuda")
        ms = min(plan.time(x.data_ptr(), y.data_ptr(), 200, 1000) for _ in range(2))
        out["ours_ms"] = ms
        out["ours_gflops"] = 2.0 * int(rp[m]) / ms / 1e6
        if "ref_ms" in out:
            out["speedup_vs_reference_gpu_kernel"] = out["ref_ms"] / ms
        print(json.dumps(out), flush=True)
        os.remove(path)


if __name__ == "__main__":
    main()
