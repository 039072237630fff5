"""First-contact diagnostic for a GPU box: runs conversion + SpMV on a few cases and prints where
things differ instead of stopping at the first assert.  python -m tests.gpu_diag"""
import sys
import traceback

import numpy as np

from oracle import oracle_py as O
from tests.cases import CASES, x_for
from tilespmv_b200 import api


def main(names):
    for name in names:
        m, n, rp, ci, v = CASES[name]()
        ora = O.Oracle("f64")
        Mo = ora.tile_create(m, n, rp, ci, v)
        want = ora.arrays(Mo, m)
        print(f"== {name}: m={m} n={n} nnz={len(ci)} tiles={Mo.tilenum} fmt_hist="
              f"{[int((want['Format'] == f).sum()) for f in range(7)]}", flush=True)
        try:
            dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
            got = dm.export().arrays()
            nbad = 0
            for k in want:
                if got[k].shape != want[k].shape:
                    print(f"   {k}: SHAPE {got[k].shape} vs {want[k].shape}")
                    nbad += 1
                elif got[k].tobytes() != want[k].tobytes():
                    bad = np.flatnonzero(got[k] != want[k])
                    print(f"   {k}: {len(bad)} mismatches first {bad[:6]} got {got[k][bad[:6]]} want {want[k][bad[:6]]}")
                    nbad += 1
            print(f"   conversion: {'OK' if nbad == 0 else str(nbad) + ' arrays differ'}", flush=True)
        except Exception:
            traceback.print_exc()
            continue
        try:
            for src, label in ((dm, "gpu-converted"),):
                plan = api.Plan(src)
                pi = plan.info()
                print(f"   plan[{label}]: chunks={pi.nchunks} stream={pi.stream_bytes} B_alg={pi.algorithmic_bytes} "
                      f"split={pi.split_rows} grid={pi.grid} smem={pi.smem_bytes}", flush=True)
                for mode in (1, 0):
                    x = x_for(n, mode)
                    y_ref, _, _ = ora.tilespmv_cpu(Mo, m, n, x)
                    y = plan.spmv_host(x)
                    scale = ora.csr_abs_spmv(m, rp, ci, v, x)
                    err = np.abs(y - y_ref)
                    bad = np.flatnonzero(err > 1e-12 * np.maximum(scale, 1e-300))
                    print(f"   spmv mode {mode}: max err {err.max(initial=0):.3e}, bad rows {len(bad)} {bad[:10]}", flush=True)
                    if len(bad):
                        print("      got ", y[bad[:6]], "\n      want", y_ref[bad[:6]])
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main(sys.argv[1:] or ["seven_formats", "lap2d_64", "lap3d27_24", "banded_8k", "band_contig_8k", "rmat_12",
                          "uniform_8k", "ragged_seven", "ragged_band", "empty_rows", "empty_matrix", "band_unsorted"])
