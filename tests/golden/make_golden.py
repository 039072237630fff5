"""Generates the committed golden fixtures from the UNMODIFIED reference CPU path (oracle/_ref).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each .npz holds the input CSR, every Tile_matrix array, ptroffset1/2, the schedule and y for the
reference driver's data (x[i] = i % 10, main.cu:93-97) as produced by Tile_create
(csr2tile.h:629) and tilespmv_cpu (tilespmv_cpu.h:3).  The fixtures travel to the GPU box, where
/root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle_py as O  # noqa: E402
from tests.cases import CASES, HYB_CASES, x_for  # noqa: E402

GOLDEN = ["seven_formats", "lap2d_64", "lap3d27_24", "banded_2k_real", "band_contig_8k", "rmat_10_real",
          "uniform_2k", "rmat_unsorted", "ragged_band", "ragged_seven", "empty_rows"]


# fixtures of the NON-DEFAULT HYB rule: produced by the reference built with csr2tile.h:308-317 un-commented
# (oracle/Makefile, target ref_hyb) into tests/golden/hyb/
GOLDEN_HYB = ["hyb_rich", "hyb_rich_real", "hyb_ragged", "rmat_10_real"]


def main():
    for precision in ("f64", "f32"):
        for name, variant in [(n, "ref") for n in GOLDEN] + [(n, "refhyb") for n in GOLDEN_HYB]:
            ref = O.Reference(precision, variant)
            if precision == "f32" and name not in ("seven_formats", "banded_2k_real", "ragged_seven", "hyb_rich"):
                continue
            m, n, rp, ci, v = (CASES if variant == "ref" else HYB_CASES)[name]()
            v = v.astype(ref.val_dtype)
            M = ref.tile_create(m, n, rp, ci, v)
            arrs = ref.arrays(M, m)
            x = x_for(n, 1, ref.val_dtype)
            y, p1, p2, sched = ref.tilespmv_cpu(M, m, n, rp, ci, v, x)
            out = {"in_shape": np.array([m, n]), "in_rowptr": rp, "in_colidx": ci, "in_val": v,
                   "x": x, "y": y, "ptroffset1": p1, "ptroffset2": p2,
                   "rowblkblock": np.array([sched[0]]), "blkcoostylerowidx": sched[1],
                   "blkcoostylerowidx_colstart": sched[2], "blkcoostylerowidx_colstop": sched[3]}
            out.update({"tm_" + k: a for k, a in arrs.items()})
            path = os.path.join(HERE, f"{name}_{precision}.npz") if variant == "ref" else \
                os.path.join(HERE, "hyb", f"{name}_{precision}.npz")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
