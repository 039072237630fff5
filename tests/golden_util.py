import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(precision=None, hyb=False):
    """hyb=True: the fixtures of the non-default HYB rule (tests/golden/hyb/)."""
    pat = f"*_{precision}.npz" if precision else "*.npz"
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "hyb" if hyb else "", pat)))


def load(path):
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    precision = name[-3:]
    d = {k: z[k] for k in z.files}
    d["name"], d["precision"] = name, precision
    return d


def tile_arrays(d):
    return {k[3:]: v for k, v in d.items() if k.startswith("tm_")}


def assert_tile_arrays_equal(got, want, what=""):
    assert got.keys() == want.keys(), (sorted(got), sorted(want))
    for k in want:
        assert got[k].shape == want[k].shape, f"{what}{k}: shape {got[k].shape} vs {want[k].shape}"
        if got[k].tobytes() != want[k].tobytes():
            bad = np.flatnonzero(got[k] != want[k])
            raise AssertionError(f"{what}{k}: {len(bad)} mismatches, first at {bad[:8]}: "
                                 f"got {got[k][bad[:8]]} want {want[k][bad[:8]]}")
