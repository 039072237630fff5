"""The C restatement against the committed golden fixtures (tests/golden/*.npz), which were
produced by the unmodified reference CPU path (tests/golden/make_golden.py).  Runs without
/root/reference, e.g. on the GPU box."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tests import golden_util as G


@pytest.mark.parametrize("path", G.golden_files(), ids=os.path.basename)
def test_oracle_reproduces_golden(path):
    d = G.load(path)
    ora = O.Oracle(d["precision"])
    m, n = (int(v) for v in d["in_shape"])
    M = ora.tile_create(m, n, d["in_rowptr"], d["in_colidx"], d["in_val"])
    G.assert_tile_arrays_equal(ora.arrays(M, m), G.tile_arrays(d), d["name"] + ":")
    y, p1, p2 = ora.tilespmv_cpu(M, m, n, d["x"])
    assert y.tobytes() == d["y"].tobytes()
    assert np.array_equal(p1, d["ptroffset1"]) and np.array_equal(p2, d["ptroffset2"])
    rbb, a, b, c = ora.schedule(M)
    assert rbb == int(d["rowblkblock"][0])
    assert np.array_equal(a, d["blkcoostylerowidx"])
    assert np.array_equal(b, d["blkcoostylerowidx_colstart"])
    assert np.array_equal(c, d["blkcoostylerowidx_colstop"])
    # the tile SpMV agrees with the plain CSR loop (main.cu:101-110) exactly on integer data
    yg = ora.csr_spmv(m, d["in_rowptr"], d["in_colidx"], d["in_val"], d["x"])
    if d["precision"] == "f64" and np.all(d["in_val"] == np.round(d["in_val"])):
        assert y.tobytes() == yg.tobytes()
    else:
        scale = ora.csr_abs_spmv(m, d["in_rowptr"], d["in_colidx"], d["in_val"], d["x"])
        tol = 1e-12 if d["precision"] == "f64" else 1e-5
        assert np.all(np.abs(y - yg) <= tol * np.maximum(scale, 1e-300))
    ora.tile_destroy(M)
