"""BASELINE configs 3 / 4 / 5 at FULL size on one B200, y against the plain CSR loop (main.cu:101-110 semantics; the
reference's O(tilem * tilen) conversion cannot restate these in reasonable time).  Minutes of host-side
generation each, so they only run with TILESPMV_FULL_SIZE=1:

    TILESPMV_FULL_SIZE=1 python -m pytest tests/test_full_size.py -m gpu -q

(bench.py --gpus N verifies configs 3 / 5 inside every multi-GPU run as well: per-rank y against torch's CSR SpMV and K
iterations of x <- A*x against torch's loop.)"""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tilespmv_b200 import api, generators as g

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("TILESPMV_FULL_SIZE") != "1", reason="set TILESPMV_FULL_SIZE=1")]

TOL = {"f64": 1e-12, "f32": 1e-5}


def _check(case, precision):
    m, n, rp, ci, v = case
    dt = np.float64 if precision == "f64" else np.float32
    v = v.astype(dt)
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm)
    x = np.random.default_rng(11).uniform(-1, 1, n).astype(dt)
    y = plan.spmv_host(x)
    ora = O.Oracle(precision)
    y_ref = ora.csr_spmv(m, rp, ci, v, x, parallel=True)
    scale = ora.csr_abs_spmv(m, rp, ci, v, x)
    err = np.abs(y.astype(np.float64) - y_ref.astype(np.float64))
    assert np.all(err <= TOL[precision] * np.maximum(scale.astype(np.float64), 1e-300))
    return dm.info(), plan.info()


@pytest.mark.timeout(1800)
def test_config3_banded_8m_full_size():
    di, pi = _check(g.banded(8_000_000, val_mode=0), "f64")
    assert di.nnz == 295_998_847 and di.tiles_by_format[0] > 0.99 * di.tilenum  # ~99.9 % CSR tiles


@pytest.mark.timeout(3600)
def test_config4_rmat_scale_24_fp32_full_size():
    di, pi = _check(g.rmat(24, val_mode=0), "f32")
    assert di.rowA == 1 << 24 and di.nnz_side > 0.9 * di.nnz and pi.split_rows > 0


@pytest.mark.timeout(1800)
def test_config5_one_of_eight_row_blocks_full_width():
    di, pi = _check(g.uniform_rows(50_000_000, 0, 6_250_000, val_mode=0), "f64")
    assert di.nnz == di.nnz_side == 125_000_000 and pi.xpanels > 1
