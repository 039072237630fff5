"""BASELINE configs 3 / 4 / 5 at FULL size on one B200, y against the plain CSR loop (main.cu:101-110 semantics; the
reference's O(tilem * tilen) conversion cannot restate these in reasonable time).  Together they take
~35 s on a B200 box (7 s + 25 s + 2 s, profiles/r02b_full_size.log), so they are part of the default `-m gpu` run whenever
the GPU has >= 60 GB and the host >= 32 GB of free memory; TILESPMV_SKIP_FULL_SIZE=1 leaves them out.

(bench.py --gpus N verifies configs 3 / 5 inside every multi-GPU run as well: per-rank y against torch's CSR SpMV and K
iterations of x <- A*x against torch's loop.)"""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tilespmv_b200 import api, generators as g


def _room():
    """(ok, why): enough device and host memory for the full-size inputs (config 4: ~5 GB of host CSR, ~9 GB on the device)."""
    if os.environ.get("TILESPMV_SKIP_FULL_SIZE") == "1":
        return False, "TILESPMV_SKIP_FULL_SIZE=1"
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "needs a GPU"
        if torch.cuda.get_device_properties(0).total_memory < 60 << 30:
            return False, "needs a GPU with >= 60 GB"
        with open("/proc/meminfo") as f:
            avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
        if avail_kb < 32 << 20:
            return False, "needs >= 32 GB of free host memory"
    except Exception as e:  # no torch / no /proc: leave the big cases out rather than fail
        return False, f"cannot size the box: {e}"
    return True, ""


_OK, _WHY = _room()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _OK, reason=_WHY or "full-size cases")]

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
