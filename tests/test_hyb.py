"""The NON-DEFAULT HYB rule.  Upstream ships the selection of format 3 commented out
(csr2tile.h:308-316) while the encoder (:505-548, :984-1008) and the CPU evaluation (tilespmv_cpu.h:190-223)
are live code.  `oracle/_ref/libtilespmv_refhyb_*.so` is the reference with that one rule un-commented
(oracle/Makefile, target ref_hyb); the fixtures in tests/golden/hyb/ were produced by it.

CPU tests: the restatement with enable_hyb against the refhyb build and against the fixtures; the default
never emits format 3.  GPU tests (-m gpu): TILESPMV_ENABLE_HYB conversion bit-exact, y through plan / kernel,
and a reference-built Tile_matrix that contains HYB tiles uploaded and multiplied.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tests import golden_util as G
from tests.cases import CASES, HYB_CASES, x_for

needs_refhyb = pytest.mark.skipif(not O.ref_available("f64", "refhyb"), reason="oracle/_ref refhyb variant not built")


@needs_refhyb
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(HYB_CASES))
def test_oracle_with_hyb_matches_the_reference_with_its_hyb_rule_enabled(name, precision):
    m, n, rp, ci, v = HYB_CASES[name]()
    ora, ref = O.Oracle(precision, enable_hyb=True), O.Reference(precision, "refhyb")
    v = v.astype(ora.val_dtype)
    Mo, Mr = ora.tile_create(m, n, rp, ci, v), ref.tile_create(m, n, rp, ci, v)
    G.assert_tile_arrays_equal(ora.arrays(Mo, m), ref.arrays(Mr, m), name + ":")
    for mode in (1, 0):
        x = x_for(n, mode, ora.val_dtype)
        yo, p1o, p2o = ora.tilespmv_cpu(Mo, m, n, x)
        yr, p1r, p2r, _ = ref.tilespmv_cpu(Mr, m, n, rp, ci, v, x)
        assert yo.tobytes() == yr.tobytes()
        assert np.array_equal(p1o, p1r) and np.array_equal(p2o, p2r)
    ora.tile_destroy(Mo)


@pytest.mark.parametrize("path", G.golden_files(hyb=True), ids=os.path.basename)
def test_oracle_with_hyb_reproduces_the_hyb_fixtures(path):
    d = G.load(path)
    ora = O.Oracle(d["precision"], enable_hyb=True)
    m, n = (int(v) for v in d["in_shape"])
    M = ora.tile_create(m, n, d["in_rowptr"], d["in_colidx"], d["in_val"])
    got = ora.arrays(M, m)
    G.assert_tile_arrays_equal(got, G.tile_arrays(d), d["name"] + ":")
    if d["name"].startswith("hyb_"):
        assert (got["Format"] == 3).sum() > 0
    y, p1, p2 = ora.tilespmv_cpu(M, m, n, d["x"])
    assert y.tobytes() == d["y"].tobytes()
    assert np.array_equal(p1, d["ptroffset1"]) and np.array_equal(p2, d["ptroffset2"])
    ora.tile_destroy(M)


def test_default_never_emits_hyb_and_spilled_entries_live_in_the_side_matrix():
    m, n, rp, ci, v = HYB_CASES["hyb_rich"]()
    Md = O.Oracle("f64").tile_create(m, n, rp, ci, v)
    ad = O.Oracle("f64").arrays(Md, m)
    assert (ad["Format"] == 3).sum() == 0 and Md.hybsize == 0
    ora = O.Oracle("f64", enable_hyb=True)
    Mh = ora.tile_create(m, n, rp, ci, v)
    ah = ora.arrays(Mh, m)
    hyb = ah["Format"] == 3
    assert hyb.sum() > 0 and np.all(ad["Format"][hyb] == 0)  # only would-be CSR tiles turn into HYB
    assert np.all(ah["Format"][~hyb] == ad["Format"][~hyb])
    spill = np.diff(ah["hyb_coocount"])
    assert np.all(spill[hyb] <= 4) and Mh.hybcoosize == spill.sum()
    # every spilled entry is ALSO in the side matrix (new_coocount, csr2tile.h:316, :538-545)
    assert Mh.coototal == Md.coototal + Mh.hybcoosize
    # slots = spill + width * rowlen; blknnznnz is its 8-bit wrap
    w = ah["tilewidth"].astype(np.int64)
    assert np.all(np.diff(ah["blknnz"])[hyb] == spill[hyb] + w[hyb] * 16)


# --------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(HYB_CASES))
def test_gpu_conversion_with_hyb_is_bit_exact_and_spmv_matches(name, precision):
    from tests.test_gpu_parity import check_matrix
    pi = check_matrix(HYB_CASES[name](), precision, enable_hyb=True)
    assert pi.nchunks > 0


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(chunk_bytes=2560, xstage_bytes=128), dict(chunk_bytes=8192, xstage_bytes=4096)])
def test_gpu_hyb_with_split_rows(cfg):
    from tests.test_gpu_parity import check_matrix
    check_matrix(HYB_CASES["hyb_rich"](), "f64", plan_kwargs=cfg, enable_hyb=True)


@pytest.mark.gpu
@pytest.mark.parametrize("path", G.golden_files(hyb=True), ids=os.path.basename)
def test_gpu_hyb_fixtures(path):
    """Conversion with TILESPMV_ENABLE_HYB against fixtures produced by the reference's own (re-enabled) code."""
    from tests.test_gpu_parity import assert_y_close
    from tilespmv_b200 import api
    d = G.load(path)
    m, n = (int(v) for v in d["in_shape"])
    dm = api.DeviceTileMatrix.from_csr(m, n, d["in_rowptr"], d["in_colidx"], d["in_val"], enable_hyb=True)
    Mg = dm.export()
    G.assert_tile_arrays_equal(Mg.arrays(), G.tile_arrays(d), d["name"] + ":")
    p1, p2, *_ = api.tilespmv_prepare(Mg, m)
    assert np.array_equal(p1, d["ptroffset1"]) and np.array_equal(p2, d["ptroffset2"])
    y = api.Plan(dm).spmv_host(d["x"])
    if d["precision"] == "f64" and np.all(d["in_val"] == np.round(d["in_val"])):
        assert y.tobytes() == d["y"].tobytes()
    else:
        scale = O.Oracle(d["precision"]).csr_abs_spmv(m, d["in_rowptr"], d["in_colidx"], d["in_val"], d["x"])
        assert_y_close(y, d["y"], scale, d["precision"], d["name"])


@pytest.mark.gpu
@needs_refhyb
def test_gpu_upload_of_a_reference_built_tile_matrix_with_hyb_tiles():
    """Tile_matrix structs produced by the reference's own HYB encoder go through upload -> plan -> SpMV."""
    from tilespmv_b200 import api
    m, n, rp, ci, v = HYB_CASES["hyb_ragged"]()
    ref = O.Reference("f64", "refhyb")
    Mr = ref.tile_create(m, n, rp, ci, v)
    assert (ref.arrays(Mr, m)["Format"] == 3).sum() > 0
    M = api.HostTileMatrix(api.F64, m, n)
    C.memmove(C.byref(M.struct), C.byref(Mr), C.sizeof(Mr))
    dm = api.DeviceTileMatrix.upload(M)
    plan = api.Plan(dm)
    x = x_for(n, 1)
    y_ref, *_ = ref.tilespmv_cpu(Mr, m, n, rp, ci, v, x)
    assert plan.spmv_host(x).tobytes() == y_ref.tobytes()


@pytest.mark.gpu
def test_gpu_default_conversion_is_unchanged_by_the_flag_plumbing():
    from tests.test_gpu_parity import check_matrix
    check_matrix(CASES["rmat_12"](), "f64")  # default: no HYB, bit-exact with the default oracle
