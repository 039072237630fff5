"""Pins the C restatement (oracle/tilespmv_oracle.c) against the UNMODIFIED reference CPU path
compiled into oracle/_ref (oracle/ref_shim.c): every Tile_matrix array, ptroffset1/2, the
warp-chunk schedule and y must be byte-identical.  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests.cases import BIG_CASES, CASES, random_case, x_for

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")


def _compare(case, precision):
    m, n, rp, ci, v = case
    ora, ref = O.Oracle(precision), O.Reference(precision)
    v = v.astype(ora.val_dtype)
    Mo = ora.tile_create(m, n, rp, ci, v)
    Mr = ref.tile_create(m, n, rp, ci, v)
    ao, ar = ora.arrays(Mo, m), ref.arrays(Mr, m)
    assert ao.keys() == ar.keys()
    for k in ar:
        assert ao[k].shape == ar[k].shape, k
        assert ao[k].tobytes() == ar[k].tobytes(), f"{k} differs"
    for mode in (1, 0):
        x = x_for(n, mode, ora.val_dtype)
        yo, p1o, p2o = ora.tilespmv_cpu(Mo, m, n, x)
        yr, p1r, p2r, sched_r = ref.tilespmv_cpu(Mr, m, n, rp, ci, v, x)
        assert yo.tobytes() == yr.tobytes()
        assert np.array_equal(p1o, p1r) and np.array_equal(p2o, p2r)
    sched_o = ora.schedule(Mo)
    assert sched_o[0] == sched_r[0]
    for a, b in zip(sched_o[1:], sched_r[1:]):
        assert np.array_equal(a, b)
    # ptroffset1 equals the tile's own format prefix
    fmt = ao["Format"]
    names = {0: "csr_offset", 1: "coo_offset", 2: "ell_offset", 4: "dns_offset", 5: "dnsrow_offset",
             6: "dnscol_offset"}
    for f, name in names.items():
        sel = fmt == f
        assert np.array_equal(p1o[sel], ao[name][:-1][sel])
    assert np.array_equal(p2o[fmt == 0], ao["csrptr_offset"][:-1][fmt == 0])
    ora.tile_destroy(Mo)
    return ao


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_f64(name):
    _compare(CASES[name](), "f64")


@pytest.mark.parametrize("name", ["seven_formats", "lap2d_64", "banded_8k_real", "rmat_12_real", "ragged_band",
                                  "band_unsorted"])
def test_oracle_matches_reference_f32(name):
    _compare(CASES[name](), "f32")


@pytest.mark.parametrize("name", sorted(BIG_CASES))
def test_oracle_matches_reference_big(name):
    _compare(BIG_CASES[name](), "f64")


@pytest.mark.parametrize("seed", range(24))
def test_oracle_matches_reference_on_random_structure(seed):
    """The inputs of the GPU differential test (tests/test_gpu_parity.py): all seven formats occur."""
    _compare(random_case(np.random.default_rng(1000 + seed)), "f64" if seed % 3 else "f32")


def test_seven_format_fixture_values():
    """The expected dump of SURVEY.md Appendix C.3."""
    a = _compare(CASES["seven_formats"](), "f64")
    assert a["Format"].tolist() == [5, 6, 1, 4, 0, 2]
    assert a["tile_ptr"].tolist() == [0, 3, 6]
    assert a["tile_nnz"].tolist() == [0, 32, 64, 67, 323, 353, 385]
    assert a["blknnznnz"].tolist() == [32, 32, 3, 0, 30, 32, 0]
    assert a["tilewidth"].tolist() == [0, 0, 0, 0, 0, 2]
    assert a["denserowid"].tolist() == [3, 7] and a["densecolid"].tolist() == [1, 5]
    assert a["Blockcsr_Ptr"].tolist() == [0, 10, 11, 19, 19, 22, 22, 22, 27, 27, 27, 29, 29, 29, 29, 29]
    assert a["coo_compressed_Idx"].tolist() == [1, 87, 240]
    assert a["deferredcoo_colidx"].tolist() == [33, 39, 32]


def test_mtx_reader_matches_reference(tmp_path):
    from tilespmv_b200 import generators as g
    m, n, rp, ci, v = g.banded(512, val_mode=0)
    path = str(tmp_path / "band.mtx")
    g.write_mtx(path, m, n, rp, ci, v)
    sym = str(tmp_path / "sym.mtx")
    with open(sym, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n% comment\n5 5 4\n1 1\n3 1\n5 2\n4 4\n")
    for p in (path, sym):
        rc_o, o = O.Oracle("f64").mtx_read(p)
        rc_r, r = O.Reference("f64").mtx_read(p)
        assert rc_o == rc_r == 0
        assert o[:3] == r[:3]
        for a, b in zip(o[3:], r[3:]):
            assert a.tobytes() == b.tobytes()
    assert O.Oracle("f64").mtx_read(str(tmp_path / "missing.mtx"))[0] == -1
