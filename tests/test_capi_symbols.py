"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tilespmv.h
declares, and refuses to compute without a GPU (no CPU fallback).  No kernels are launched."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tilespmv_b200 import _capi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "tilespmv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = src[:src.index("#ifdef TILESPMV_REFERENCE_NAMES")]
    names = set(re.findall(r"\b((?:Tile_|call_tilespmv_|tilespmv_)[A-Za-z0-9_]+)\s*\(", src))
    return sorted(n for n in names if not n.startswith("tilespmv_dmat ") and n not in ("tilespmv_dmat", "tilespmv_plan"))


def test_library_exports_every_declared_symbol():
    L = _capi.load()
    declared = _declared_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/tilespmv.h but not exported"
    assert set(declared) == set(_capi.EXPORTS)
    assert b"sm_100a" in L.tilespmv_version()


def test_struct_layout_matches_header():
    # 3 ints + 47 pointer/int fields, natural alignment: the struct the reference's main.cu mallocs
    assert C.sizeof(_capi.TileMatrixF64) == C.sizeof(_capi.TileMatrixF32)
    assert [f[0] for f in _capi.TileMatrixF64._fields_][:7] == ["tilem", "tilen", "tilenum", "tile_ptr",
                                                                 "tile_columnidx", "tile_nnz", "Format"]
    assert len(_capi.TileMatrixF64._fields_) == 50


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    rp = np.array([0, 1, 2], np.int32)
    ci = np.array([0, 1], np.int32)
    v = np.array([1.0, 2.0])
    with pytest.raises(api.TileSpMVError) as e:
        api.DeviceTileMatrix.from_csr(2, 2, rp, ci, v)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(api.TileSpMVError):
        api.Tile_create(2, 2, rp, ci, v)


def test_prepare_is_host_bookkeeping_only():
    """tilespmv_prepare on a Tile_matrix built by the oracle equals the golden ptroffset/schedule."""
    from oracle import oracle_py as O
    from tests import golden_util as G
    for path in G.golden_files("f64"):
        d = G.load(path)
        m, n = (int(v) for v in d["in_shape"])
        ora = O.Oracle("f64")
        Mo = ora.tile_create(m, n, d["in_rowptr"], d["in_colidx"], d["in_val"])
        M = api.HostTileMatrix(api.F64, m, n)
        C.memmove(C.byref(M.struct), C.byref(Mo), C.sizeof(Mo))  # borrow the oracle's arrays
        p1, p2, rbb, a, b, c = api.tilespmv_prepare(M, m)
        assert np.array_equal(p1, d["ptroffset1"]) and np.array_equal(p2, d["ptroffset2"])
        assert rbb == int(d["rowblkblock"][0])
        assert np.array_equal(a, d["blkcoostylerowidx"])
        assert np.array_equal(b, d["blkcoostylerowidx_colstart"])
        assert np.array_equal(c, d["blkcoostylerowidx_colstop"])
        ora.tile_destroy(Mo)


def test_mmio_front_end_matches_oracle(tmp_path):
    from oracle import oracle_py as O
    from tilespmv_b200 import generators as g
    m, n, rp, ci, v = g.banded(300, val_mode=0)
    p1 = str(tmp_path / "a.mtx")
    g.write_mtx(p1, m, n, rp, ci, v)
    p2 = str(tmp_path / "s.mtx")
    with open(p2, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer symmetric\n%c\n6 6 5\n1 1 3\n3 1 -2\n5 2 7\n4 4 1\n6 1 9\n")
    for p in (p1, p2):
        rc, got = api.mmio_allinone(p)
        rco, want = O.Oracle("f64").mtx_read(p)
        assert rc == rco == 0
        assert got[:3] == want[:3]
        for a, b in zip(got[3:], want[3:]):
            assert a.tobytes() == b.tobytes()
    assert api.mmio_allinone(str(tmp_path / "nope.mtx"))[0] == -1
    bad = str(tmp_path / "bad.mtx")
    open(bad, "w").write("hello world\n1 1 1\n")
    assert api.mmio_allinone(bad)[0] == -2


def _same(got, want):
    assert got[:3] == want[:3]
    for a, b in zip(got[3:], want[3:]):
        assert a.tobytes() == b.tobytes()


def test_mmio_parallel_parser_cache_and_fallback(tmp_path, monkeypatch):
    """Files with >= 4096 entries take the parallel parser; anything unusual falls back to the serial loop that
    mimics the reference's fscanf.  Every variant must equal the oracle's reader (itself pinned against the reference's
    mmio_allinone).  TILESPMV_MTX_CACHE switches on the binary CSR cache."""
    from oracle import oracle_py as O
    from tilespmv_b200 import generators as g
    ora = O.Oracle("f64")
    m, n, rp, ci, v = g.rmat(11, val_mode=0)  # ~25 k entries, rows of very different lengths
    big = str(tmp_path / "big.mtx")
    g.write_mtx(big, m, n, rp, ci, v)
    rc, got = api.mmio_allinone(big)
    rco, want = ora.mtx_read(big)
    assert rc == rco == 0
    _same(got, want)
    # symmetric pattern file, exponent / signed values, tabs, blank lines at the end
    rng = np.random.default_rng(3)
    sym = str(tmp_path / "sym.mtx")
    with open(sym, "w") as f:
        pairs = sorted({(int(max(a, b)), int(min(a, b))) for a, b in rng.integers(1, 3000, size=(6000, 2))})
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n% a comment\n%another\n")
        f.write(f"3000 3000 {len(pairs)}\n")
        for a, b in pairs:
            f.write(f"{a}\t{b}\n")
        f.write("\n\n")
    val = str(tmp_path / "val.mtx")
    with open(val, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2000 1000 5000\n")
        for k in range(5000):
            x = [f"{rng.uniform(-1, 1):.17e}", f"+{rng.uniform(0, 9):.3f}", f"{rng.integers(-5, 5)}", "1e-3", "-.5"][k % 5]
            f.write(f"  {rng.integers(1, 2001)} {rng.integers(1, 1001)}   {x}  \n")
    for p in (sym, val):
        rc, got = api.mmio_allinone(p)
        rco, want = ora.mtx_read(p)
        assert rc == rco == 0
        _same(got, want)
    # an out-of-range index in the middle (undefined behaviour upstream: it would write out of bounds): the parallel
    # parser hands over to the serial loop, which stops at the first invalid entry
    bad = str(tmp_path / "bad.mtx")
    lines = open(val).read().split("\n")
    lines[2500] = "99999 1 1.0"
    open(bad, "w").write("\n".join(lines))
    rc, got = api.mmio_allinone(bad)
    assert rc == 0 and len(got[4]) == 2498
    rcv, full = api.mmio_allinone(val)
    order = np.argsort(np.repeat(np.arange(2000), np.diff(full[3])), kind="stable")  # rows are in file order
    assert len(full[4]) == 5000 and order is not None
    # binary cache: second read comes from the cache file and is identical; a changed source invalidates it
    cache = tmp_path / "cache"
    cache.mkdir()
    monkeypatch.setenv("TILESPMV_MTX_CACHE", str(cache))
    rc, first = api.mmio_allinone(big)
    assert rc == 0 and len(list(cache.iterdir())) == 1
    rc, second = api.mmio_allinone(big)
    _same(second, first)
    _same(second, ora.mtx_read(big)[1])
    rc, f32 = api.mmio_allinone(big, api.F32)
    assert rc == 0 and f32[5].dtype == np.float32 and len(list(cache.iterdir())) == 2


def test_mmio_blank_lines_before_the_size_line(tmp_path):
    """A blank / whitespace-only line between the comment block and the size line: the reference (mmio.h:600-612)
    keeps reading lines until one yields three integers; the entries then start right after THAT line."""
    from oracle import oracle_py as O
    ora = O.Oracle("f64")
    small = str(tmp_path / "blank_small.mtx")
    with open(small, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% comment\n\n   \n4 5 3\n1 1 1.5\n4 5 -2\n2 3 7\n")
    big = str(tmp_path / "blank_big.mtx")
    rng = np.random.default_rng(9)
    with open(big, "w") as f:  # >= 4096 entries: the parallel parser
        f.write("%%MatrixMarket matrix coordinate real general\n%c\n\n900 800 5000\n")
        for _ in range(5000):
            f.write(f"{rng.integers(1, 901)} {rng.integers(1, 801)} {rng.uniform(-1, 1):.17e}\n")
    for p, nnz in ((small, 3), (big, 5000)):
        rc, got = api.mmio_allinone(p)
        rco, want = ora.mtx_read(p)
        assert rc == rco == 0 and len(got[4]) == nnz
        _same(got, want)
