"""GPU parity tests (run with -m gpu on a B200): every call goes through the C-ABI of
libtilespmv_b200.so; the oracle (oracle/) is only the checker.

  * Tile_create on the GPU is BIT-EXACT with the reference conversion (all Tile_matrix arrays)
  * y = A*x matches tilespmv_cpu: bit-exact on the reference driver's integer data
    (val = j%10, x = i%10, main.cu:68-69,93-97), and within 1e-12 (fp64) / 1e-5 (fp32) of
    sum_j |a_ij||x_j| on seeded uniform(-1,1) data -- the tolerance BASELINE.json states
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from tests import golden_util as G
from tests.cases import BIG_CASES, CASES, random_case, x_for
from tilespmv_b200 import api

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}


def _prec(precision):
    return api.F64 if precision == "f64" else api.F32


def assert_y_close(y, y_ref, scale, precision, what=""):
    tol = TOL[precision]
    err = np.abs(y.astype(np.float64) - y_ref.astype(np.float64))
    bound = tol * np.maximum(scale.astype(np.float64), np.finfo(np.float64).tiny)
    bad = np.flatnonzero(err > bound)
    assert len(bad) == 0, (f"{what}: {len(bad)} rows out of tolerance, first {bad[:8]}, got {y[bad[:8]]}, "
                           f"want {y_ref[bad[:8]]}")


def check_matrix(case, precision, plan_kwargs=None, exact_modes=(1,), real_modes=(0,), enable_hyb=False):
    m, n, rp, ci, v = case
    ora = O.Oracle(precision, enable_hyb=enable_hyb)
    v = v.astype(ora.val_dtype)
    Mo = ora.tile_create(m, n, rp, ci, v)
    want = ora.arrays(Mo, m)
    # --- conversion on the GPU, bit-exact ---
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v, enable_hyb=enable_hyb)
    Mg = dm.export()
    G.assert_tile_arrays_equal(Mg.arrays(), want, "Tile_matrix.")
    info = dm.info()
    assert info.tilenum == Mo.tilenum and info.nnz == int(rp[m]) and info.nnz_side == Mo.coototal
    assert list(info.tiles_by_format) == [int((want["Format"] == f).sum()) for f in range(7)]
    sc = want["scalars"]  # tilem, tilen, tilenum, csrsize, csrptrlen, coosize, ellsize, hybsize, hybellsize, hybcoosize, dns...
    assert list(info.slots_by_format) == [int(sc[3]), int(sc[5]), int(sc[6]), int(sc[7]), int(sc[10]), int(sc[11]), int(sc[12])]
    # --- SpMV ---
    plan = api.Plan(dm, **(plan_kwargs or {}))
    integer_vals = bool(np.all(v == np.round(v)))
    for mode in exact_modes + real_modes:
        x = x_for(n, mode, ora.val_dtype)
        y_ref, _, _ = ora.tilespmv_cpu(Mo, m, n, x)
        y = plan.spmv_host(x)
        if mode == 1 and integer_vals and (precision == "f64" or np.abs(y_ref).max(initial=0) < 2 ** 24):
            assert y.tobytes() == y_ref.tobytes(), f"integer data must be bit-exact (mode {mode})"
        scale = ora.csr_abs_spmv(m, rp, ci, v, x)
        assert_y_close(y, y_ref, scale, precision, f"mode {mode}")
    pi = plan.info()
    plan.destroy()
    dm.destroy()
    Mg.destroy()
    ora.tile_destroy(Mo)
    return pi


@pytest.mark.parametrize("name", sorted(CASES))
def test_convert_and_spmv_f64(name):
    check_matrix(CASES[name](), "f64")


@pytest.mark.parametrize("name", ["seven_formats", "lap2d_64", "banded_8k_real", "rmat_12_real", "ragged_band",
                                  "band_unsorted", "ragged_seven", "uniform_8k", "dense_48", "hub_rows"])
def test_convert_and_spmv_f32(name):
    check_matrix(CASES[name](), "f32")


@pytest.mark.parametrize("name", ["seven_formats", "banded_8k", "banded_8k_real", "band_contig_8k", "rmat_12_real",
                                  "ragged_band", "ragged_seven", "band_unsorted", "dense_48"])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_csr_groups_and_individual_csr_tiles_agree(name, precision):
    """The planner merges the CSR tiles of a block row into one CSR group (default); with
    TILESPMV_PLAN_NO_CSR_GROUPS every CSR tile stays an individual tile.  Both must match the oracle."""
    pg = check_matrix(CASES[name](), precision)
    pi = check_matrix(CASES[name](), precision, plan_kwargs=dict(csr_groups=False))
    assert pi.csr_groups == 0
    if name.startswith("band"):
        assert pg.csr_groups > 0


@pytest.mark.parametrize("name,precision,panel", [("uniform_8k", "f64", 8192), ("uniform_8k", "f32", 4096), ("rmat_12", "f64", 4096),
                                                  ("rmat_12_real", "f32", 1024), ("ragged_rmat", "f64", 2048),
                                                  ("seven_formats", "f64", 128), ("lap2d_64", "f64", 4096),
                                                  ("empty_rows", "f64", 128), ("band_contig_8k", "f64", 8192), ("hub_rows", "f64", 8192)])
def test_x_panels_accumulating_sub_plans(name, precision, panel):
    """xpanel_bytes forces the side matrix into column panels (one accumulating launch per panel, the layout
    meant for x far larger than L2): same results, also with rows cut into pieces inside a panel."""
    pi = check_matrix(CASES[name](), precision, plan_kwargs=dict(xpanel_bytes=panel))
    if name not in ("band_contig_8k",):
        assert pi.xpanels > 1 and pi.launches_per_spmv >= pi.xpanels - 1
    pi = check_matrix(CASES[name](), precision, plan_kwargs=dict(xpanel_bytes=panel, chunk_bytes=2560, xstage_bytes=128))
    pi = check_matrix(CASES[name](), precision, plan_kwargs=dict(xpanel_bytes=-1))
    assert pi.xpanels == 1


@pytest.mark.parametrize("name", sorted(BIG_CASES))
def test_convert_and_spmv_big(name):
    pi = check_matrix(BIG_CASES[name](), "f64")
    assert pi.nchunks > 0 and pi.stream_bytes > 0


@pytest.mark.parametrize("name", ["seven_formats", "rmat_12", "band_contig_8k", "uniform_8k", "lap3d27_24",
                                  "ragged_rmat"])
@pytest.mark.parametrize("cfg", [dict(chunk_bytes=2560, xstage_bytes=128), dict(chunk_bytes=2560, xstage_bytes=256),
                                 dict(chunk_bytes=8192, xstage_bytes=4096), dict(chunk_bytes=4096, xstage_bytes=3072, ctas_per_sm=1)])
def test_chunk_configurations_and_split_rows(name, cfg):
    """Tiny x-staging budgets force block rows to be cut into pieces (scratch + fix-up kernel)."""
    pi = check_matrix(CASES[name](), "f64", plan_kwargs=cfg)
    if cfg["xstage_bytes"] == 128 and name in ("rmat_12", "band_contig_8k", "uniform_8k", "lap3d27_24"):
        assert pi.split_rows > 0 and pi.launches_per_spmv >= 2


@pytest.mark.parametrize("path", G.golden_files(), ids=os.path.basename)
def test_golden_fixtures_through_the_drop_in_entry_points(path):
    """Tile_create -> tilespmv_prepare -> call_tilespmv_cuda exactly as main.cu:87-180 calls them,
    checked against fixtures produced by the unmodified reference CPU path."""
    d = G.load(path)
    m, n = (int(v) for v in d["in_shape"])
    os.environ["TILESPMV_BENCH_REPEAT"] = "3"
    os.environ["TILESPMV_WARMUP_NUM"] = "1"
    cwd = os.getcwd()
    M = api.Tile_create(m, n, d["in_rowptr"], d["in_colidx"], d["in_val"])
    G.assert_tile_arrays_equal(M.arrays(), G.tile_arrays(d), d["name"] + ":")
    p1, p2, rbb, a, b, c = api.tilespmv_prepare(M, m)
    assert np.array_equal(p1, d["ptroffset1"]) and np.array_equal(p2, d["ptroffset2"])
    assert rbb == int(d["rowblkblock"][0]) and np.array_equal(a, d["blkcoostylerowidx"])
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            y = api.call_tilespmv_cuda("golden.mtx", M, m, n, int(d["in_rowptr"][m]), d["x"])
            rows = open("results.csv").read().strip().split("\n")
        finally:
            os.chdir(cwd)
    assert rows[-1].startswith(f"golden.mtx,{m},{n},{int(d['in_rowptr'][m])},")
    if d["precision"] == "f64" and np.all(d["in_val"] == np.round(d["in_val"])):
        assert y.tobytes() == d["y"].tobytes()
    else:
        ora = O.Oracle(d["precision"])
        scale = ora.csr_abs_spmv(m, d["in_rowptr"], d["in_colidx"], d["in_val"], d["x"])
        assert_y_close(y, d["y"], scale, d["precision"], d["name"])
    M.destroy()


def test_upload_of_a_reference_built_tile_matrix():
    """A Tile_matrix produced on the CPU (here: by the oracle) can be uploaded and multiplied."""
    m, n, rp, ci, v = CASES["seven_formats"]()
    ora = O.Oracle("f64")
    Mo = ora.tile_create(m, n, rp, ci, v)
    M = api.HostTileMatrix(api.F64, m, n)
    C.memmove(C.byref(M.struct), C.byref(Mo), C.sizeof(Mo))
    dm = api.DeviceTileMatrix.upload(M)
    plan = api.Plan(dm)
    x = x_for(n, 1)
    y_ref, _, _ = ora.tilespmv_cpu(Mo, m, n, x)
    assert plan.spmv_host(x).tobytes() == y_ref.tobytes()
    ora.tile_destroy(Mo)


def test_host_batch_pipeline_equals_single_calls():
    """tilespmv_plan_spmv_host_batch (3-stream pipeline over a ring of device buffers) gives bit-identical
    results to one tilespmv_plan_spmv_host call per vector, for more vectors than ring slots, pinned or not."""
    import torch
    m, n, rp, ci, v = CASES["rmat_12"]()  # has rows cut across chunks: the shared scratch must not race
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm, chunk_bytes=2560, xstage_bytes=256)
    assert plan.info().split_rows > 0
    rng = np.random.default_rng(5)
    nvec = 8
    xs = [torch.from_numpy(rng.uniform(-1, 1, n)) for _ in range(nvec)]
    want = [plan.spmv_host(x.numpy()) for x in xs]
    for pinned in (True, False):
        xin = [x.pin_memory() if pinned else x.clone() for x in xs]
        ys = [torch.full((m,), float("nan"), dtype=torch.float64) for _ in range(nvec)]
        if pinned:
            ys = [y.pin_memory() for y in ys]
        plan.spmv_host_batch([x.data_ptr() for x in xin], [y.data_ptr() for y in ys])
        for y, w in zip(ys, want):
            assert y.numpy().tobytes() == w.tobytes()
    plan.spmv_host_batch([], [])


@pytest.mark.parametrize("name", ["lap2d_64", "rmat_12_real", "band_contig_8k"])
def test_iterate_graph_equals_a_loop_of_spmv_calls(name):
    """tilespmv_plan_iterate (niters ping-pong launches replayed as one CUDA graph) is bit-identical to calling
    tilespmv_plan_spmv in a loop; the graph is re-used across calls and rebuilt when niters changes."""
    import torch
    m, n, rp, ci, v = CASES[name]()
    assert m == n
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v * 0.01)
    plan = api.Plan(dm, chunk_bytes=2560, xstage_bytes=256) if name == "rmat_12_real" else api.Plan(dm)
    x0 = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, n)).cuda()
    launches0 = api._capi.load().tilespmv_kernel_launch_count()
    for niters in (5, 5, 4, 1):
        a, b = x0.clone(), torch.empty_like(x0)
        for i in range(niters):
            src, dst = (a, b) if i % 2 == 0 else (b, a)
            plan.spmv(src.data_ptr(), dst.data_ptr())
        want = (a if niters % 2 == 0 else b).clone()
        xa, xb = x0.clone(), torch.full_like(x0, float("nan"))
        plan.iterate(xa.data_ptr(), xb.data_ptr(), niters)
        torch.cuda.synchronize()
        got = xa if niters % 2 == 0 else xb
        assert torch.equal(got, want), niters
    per = plan.info().launches_per_spmv
    assert api._capi.load().tilespmv_kernel_launch_count() - launches0 == 2 * (5 + 5 + 4 + 1) * per
    with pytest.raises(api.TileSpMVError):
        m2, n2, rp2, ci2, v2 = CASES["seven_formats"]()  # 32 x 40: not square
        api.Plan(api.DeviceTileMatrix.from_csr(m2, n2, rp2, ci2, v2)).iterate(xa.data_ptr(), xb.data_ptr(), 2)


@pytest.mark.parametrize("name", ["lap2d_256", "lap3d27_48", "uniform_64k", "rmat_15"])
def test_back_to_back_launches_see_the_previous_result(name):
    """Consecutive launches of one stream overlap their set-up with the previous kernel's tail (programmatic dependent
    launch: the kernel fetches its first chunks of the immutable stream, then waits for the grid dependency before it
    touches x or y).  A chain x <- A*x issued back to back must equal the same chain with a device synchronisation after
    every launch, bit for bit -- on the default stream and on a side stream, with sub-plans (x panels) in between."""
    import torch
    m, n, rp, ci, v = BIG_CASES[name]()
    assert m == n
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v * (1.0 / 64))
    # default plan, x panels (accumulating sub-plans), tiny staging budgets (rows cut into pieces: the fix-up kernels are
    # links of the chain too)
    for kw in ({}, {"xpanel_bytes": 64 * 1024}, {"chunk_bytes": 2560, "xstage_bytes": 128}):
        plan = api.Plan(dm, **kw)
        x0 = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, n)).cuda()
        niters = 12

        def chain(sync_every, stream):
            a, b = x0.clone(), torch.full_like(x0, float("nan"))
            torch.cuda.synchronize()
            for i in range(niters):
                src, dst = (a, b) if i % 2 == 0 else (b, a)
                plan.spmv(src.data_ptr(), dst.data_ptr(), stream)
                if sync_every:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            return (a if niters % 2 == 0 else b).clone()

        want = chain(True, 0)
        assert torch.isfinite(want).all()
        st = torch.cuda.Stream()
        for _ in range(3):
            assert torch.equal(chain(False, 0), want)
            assert torch.equal(chain(False, st.cuda_stream), want)
        plan.destroy()


def test_device_pointer_path_streams_and_linearity():
    """tilespmv_plan_spmv on torch device buffers and a non-default stream; A(ax+by) = aAx + bAy."""
    import torch
    m, n, rp, ci, v = BIG_CASES["lap3d27_48"]()
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm)
    g = torch.Generator(device="cpu").manual_seed(0)
    x1 = torch.randint(-8, 9, (n,), generator=g).double().cuda()
    x2 = torch.randint(-8, 9, (n,), generator=g).double().cuda()
    y1, y2, y3 = (torch.full((m,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan.spmv(x1.data_ptr(), y1.data_ptr(), st.cuda_stream)
        plan.spmv(x2.data_ptr(), y2.data_ptr(), st.cuda_stream)
        x3 = 3 * x1 - 2 * x2
        plan.spmv(x3.data_ptr(), y3.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert torch.equal(y3, 3 * y1 - 2 * y2)  # small integers: exact
    # against torch's own CSR SpMV as an independent check
    A = torch.sparse_csr_tensor(torch.from_numpy(rp).long(), torch.from_numpy(ci).long(),
                                torch.from_numpy(v), size=(m, n)).cuda()
    assert torch.equal(A @ x1, y1)
    ms = plan.time(x1.data_ptr(), y1.data_ptr(), warmup=2, iters=5)
    assert ms > 0
    with pytest.raises(api.TileSpMVError):
        plan.spmv(x1.data_ptr() + 8, y1.data_ptr())  # misaligned x


def test_config1_lap2d_1024_against_oracle():
    """BASELINE config 1: 2-D 5-point Laplacian 1024^2 through conversion + SpMV, full size."""
    from tilespmv_b200 import generators as g
    pi = check_matrix(g.lap2d(1024, val_mode=1), "f64")
    assert pi.algorithmic_bytes == 67197320  # B_alg(C1) of SURVEY.md Appendix D


def test_config2_lap3d27_160_against_oracle_full_size():
    """BASELINE config 2 (the bench workload): 3-D 27-point Laplacian 160^3, conversion bit-exact against the
    oracle, y bit-exact on integer data and within 1e-12 on real data, at full size."""
    from tilespmv_b200 import generators as g
    pi = check_matrix(g.lap3d27(160, val_mode=1), "f64")
    assert pi.algorithmic_bytes == 1037093192  # B_alg(C2) of SURVEY.md Appendix D
    assert pi.split_rows == 0 and pi.launches_per_spmv == 1


@pytest.mark.parametrize("name,precision", [("banded_1m", "f64"), ("rmat_18", "f32"), ("uniform_2m", "f64"),
                                            ("band_contig_1m", "f32")])
def test_reduced_configs_3_4_5_and_size_independent_properties(name, precision):
    """Configs 3/4/5 at the largest size the oracle finishes in seconds, plus properties that need no oracle:
    linearity A(a*x + b*z) = a*A*x + b*A*z and y(e_j) = column j."""
    from tilespmv_b200 import generators as g
    case = {"banded_1m": lambda: g.banded(1 << 20, val_mode=0), "rmat_18": lambda: g.rmat(18, val_mode=0),
            "uniform_2m": lambda: g.uniform(1 << 21, val_mode=1), "band_contig_1m": lambda: g.band_contig(1 << 20, val_mode=0)}[name]()
    check_matrix(case, precision)
    m, n, rp, ci, v = case
    dt = np.float64 if precision == "f64" else np.float32
    v = v.astype(dt)
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm)
    rng = np.random.default_rng(3)
    x, z = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, n).astype(dt)
    a, b = dt(0.75), dt(-1.5)
    lhs = plan.spmv_host((a * x + b * z).astype(dt)).astype(np.float64)
    rhs = a * plan.spmv_host(x).astype(np.float64) + b * plan.spmv_host(z).astype(np.float64)
    ora = O.Oracle(precision)
    scale = ora.csr_abs_spmv(m, rp, ci, v, (np.abs(x) + np.abs(z)).astype(dt)).astype(np.float64) * 2.25
    assert np.all(np.abs(lhs - rhs) <= 8 * TOL[precision] * np.maximum(scale, 1e-300))
    j = int(ci[len(ci) // 2])  # a column that certainly has an entry
    e = np.zeros(n, dt)
    e[j] = 1
    col = plan.spmv_host(e)
    want = np.zeros(m, np.float64)
    rows = np.repeat(np.arange(m), np.diff(rp))
    sel = ci == j
    np.add.at(want, rows[sel], v[sel].astype(np.float64))
    assert np.array_equal(col.astype(np.float64), want.astype(dt).astype(np.float64))


def _y_against_csr(case, precision, plan_kwargs=None):
    """y through the library vs the plain CSR loop (main.cu:101-110) for inputs whose conversion the oracle would
    take too long to restate; tolerance relative to sum |a||x| as everywhere."""
    m, n, rp, ci, v = case
    dt = np.float64 if precision == "f64" else np.float32
    v = v.astype(dt)
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm, **(plan_kwargs or {}))
    x = np.random.default_rng(11).uniform(-1, 1, n).astype(dt)
    y = plan.spmv_host(x)
    ora = O.Oracle(precision)
    y_ref = ora.csr_spmv(m, rp, ci, v, x, parallel=True)
    scale = ora.csr_abs_spmv(m, rp, ci, v, x)
    assert_y_close(y, y_ref, scale, precision)
    return dm.info(), plan.info()


def test_config5_shard_at_full_width_uses_x_panels():
    """One row block of BASELINE config 5 at its true width: 1 M rows x 50 M columns, 20 per row.  x is 400 MB, far
    larger than L2, so the planner cuts the side matrix into column panels on its own."""
    from tilespmv_b200 import generators as g
    di, pi = _y_against_csr(g.uniform_rows(50_000_000, 0, 1 << 20, val_mode=0), "f64")
    assert di.nnz_side == di.nnz == 20 << 20
    assert pi.xpanels > 1 and pi.launches_per_spmv >= pi.xpanels


def test_config4_rmat_scale_22_fp32_hub_rows():
    """BASELINE config 4 at scale 22 (65 M nnz, fp32): almost everything is extracted side entries, hub rows are cut
    into hundreds of pieces whose partial sums the fix-up kernels add."""
    from tilespmv_b200 import generators as g
    di, pi = _y_against_csr(g.rmat(22, val_mode=0), "f32")
    assert di.nnz_side > 0.9 * di.nnz and pi.split_rows > 0 and pi.launches_per_spmv >= 2


@pytest.mark.parametrize("seed", range(24))
def test_random_matrices_all_plan_variants(seed):
    """Differential test: conversion bit-exact against the oracle, y against tilespmv_cpu, for every plan variant
    (CSR groups on / off, x panels, tiny chunks that cut rows, flat side chunks on / off, HYB rule) on random structure."""
    rng = np.random.default_rng(1000 + seed)
    case = random_case(rng)
    precision = "f64" if seed % 3 else "f32"
    variants = [dict(), dict(csr_groups=False), dict(xpanel_bytes=256), dict(chunk_bytes=2560, xstage_bytes=128),
                dict(chunk_bytes=2560, xstage_bytes=256, xpanel_bytes=512, csr_groups=False), dict(flat_side=False),
                dict(flat_side=False, xpanel_bytes=384, chunk_bytes=2560, xstage_bytes=256)]
    check_matrix(case, precision, plan_kwargs=variants[seed % len(variants)], enable_hyb=seed % 4 == 0)
    check_matrix(case, precision, plan_kwargs=variants[(seed + 1) % len(variants)])


# ---- per-format cost profiling (cf. DEBUG_FORMATCOST / formatprofile, tilespmv_cuda.h:102-111, main.cu:12) ----
@pytest.mark.parametrize("name", ["seven_formats", "band_contig_8k", "rmat_12", "lap3d27_24", "ragged_seven", "hyb_rich"])
def test_single_format_plans_partition_the_matrix(name):
    """A plan restricted to one tile format multiplies exactly the nonzeros of that format (bit 1 = the extracted side
    entries): on the reference driver's integer data the seven partial y's add up to y bit for bit, the empty mask gives
    zeros, and tilespmv_format_profile reports a time for every format that is present."""
    import torch
    from tests.cases import HYB_CASES
    hyb = name == "hyb_rich"
    m, n, rp, ci, v = (HYB_CASES if hyb else CASES)[name]()
    v = (np.arange(len(ci)) % 10).astype(np.float64)
    x = x_for(n, 1)
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v, enable_hyb=hyb)
    info = dm.info()
    y_all = api.Plan(dm).spmv_host(x)
    total = np.zeros(m)
    for f in range(7):
        y_f = api.Plan(dm, format_mask=1 << f).spmv_host(x)
        present = info.tiles_by_format[f] > 0 or (f == 1 and info.nnz_side > 0)
        assert present or not y_f.any(), f"format {f} is absent but its plan produced values"
        total += y_f
    assert total.tobytes() == y_all.tobytes()
    assert not api.Plan(dm, format_mask=0x80).spmv_host(x).any()
    d_x, d_y = torch.from_numpy(x).cuda(), torch.empty(max(m, 1), dtype=torch.float64, device="cuda")
    ms, nnz = api.format_profile(dm, d_x.data_ptr(), d_y.data_ptr(), warmup=1, iters=3)
    assert nnz[8] == int(rp[m]) and sum(nnz[:7]) == nnz[8] and nnz[1] == info.nnz_side
    for f in range(7):
        present = info.tiles_by_format[f] > 0 or (f == 1 and info.nnz_side > 0)
        assert (ms[f] > 0) == present and (nnz[f] > 0) == present
    assert ms[7] > 0 and ms[8] > 0


# ---- binary cache of the packed plan ----
@pytest.mark.parametrize("name,kw", [("seven_formats", {}), ("banded_8k_real", {}), ("hub_rows", {}), ("rmat_12_real", dict(xpanel_bytes=8192))])
def test_plan_cache_round_trip(tmp_path, name, kw):
    """tilespmv_plan_save / tilespmv_plan_load: the loaded plan (with its x-panel sub-plans and split rows) gives the
    same y bit for bit without the Tile_matrix; corrupt or foreign files are refused."""
    m, n, rp, ci, v = CASES[name]()
    x = x_for(n, 0)
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
    plan = api.Plan(dm, **kw)
    y = plan.spmv_host(x)
    path = str(tmp_path / "plan.bin")
    plan.save(path)
    i0 = plan.info()
    plan.destroy()
    dm.destroy()
    loaded = api.Plan.load(path, api.F64, m, n)
    i1 = loaded.info()
    assert (i1.nchunks, i1.stream_bytes, i1.split_rows, i1.xpanels, i1.launches_per_spmv, i1.algorithmic_bytes) == \
           (i0.nchunks, i0.stream_bytes, i0.split_rows, i0.xpanels, i0.launches_per_spmv, i0.algorithmic_bytes)
    assert loaded.spmv_host(x).tobytes() == y.tobytes()
    assert loaded.spmv_host(2 * x).tobytes() == (2 * y).tobytes()
    blob = bytearray(open(path, "rb").read())
    blob[len(blob) // 2] ^= 0x5a
    bad = str(tmp_path / "bad.bin")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(api.TileSpMVError):
        api.Plan.load(bad, api.F64, m, n)
    open(bad, "wb").write(bytes(blob[: len(blob) // 3]))
    with pytest.raises(api.TileSpMVError):
        api.Plan.load(bad, api.F64, m, n)
    with pytest.raises(api.TileSpMVError):
        api.Plan.load(str(tmp_path / "missing.bin"), api.F64, m, n)
    # a damaged header that promises 2^60 bytes of payload is an I/O error, not a 2^60-byte allocation
    blob = bytearray(open(path, "rb").read())
    blob[24:32] = (1 << 60).to_bytes(8, "little")  # PlanFileHeader.payload_bytes
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(api.TileSpMVError):
        api.Plan.load(bad, api.F64, m, n)
