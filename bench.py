#!/usr/bin/env python
"""bench.py -- fp64 SpMV GFLOP/s (2*nnz/t) + achieved HBM GB/s for the TileSpMV hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|c3|c5]

N = 1 (workload c2): BASELINE config 2 (3-D 27-point Laplacian 160^3, fp64, ~109 M nnz); a step is one y = A*x through
tilespmv_plan_spmv on device-resident data.

N > 1: the north star's multi-GPU loop -- the matrix cut into N row blocks (one rank per GPU), x replicated, a step is
one iteration of x <- A*x INCLUDING the per-iteration all-gather of the y slices (strong scaling: the global matrix is
fixed).  workload auto = BASELINE config 3 (banded 8 M rows, ~296 M nnz) at N = 2 / 4 and config 5 (uniform random
50 M x 50 M, 10^9 nnz) at N = 8.  `value` is the loop with the library's pipelined exchange (tilespmv_dist_iterate);
the NCCL and fused-epilogue exchanges, the SpMV without exchange, the same matrix on ONE GPU (rank 0 runs all N
row-block plans back to back) and the weak-scaling stencil of round 1 are reported next to it.  All results are
verified inside the run (exchanges bitwise equal, K iterations against torch's CSR SpMV + torch all-gather).

Prints ONE JSON line (rank 0).  `value` is device-timed with CUDA events on the launching stream (max over ranks);
`e2e` is the same metric through HOST buffers (pinned H2D of x + D2H of y inside the timed region); `roofline` relates
the SpMV kernel to the measured HBM copy bandwidth (MEASURED_PEAKS.json); `cpu_baseline` times the reference's own CPU
path (oracle/_ref, else the oracle port) on a bounded sample on this box's host cores.

--impl reference times only that CPU path (rank 0) on the same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fp64 SpMV GFLOP/s (2*nnz/t)"
UNIT = "GFLOP/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
NVLINK_PEER_GBS = 770.0    # measured peer copy per direction (B200_PROFILING.md)
MODES = ("nccl", "fused", "pipelined", "halo")
DEFAULT_MODES = ("nccl", "fused", "pipelined")  # halo: --with-halo (no gain over fused on this pool)
GATHER_PROBE_GPS = 263.0   # scattered 8-byte loads from an L2-resident window, tools/gather_probe.cu (profiles/r01_gather_probe.log)
LAUNCH_NOTE = ("steps are launched back to back on one stream; every launch carries the programmatic-stream-serialization attribute: "
               "its set-up and first fetches of the (immutable) packed matrix stream overlap the previous launch's tail, and the "
               "kernel waits for the grid dependency before it reads x or writes y (TILESPMV_NO_PDL=1 switches it off)")
SEGMENT = 50               # x <- A*x restarts from x0 every SEGMENT iterations (keeps the iterates finite for any K)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """dram__bytes_read+write per launch of tile_spmv_kernel from the committed ncu capture (a RECORDED figure of an
    earlier run of the same kernel on the same workload, not measured by this run)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def emit(obj):
    """One JSON line in ONE write (NCCL_DEBUG banners of other ranks must not land inside it)."""
    sys.stdout.flush()
    os.write(1, (json.dumps(obj) + "\n").encode())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread, self.idx = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gpu_index(local_rank):
    if "CUDA_VISIBLE_DEVICES" in os.environ:
        return os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]
    return local_rank


def nvlink_kib(idx):
    """Sum of the NVLink data counters of one GPU (nvidia-smi nvlink -gt d): (tx KiB, rx KiB) or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(idx)], capture_output=True, text=True, timeout=20).stdout
        tx = rx = 0
        seen = False
        for line in out.splitlines():
            t = line.replace(":", " ").split()
            if "Tx" in t and "KiB" in t:
                tx += int(t[t.index("KiB") - 1])
                seen = True
            if "Rx" in t and "KiB" in t:
                rx += int(t[t.index("KiB") - 1])
                seen = True
        return (tx, rx) if seen else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def pick_workload(args, world):
    if args.workload != "auto":
        return args.workload
    if world == 1:
        return "c2"
    return "c5" if world >= 8 else "c3"


def workload_rows(wl, args, world):
    """Global size and the row cuts (equal row blocks, multiples of 16: the row weights of these matrices are uniform,
    so the byte-balanced cuts of tilespmv_partition_rows coincide with them up to one block row)."""
    n = {"c3": args.c3_rows, "c5": args.c5_rows}[wl]
    per = (n // 16 // world) * 16
    return n, [(r * per, (r + 1) * per if r < world - 1 else n) for r in range(world)]


def workload_gen(wl, n, r0, r1):
    from tilespmv_b200 import generators as g
    if wl == "c3":
        return g.banded_rows(n, r0, r1 - r0)
    return g.uniform_rows(n, r0, r1 - r0)


def workload_config(wl, args, world):
    if wl == "c2":
        G = args.grid
        return {"workload": f"BASELINE config 2: 3-D 27-point Laplacian {G}^3 fp64 ({G ** 3} rows, ~{(3 * G - 2) ** 3 / 1e6:.0f} M nnz) on 1 GPU, "
                            "one y = A*x per step",
                "cache": "inputs larger than L2 (packed stream ~1.07 GB vs 126 MB L2); no flush",
                "launch": LAUNCH_NOTE}
    n = {"c3": args.c3_rows, "c5": args.c5_rows}[wl]
    what = {"c3": f"BASELINE config 3: banded FEM-like {n} x {n}, half-bandwidth 64, 37 nnz/row (~{n * 37 / 1e6:.0f} M nnz) fp64",
            "c5": f"BASELINE config 5: uniform random {n} x {n}, 20 nnz/row ({n * 20 / 1e9:.2f} G nnz) fp64"}[wl]
    return {"workload": what + f", row-block sharded over {world} GPUs; one step = one iteration of x <- A*x incl. the all-gather of x",
            "partition": f"{world} contiguous row blocks of tiles (cuts at multiples of 16 rows), x replicated",
            "exchange": "value: the fastest verified exchange of tilespmv_dist_iterate (named in headline_exchange); `iterate` lists all three",
            "cache": "inputs larger than L2 (packed stream per GPU > 126 MB L2); no flush", "launch": LAUNCH_NOTE}


# ------------------------------------------------------------------------------------------------
# CPU baseline (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _quiet_stdout():
    class Q:
        def __enter__(self):
            sys.stdout.flush()
            self.devnull = os.open(os.devnull, os.O_WRONLY)
            self.saved = os.dup(1)
            os.dup2(self.devnull, 1)

        def __exit__(self, *a):
            os.dup2(self.saved, 1)
            os.close(self.devnull)
            os.close(self.saved)
    return Q()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_sample(wl, args):
    """A bounded sample of the workload the CPU path finishes in a fraction of a second per call."""
    from tilespmv_b200 import generators as g
    if wl == "c2":
        G = args.grid
        m, n, rp, ci, v = g.lap3d27(G, val_mode=0)
        return (m, n, rp, ci, v), f"3-D 27-pt Laplacian {G}^3 fp64 ({int(rp[m])} nnz: the whole config-2 matrix)"
    if wl == "c3":
        rows = min(args.c3_rows, args.cpu_sample_rows or 1_000_000)
        m, n, rp, ci, v = g.banded_rows(args.c3_rows, 0, rows)
        return (m, n, rp, ci, v), f"first {rows} rows of the config-3 banded matrix ({int(rp[m])} nnz, x has {n} entries)"
    rows = min(args.c5_rows, args.cpu_sample_rows or 500_000)
    m, n, rp, ci, v = g.uniform_rows(args.c5_rows, 0, rows)
    return (m, n, rp, ci, v), f"first {rows} rows of the config-5 uniform matrix ({int(rp[m])} nnz, x has {n} entries)"


def cpu_baseline(wl, args, steps=3, warm=0, budget_s=150.0):
    """Reference CPU path (tilespmv_cpu, tilespmv_cpu.h:3-285; serial => 1 core) on a bounded sample.  Honours
    steps / warm; if the projected time exceeds budget_s the call count stays and the remaining calls are skipped
    (reported in `calls`)."""
    from oracle import oracle_py as O
    (m, n, rp, ci, v), what = cpu_sample(wl, args)
    ora = O.Oracle("f64")
    t0 = time.time()
    M = ora.tile_create(m, n, rp, ci, v)  # conversion by the O(nnz log nnz) port (untimed set-up)
    t_conv = time.time() - t0
    x = np.random.default_rng(1).uniform(-1, 1, n)
    times = []
    t_start = time.time()
    if O.ref_available("f64"):
        ref, kind = O.Reference("f64"), "reference"
        with _quiet_stdout():  # the reference prints an errcount line per call
            for _ in range(steps + warm):
                ms, _ = ref.time_tilespmv_cpu(M, m, n, rp, ci, v, x)
                times.append(ms)
                if time.time() - t_start > budget_s and len(times) > warm:
                    break
    else:
        kind = "port"
        for _ in range(steps + warm):
            ms, _ = ora.time_tilespmv_cpu(M, m, n, x)
            times.append(ms)
            if time.time() - t_start > budget_s and len(times) > warm:
                break
    nnz = int(rp[m])
    timed = times[warm:] if len(times) > warm else times
    best, mean = min(timed), float(np.mean(timed))
    t0 = time.time()
    ora.csr_spmv(m, rp, ci, v, x)  # context only: the plain serial CSR loop of main.cu:101-110 on the same sample
    t_csr = time.time() - t0
    return {"value": 2.0 * nnz / (best * 1e-3) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "plain_csr_serial_gflops": 2.0 * nnz / t_csr / 1e9,
            "sample": f"{what}, tilespmv_cpu whole call, best of {len(timed)}; Tile_matrix built by the oracle port in {t_conv:.1f}s "
                      f"with {ora.threads()} threads (untimed)",
            "ms_per_call": best, "mean_ms_per_call": mean, "calls": len(timed), "nnz_sample": nnz,
            "host_threads_available": ora.threads(), "cpu_model": cpu_model()}


def conversion_baseline(gpu=True):
    """Tile_create (csr2tile.h:629-1020) of the reference with all host threads beside the GPU conversion, on BASELINE
    config 1 (2-D 5-pt Laplacian 1024^2: the reference's own CPU-runnable case; its O(tilem*tilen) scratch makes config
    2 a minute-long call).  Conversion bytes = CSR in + Tile_matrix out."""
    from oracle import oracle_py as O
    from tilespmv_b200 import generators as g
    m, n, rp, ci, v = g.lap2d(1024, val_mode=0)
    out = {"workload": "BASELINE config 1: 2-D 5-pt Laplacian 1024^2 fp64 (5.24 M nnz)"}
    if O.ref_available("f64"):
        ref = O.Reference("f64")
        with _quiet_stdout():
            t0 = time.time()
            ref.tile_create(m, n, rp, ci, v)  # leaked on purpose: the reference's Tile_destroy does not free everything
            out["reference_tile_create_s"] = time.time() - t0
        out["reference_threads"] = ref.threads()
        out["kind"] = "reference"
    else:
        ora = O.Oracle("f64")
        t0 = time.time()
        M = ora.tile_create(m, n, rp, ci, v)
        out["reference_tile_create_s"] = time.time() - t0
        ora.tile_destroy(M)
        out["reference_threads"] = ora.threads()
        out["kind"] = "port"
    if gpu:
        import torch
        from tilespmv_b200 import api
        d_rp, d_ci, d_v = torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda(), torch.from_numpy(v).cuda()
        best_h, best_d = 1e9, 1e9
        for _ in range(3):
            t0 = time.time()
            dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)
            torch.cuda.synchronize()
            best_h = min(best_h, time.time() - t0)
            dbytes = dm.info().device_bytes
            dm.destroy()
            t0 = time.time()
            dm = api.DeviceTileMatrix.from_csr(m, n, d_rp.data_ptr(), d_ci.data_ptr(), d_v.data_ptr(), on_device=True, precision=api.F64)
            torch.cuda.synchronize()
            best_d = min(best_d, time.time() - t0)
            dm.destroy()
        csr_bytes = rp.nbytes + ci.nbytes + v.nbytes
        out.update({"gpu_convert_s_host_csr": best_h, "gpu_convert_s_device_csr": best_d,
                    "conversion_bytes": int(csr_bytes + dbytes),
                    "gpu_GBps_on_conversion_bytes": (csr_bytes + dbytes) / best_d / 1e9,
                    "speedup_vs_reference": out["reference_tile_create_s"] / best_h,
                    "speedup_vs_reference_device_csr": out["reference_tile_create_s"] / best_d,
                    "note": "bytes = CSR read once + every Tile_matrix array written once; the GPU conversion is a sort + 13 "
                            "scans + scatter, i.e. several passes over those bytes, so this is a lower bound on its traffic"})
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    world = max(world, args.gpus)  # the workload follows --gpus even when the arm is started without torchrun
    wl = pick_workload(args, world)
    cb = cpu_baseline(wl, args, steps=args.steps, warm=args.warmup)
    value = 2.0 * cb["nnz_sample"] / (cb["mean_ms_per_call"] * 1e-3) / 1e9
    cb = dict(cb, value=value)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": cb["calls"],
           "warmup": args.warmup, "ms_per_step": cb["mean_ms_per_call"], "higher_is_better": True,
           "scaling": "weak" if wl == "c2" else "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args, world),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------------------
# our arm, one GPU: BASELINE config 2, one SpMV per step
# ------------------------------------------------------------------------------------------------
def torch_csr(rp, ci, v, m, n):
    import torch
    return torch.sparse_csr_tensor(torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda(), torch.from_numpy(v).cuda(), size=(m, n))


def run_single(args, local_rank):
    import torch
    from tilespmv_b200 import _capi, api, generators as g

    torch.cuda.set_device(local_rank)
    L = _capi.load()
    G = args.grid
    # first contact with the device (context, module load) is not part of any conversion figure
    t0 = time.time()
    wm = g.lap3d27(16, val_mode=0)
    api.DeviceTileMatrix.from_csr(*wm).destroy()
    torch.cuda.synchronize()
    t_first = time.time() - t0
    t0 = time.time()
    m, n, rp, ci, v = g.lap3d27(G, val_mode=0)
    nnz = int(rp[m])
    t_gen = time.time() - t0
    t0 = time.time()
    dm = api.DeviceTileMatrix.from_csr(m, n, rp, ci, v)  # GPU csr2tile (incl. H2D of the CSR)
    torch.cuda.synchronize()
    t_conv = time.time() - t0
    t0 = time.time()
    plan = api.Plan(dm, chunk_bytes=args.chunk_bytes, xstage_bytes=args.xstage_bytes, ctas_per_sm=args.ctas_per_sm)
    torch.cuda.synchronize()
    t_plan = time.time() - t0
    pi, di = plan.info(), dm.info()

    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1)
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    # sanity: the kernel's y against torch's own CSR SpMV on the same device data
    plan.spmv(x.data_ptr(), y.data_ptr(), stream)
    A = torch_csr(rp, ci, v, m, n)
    y_chk = A @ x
    scale = torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), A.values().abs(), size=(m, n)) @ x.abs()
    ok = bool(((y - y_chk).abs() <= 1e-12 * scale.clamp_min(1e-300)).all())
    del A, y_chk, scale
    if not ok:
        raise SystemExit("bench: SpMV result check failed")

    sampler = ClockSampler(gpu_index(local_rank))
    for _ in range(args.warmup):
        plan.spmv(x.data_ptr(), y.data_ptr(), stream)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = L.tilespmv_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.spmv(x.data_ptr(), y.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    launches = L.tilespmv_kernel_launch_count() - launches0
    ms_step = e0.elapsed_time(e1) / args.steps
    value = 2.0 * nnz / (ms_step * 1e-3) / 1e9

    # per-launch distribution (outside the timed region): 200 launches, one CUDA-event pair each
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for a_, b_ in evs:
        a_.record()
        plan.spmv(x.data_ptr(), y.data_ptr(), stream)
        b_.record()
    torch.cuda.synchronize()
    us = np.sort(np.array([a_.elapsed_time(b_) * 1e3 for a_, b_ in evs]))
    per_launch = {"n": 200, "min": float(us[0]), "median": float(us[100]), "p95": float(us[189]), "max": float(us[-1]),
                  "unit": "us", "note": "event pair around every launch: includes ~2 us of launch gap"}

    # ---- end-to-end with HOST buffers (pinned), host<->device copies inside the timed region ----
    # one batch call of >= 64 vectors (the pipeline's fill and drain are inside the timed call; measured: 0.740 ms per step
    # with 20 vectors, 0.743 with 64 -- PCIe at ~44 GB/s each way is the limit either way)
    e2e_steps = min(max(args.steps, 64), 256)
    yh = torch.empty(m, dtype=torch.float64).pin_memory()
    xh = torch.empty(n, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    for _ in range(2):
        _capi.check(L.tilespmv_plan_spmv_host(plan.handle, xh.data_ptr(), yh.data_ptr()))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _capi.check(L.tilespmv_plan_spmv_host(plan.handle, xh.data_ptr(), yh.data_ptr()))
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    # the pipelined batch call: every step still copies ITS x host->device and ITS y device->host, but step i+1's
    # H2D, step i's kernel and step i-1's D2H overlap (PCIe is full duplex); ring of 4 distinct pinned buffers
    ring = 4
    xring = [xh] + [xh.clone().pin_memory() for _ in range(ring - 1)]
    yring = [yh] + [torch.empty(m, dtype=torch.float64).pin_memory() for _ in range(ring - 1)]
    xp = [xring[i % ring].data_ptr() for i in range(e2e_steps)]
    yp = [yring[i % ring].data_ptr() for i in range(e2e_steps)]
    plan.spmv_host_batch(xp, yp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.spmv_host_batch(xp, yp)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_value = 2.0 * nnz / (e2e_ms * 1e-3) / 1e9
    e2e_ok = bool(torch.equal(yh.cuda(), y))
    clocks = sampler.stop()

    peak, peak_src = measured_peak()
    b_alg = pi.algorithmic_bytes
    achieved = b_alg / (ms_step * 1e-3) / 1e9
    cb = conv = None
    if not args.no_cpu_baseline:
        cb = cpu_baseline("c2", args)
        conv = conversion_baseline()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config("c2", args, 1),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": m * 8,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": f"tilespmv_plan_spmv_host_batch: {e2e_steps} host vectors per call (pinned host x -> host y each), "
                       "3-stream pipeline over a ring of 3 device buffers",
                "matches_device_y": e2e_ok, "serial_ms_per_step": e2e_serial_ms,
                "serial_api": "tilespmv_plan_spmv_host, one vector per call"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic("c2_lap3d27_160"),
                     "traffic_source": "RECORDED: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu "
                                       "--set full capture of this kernel on config 2, N = 1 (profiles/traffic.json); not re-measured by this run",
                     "peak_source": peak_src, "kernel": "tsp::tile_spmv_kernel<double>", "algorithmic_bytes_per_launch": b_alg,
                     "stream_bytes_per_launch": pi.stream_bytes, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "timing": "CUDA events on the launching stream over the timed steps / steps"},
        "cpu_baseline": cb,
        "conversion": conv,
        "extra": {"nnz_total": nnz, "rows_per_gpu": m, "tilenum": di.tilenum, "nnz_side": di.nnz_side,
                  "tiles_by_format": list(di.tiles_by_format), "chunks": pi.nchunks, "split_rows": pi.split_rows,
                  "grid": pi.grid, "block": pi.block, "smem_bytes": pi.smem_bytes, "chunk_bytes": pi.chunk_bytes,
                  "xstage_bytes": pi.xstage_bytes, "launches_per_spmv": pi.launches_per_spmv,
                  "csr_bytes": pi.csr_bytes, "gen_s": t_gen, "convert_s_incl_h2d": t_conv, "plan_s": t_plan,
                  "first_call_s": t_first, "first_call_note": "CUDA context + module load + a 16^3 conversion, before any timed figure",
                  "result_check_vs_torch_csr": ok, "library": os.path.basename(_capi.lib_path()),
                  "per_launch_us": per_launch},
    }
    emit(out)


# ------------------------------------------------------------------------------------------------
# our arm, N > 1 GPUs: x <- A*x with the per-iteration all-gather, strong scaling
# ------------------------------------------------------------------------------------------------
def run_multi(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from tilespmv_b200 import _capi, api, distributed as D, generators as g

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _capi.load()
    wl = pick_workload(args, world)
    n, rows = workload_rows(wl, args, world)
    r0, r1 = rows[rank]
    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(v):
        t = torch.tensor([v], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    t0 = time.time()
    m, _, rp, ci, v = workload_gen(wl, n, r0, r1)
    nnz_local = int(rp[m])
    t_gen = time.time() - t0
    comm = D.Comm(D.job_name("bench"), rank, world, nccl=True)
    t0 = time.time()
    sp = D.ShardedSpMV(comm, rows, n, rp, ci, v, plan_kwargs=dict(chunk_bytes=args.chunk_bytes, xstage_bytes=args.xstage_bytes))
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    pi, di, info = sp.plan.info(), sp.dm.info(), sp.info()
    nnz_total = allsum_int(nnz_local)

    gen = torch.Generator(device="cuda").manual_seed(1234)
    x0 = (torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1) / 8.0  # the same on every rank
    y = torch.empty(max(m, 1), dtype=torch.float64, device="cuda")

    def result(ptr):
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        _d2d(out.data_ptr(), ptr, n * 8)
        return out

    # ---------------- verification (before anything is timed) ----------------
    KV = 3
    A = torch_csr(rp, ci, v, m, n)
    Aabs = torch.sparse_csr_tensor(A.crow_indices(), A.col_indices(), A.values().abs(), size=(m, n))
    sp.spmv(x0.data_ptr(), y.data_ptr(), stream)
    y_chk = A @ x0
    bound = Aabs @ x0.abs()
    ok_spmv = bool(((y[:m] - y_chk).abs() <= 1e-12 * bound.clamp_min(1e-300)).all())
    # torch's own loop: local CSR SpMV + torch all-gather (equal row blocks except possibly the last)
    xr, br = x0.clone(), x0.abs()
    for _ in range(KV):
        parts = [torch.empty(b - a, dtype=torch.float64, device="cuda") for a, b in rows]
        dist.all_gather(parts, A @ xr)
        bparts = [torch.empty(b - a, dtype=torch.float64, device="cuda") for a, b in rows]
        dist.all_gather(bparts, Aabs @ br)
        xr, br = torch.cat(parts), torch.cat(bparts)
    del A, Aabs, y_chk, bound, parts, bparts
    verify, xs = {}, {}
    modes = MODES if (args.with_halo or args.exchange == "halo") else DEFAULT_MODES
    for mode in modes:
        try:
            xs[mode] = result(sp.iterate(x0.data_ptr(), KV, mode=mode, stream=stream))
            sp.sync(stream)
            err = float(((xs[mode] - xr).abs() / br.clamp_min(1e-300)).max())
            same = xs[mode].clone()
            dist.broadcast(same, src=0)
            verify[mode] = {"max_err_over_bound": err, "ok": err <= 1e-10, "identical_on_all_ranks": bool(torch.equal(same, xs[mode]))}
        except Exception as e:
            verify[mode] = {"ok": False, "error": str(e)[:300]}
    ref_mode = next((mo for mo in modes if mo in xs), None)
    for mode in xs:
        verify[mode]["bitwise_equal_to_" + ref_mode] = bool(torch.equal(xs[mode], xs[ref_mode]))
    for mode in verify:  # a rank-local failure fails the mode everywhere
        flag = torch.tensor([1 if (verify[mode].get("ok") and verify[mode].get("identical_on_all_ranks")
                                   and verify[mode].get("bitwise_equal_to_" + str(ref_mode), True)) else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verify[mode]["verified"] = bool(flag.item())
    del xs, xr, br
    allv = [None] * world
    dist.all_gather_object(allv, {mo: {k: v for k, v in verify[mo].items() if k in ("error", "ok", "identical_on_all_ranks")} for mo in verify})
    for mo in verify:
        bad = {r: allv[r][mo] for r in range(world) if allv[r][mo].get("error") or not allv[r][mo].get("ok")}
        if bad:
            verify[mo]["failing_ranks"] = {str(r): v for r, v in bad.items()}
    good = [mo for mo in modes if verify[mo].get("verified")]
    if not good or not ok_spmv:
        if rank == 0:
            sys.stderr.write(f"bench: verification failed: spmv {ok_spmv}, {json.dumps(verify)}\n")
        raise SystemExit(2)
    # every verified exchange is timed over the same K steps; `value` is the fastest one (--exchange forces one)
    headline = args.exchange if args.exchange in good else None

    # ---------------- timing ----------------
    def run_iters(mode, k):
        left = k
        while left > 0:
            seg = min(left, SEGMENT)
            sp.iterate(x0.data_ptr(), seg, mode=mode, stream=stream)
            left -= seg

    sampler = ClockSampler(gpu_index(local_rank))
    iterate, launches = {}, 0
    nvl = None
    if rank == 0:
        sampler.start()
        nvl = nvlink_kib(gpu_index(local_rank))
    for mode in good:
        steps = args.steps
        run_iters(mode, args.warmup)
        sp.sync(stream)
        barrier()
        l0 = L.tilespmv_kernel_launch_count()
        e0.record()
        run_iters(mode, steps)
        e1.record()
        sp.sync(stream)
        barrier()
        ms = allmax(e0.elapsed_time(e1) / steps)
        iterate[mode] = {"ms_per_iteration": ms, "value": 2.0 * nnz_total / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
                         "verified": True, "gpu_launches": int(L.tilespmv_kernel_launch_count() - l0)}
    if headline is None:
        headline = min(good, key=lambda mo: iterate[mo]["ms_per_iteration"])
    launches = iterate[headline]["gpu_launches"]
    if rank == 0 and nvl is not None:
        nv1 = nvlink_kib(gpu_index(local_rank))
        nvl = {"tx_bytes_total": (nv1[0] - nvl[0]) * 1024.0, "rx_bytes_total": (nv1[1] - nvl[1]) * 1024.0,
               "source": "nvidia-smi nvlink -gt d on rank 0's GPU before / after the timed loops of all exchanges"} if nv1 else None
    clocks = sampler.stop() if rank == 0 else None
    for mode in verify:
        if mode not in iterate:
            iterate[mode] = {"verified": False, "why": verify[mode]}
    # the SpMV alone (no exchange): what the kernel does on this rank's shard
    for _ in range(max(3, args.warmup // 4)):
        sp.spmv(x0.data_ptr(), y.data_ptr(), stream)
    barrier()
    k_spmv = max(5, min(args.steps, 100))
    e0.record()
    for _ in range(k_spmv):
        sp.spmv(x0.data_ptr(), y.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms_kernel_local = e0.elapsed_time(e1) / k_spmv
    ms_spmv = allmax(ms_kernel_local)

    # ---------------- end to end with HOST buffers: every rank uploads its slice of x, the slices are all-gathered
    # over NVLink, SpMV, D2H of the rank's y slice; 3 streams, ring of 3 device buffers ----------------
    e2e_steps = max(3, min(args.steps, 30))
    e2e = None
    if sp.info().equal_slices:
        R, ring = 3, 4
        xring = [torch.empty(m, dtype=torch.float64).pin_memory() for _ in range(ring)]
        for t in xring:
            t.copy_(x0[r0:r1].cpu())
        yring = [torch.empty(m, dtype=torch.float64).pin_memory() for _ in range(ring)]
        x_e = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(R)]
        y_e = [torch.empty(m, dtype=torch.float64, device="cuda") for _ in range(R)]
        s_in, s_comp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev_in, ev_comp, ev_out = ([torch.cuda.Event() for _ in range(R)] for _ in range(3))

        def e2e_batch(nsteps):
            for i in range(nsteps):
                b = i % R
                mine = x_e[b][r0:r1]
                with torch.cuda.stream(s_in):
                    if i >= R:
                        s_in.wait_event(ev_comp[b])
                    mine.copy_(xring[i % ring], non_blocking=True)
                    ev_in[b].record(s_in)
                with torch.cuda.stream(s_comp):
                    s_comp.wait_event(ev_in[b])
                    if i >= R:
                        s_comp.wait_event(ev_out[b])
                    dist.all_gather_into_tensor(x_e[b], mine)
                    sp.spmv(x_e[b].data_ptr(), y_e[b].data_ptr(), s_comp.cuda_stream)
                    ev_comp[b].record(s_comp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_comp[b])
                    yring[i % ring].copy_(y_e[b], non_blocking=True)
                    ev_out[b].record(s_out)
            torch.cuda.synchronize()

        e2e_batch(e2e_steps)
        barrier()
        t0 = time.perf_counter()
        e2e_batch(e2e_steps)
        barrier()
        e2e_ms = allmax((time.perf_counter() - t0) * 1e3 / e2e_steps)
        sp.spmv(x0.data_ptr(), y.data_ptr(), stream)
        torch.cuda.synchronize()
        e2e_ok = bool(torch.equal(yring[(e2e_steps - 1) % ring].cuda(), y[:m]))
        e2e = {"value": 2.0 * nnz_total / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
               "ms_per_step": e2e_ms, "steps": e2e_steps, "matches_device_y": e2e_ok,
               "api": f"per rank, {e2e_steps} steps pipelined over 3 streams: H2D of its x slice (pinned) -> NCCL all-gather of x over "
                      "NVLink + tilespmv_plan_spmv on its row block -> D2H of its y slice (bytes are the sums over all ranks)"}
        del x_e, y_e, xring, yring

    # ---------------- the same matrix on ONE GPU: rank 0 builds every row-block plan (default single-GPU options) and
    # runs them back to back; the other ranks wait ----------------
    one_gpu = None
    if not args.no_one_gpu:
        if rank == 0:
            t0 = time.time()
            plans = []
            for (a, b) in rows:
                mm, _, rp2, ci2, v2 = workload_gen(wl, n, a, b)
                dm2 = api.DeviceTileMatrix.from_csr(mm, n, rp2, ci2, v2)
                plans.append((a, api.Plan(dm2, chunk_bytes=args.chunk_bytes, xstage_bytes=args.xstage_bytes)))
                dm2.destroy()
                del rp2, ci2, v2
            torch.cuda.synchronize()
            t_build = time.time() - t0
            xa, xb = x0.clone(), torch.empty_like(x0)

            def one_iter(src, dst):
                for a, p in plans:
                    p.spmv(src.data_ptr(), dst.data_ptr() + a * 8, stream)

            k1 = max(3, min(args.steps, 20))
            for _ in range(2):
                one_iter(xa, xb)
            torch.cuda.synchronize()
            e0.record()
            for i in range(k1):
                one_iter(xa, xb) if i % 2 == 0 else one_iter(xb, xa)
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / k1
            # its result after KV steps from x0 equals the distributed one?
            xa.copy_(x0)
            for i in range(KV):
                one_iter(xa, xb) if i % 2 == 0 else one_iter(xb, xa)
            x1 = xb if KV % 2 == 1 else xa
            one_gpu = {"ms_per_iteration": ms1, "value": 2.0 * nnz_total / (ms1 * 1e-3) / 1e9, "unit": UNIT, "steps": k1,
                       "launches_per_iteration": int(sum(p.info().launches_per_spmv for _, p in plans)), "build_s": t_build,
                       "checksum_abs_after_%d_steps" % KV: float(x1.abs().sum()),
                       "how": f"rank 0's GPU runs the {world} row-block plans of the SAME global matrix back to back (x, y local; no exchange needed)"}
            for _, p in plans:
                p.destroy()
            del plans, xa, xb, x1
        barrier()

    # ---------------- the weak-scaling stencil of round 1 (one 160^3 slab per GPU), as an extra ----------------
    weak = None
    if not args.no_weak:
        Gd = args.grid
        wm, wn, wrp, wci, wv = g.lap3d27_slab(Gd * world, Gd, Gd, rank * Gd, (rank + 1) * Gd, val_mode=0)
        wrows = [(r * wm, (r + 1) * wm) for r in range(world)]
        ws = D.ShardedSpMV(comm, wrows, wn, wrp, wci, wv)
        wnnz = allsum_int(int(wrp[wm]))
        wx0 = (torch.rand(wn, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1) / 32.0
        wy = torch.empty(wm, dtype=torch.float64, device="cuda")
        weak = {"workload": f"3-D 27-pt Laplacian {Gd * world}x{Gd}x{Gd}: one {Gd}^3 slab per GPU (round 1's weak-scaling line)", "nnz_total": wnnz}
        kw = max(5, min(args.steps, 100))
        for _ in range(5):
            ws.spmv(wx0.data_ptr(), wy.data_ptr(), stream)
        barrier()
        e0.record()
        for _ in range(kw):
            ws.spmv(wx0.data_ptr(), wy.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        msw = allmax(e0.elapsed_time(e1) / kw)
        weak["spmv_no_exchange"] = {"ms_per_step": msw, "value": 2.0 * wnnz / (msw * 1e-3) / 1e9}
        for mode in good:
            ws.iterate(wx0.data_ptr(), 5, mode=mode, stream=stream)
            ws.sync(stream)
            barrier()
            e0.record()
            left = kw
            while left > 0:
                ws.iterate(wx0.data_ptr(), min(left, 20), mode=mode, stream=stream)
                left -= 20
            e1.record()
            ws.sync(stream)
            barrier()
            msm = allmax(e0.elapsed_time(e1) / kw)
            weak["iterate_" + mode] = {"ms_per_iteration": msm, "value": 2.0 * wnnz / (msm * 1e-3) / 1e9}
        ws.destroy()
        del wx0, wy

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = iterate[headline]["ms_per_iteration"]
        value = iterate[headline]["value"]
        # B_alg of SURVEY 8(d) for rank 0's shard; a row-block shard reads only the x columns its launches touch, but the
        # formula's s*n for x is kept (conservative: it is what a shard of a matrix without structure reads)
        b_alg = pi.algorithmic_bytes
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": iterate[headline]["steps"], "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_config(wl, args, world), headline_exchange=headline),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": b_alg / (ms_kernel_local * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_alg / (ms_kernel_local * 1e-3) / 1e9 / peak, "traffic": None,
                         "traffic_note": "no ncu capture of a multi-rank run (the profiling recipe forbids it); see profiles/ for the 1-GPU captures",
                         "peak_source": peak_src, "kernel": "tsp::tile_spmv_kernel<double> on rank 0's row block (all launch units of one SpMV)",
                         "algorithmic_bytes_per_launch": b_alg, "stream_bytes_per_launch": pi.stream_bytes,
                         "launches_per_spmv": pi.launches_per_spmv,
                         "timing": "CUDA events around back-to-back SpMVs of rank 0's shard WITHOUT the exchange (kernel alone)",
                         "achieved_inside_iterate": b_alg / (ms_step * 1e-3) / 1e9,
                         "frac_inside_iterate": b_alg / (ms_step * 1e-3) / 1e9 / peak,
                         "frac_of_nominal_8TBs": b_alg / (ms_kernel_local * 1e-3) / 1e9 / 8000.0,
                         "gather": ({"side_entries_rank0": int(di.nnz_side),
                                     "achieved_G_per_s": di.nnz_side / (ms_kernel_local * 1e-3) / 1e9,
                                     "probe_ceiling_G_per_s": GATHER_PROBE_GPS,
                                     "frac_of_probe": di.nnz_side / (ms_kernel_local * 1e-3) / 1e9 / GATHER_PROBE_GPS,
                                     "note": "this row block is made of extracted (side) entries: one scattered 8-byte load of x per nonzero; "
                                             "the SM's L1 accepts ~0.92 such requests per clock (tools/gather_probe.cu), which binds before HBM does"}
                                    if 2 * di.nnz_side > nnz_local else None)},
            "cpu_baseline": None,
            "iterate": dict(iterate, verification=verify,
                            collective={"nccl": "one in-place ncclAllGather per iteration (library-owned communicator)",
                                        "fused": "SpMV epilogue stores y into every peer's next x (P2P over NVLink) + 1 flag barrier",
                                        "pipelined": "copy-engine pushes of the y slice to each peer in the order of need + per-launch waits on "
                                                     "only the slices a launch reads (x panels cut at the ranks' row blocks)",
                                        "halo": ("fused stores of only the rows a peer's next launch reads + copy-engine replication of the rest of "
                                                 "every slice in the background" if info.halo_eligible else "not eligible for this matrix: ran as pipelined"),
                                        "allgather_bytes_in_per_gpu": (n - m) * 8, "allgather_bytes_out_per_gpu_unicast": (world - 1) * m * 8,
                                        "nvlink_floor_ms_unoverlapped": (n - m) * 8 / (NVLINK_PEER_GBS * 1e6)}),
            "spmv_no_exchange": {"ms_per_step": ms_spmv, "value": 2.0 * nnz_total / (ms_spmv * 1e-3) / 1e9,
                                 "note": "all ranks run their row block, no all-gather: upper bound of the loop"},
            "one_gpu": one_gpu,
            "speedup_vs_one_gpu": (one_gpu["ms_per_iteration"] / ms_step) if one_gpu else None,
            "nvlink": nvl, "weak_scaling_stencil": weak,
            "extra": {"nnz_total": nnz_total, "rows_rank0": m, "tilenum_rank0": di.tilenum, "nnz_side_rank0": di.nnz_side,
                      "tiles_by_format_rank0": list(di.tiles_by_format), "chunks": pi.nchunks, "split_rows": pi.split_rows,
                      "grid": pi.grid, "block": pi.block, "smem_bytes": pi.smem_bytes, "chunk_bytes": pi.chunk_bytes,
                      "launch_units": info.launch_units, "unit_deps_rank0": [int(info.unit_deps[u]) for u in range(info.launch_units)],
                      "xpanels": pi.xpanels, "gen_s": t_gen, "convert_plan_s": t_setup, "spmv_check_vs_torch_csr": ok_spmv,
                      "nccl_version": list(torch.cuda.nccl.version()), "library": os.path.basename(_capi.lib_path())},
        }
    barrier()
    sp.destroy()
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        emit(out)


def _d2d(dst, src, nbytes):
    import ctypes
    global _RT
    try:
        _RT
    except NameError:
        _RT = ctypes.CDLL("libcudart.so.12")
        _RT.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rc = _RT.cudaMemcpy(dst, src, nbytes, 3)
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy D2D failed: {rc}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)   # BENCH_REPEAT of the reference (common.h:16-18)
    ap.add_argument("--warmup", type=int, default=200)   # WARMUP_NUM (common.h:20-22)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c3", "c5"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "pipelined", "halo", "fused", "nccl"])
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--c3-rows", type=int, default=8_000_000)
    ap.add_argument("--c5-rows", type=int, default=50_000_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--chunk-bytes", type=int, default=0)
    ap.add_argument("--xstage-bytes", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-one-gpu", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--with-halo", action="store_true", help="also verify / time the halo exchange")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif world == 1:
        if pick_workload(args, 1) != "c2":
            raise SystemExit("bench: workloads c3 / c5 are the multi-GPU lines (launch with torchrun, --gpus N)")
        run_single(args, local_rank)
    else:
        run_multi(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
