"""cli/tilespmv_test: the reference's `./test -d <id> A.mtx` flow (main.cu:15-205) in plain C over the C-ABI."""
import os
import subprocess

import numpy as np
import pytest

from tilespmv_b200 import generators as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cli", "tilespmv_test")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cli")])


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback: without a device the conversion reports TILESPMV_ERR_NODEVICE through the CLI."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    _build()
    m, n, rp, ci, v = g.lap2d(16, val_mode=1)
    path = str(tmp_path / "a.mtx")
    g.write_mtx_fast(path, m, n, rp, ci, v)
    r = subprocess.run([CLI, "-d", "0", path], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode == 0
    assert "Tile_create failed" in r.stdout and "no CPU fallback" in r.stdout
    assert "PASS" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("binary", ["tilespmv_test", "tilespmv_test_f32"])
def test_cli_matches_the_reference_driver_conventions(tmp_path, binary):
    _build()
    m, n, rp, ci, v = g.band_contig(1000, hb=18, val_mode=1)  # 1000 rows: the driver drops the last 8 (main.cu:71)
    path = str(tmp_path / "band.mtx")
    g.write_mtx_fast(path, m, n, rp, ci, v)
    env = dict(os.environ, TILESPMV_BENCH_REPEAT="20", TILESPMV_WARMUP_NUM="5")
    r = subprocess.run([os.path.join(ROOT, "cli", binary), "-d", "0", path], capture_output=True, text=True, cwd=tmp_path,
                       env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "The number of tile" in r.stdout
    assert "CUDA SpMV runtime" in r.stdout and "GFlops" in r.stdout
    assert "Check... PASS!" in r.stdout
    line = open(tmp_path / "results.csv").read().strip().split(",")
    assert line[0] == path and int(line[1]) == 992 and int(line[2]) == n  # filename,rowA,colA,nnzA,ms,gflops
    assert len(line) == 6 and float(line[4]) > 0
    # TILESPMV_CSV_EXTENDED=1 appends: algorithmic bytes, GB/s on them, fraction of the HBM peak
    r = subprocess.run([os.path.join(ROOT, "cli", binary), "-d", "0", path], capture_output=True, text=True, cwd=tmp_path,
                       env=dict(env, TILESPMV_CSV_EXTENDED="1", TILESPMV_HBM_PEAK_GBS="6556.8"), timeout=300)
    assert r.returncode == 0, r.stderr
    ext = open(tmp_path / "results.csv").read().strip().split("\n")[-1].split(",")
    assert len(ext) == 9 and ext[:4] == line[:4] and int(ext[6]) > 0
    assert abs(float(ext[7]) - int(ext[6]) * 1e-6 / float(ext[4])) < 1e-3 * float(ext[7]) and 0 < float(ext[8]) < 1.5


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["pipelined", "fused", "nccl"])
def test_multi_gpu_loop_from_a_plain_c_host(tmp_path, exchange):
    """cli/tilespmv_multi: one forked process per rank, everything through the C-ABI (tilespmv_comm_* / tilespmv_dist_*).
    With one GPU the ranks share it (-s); NCCL needs one GPU per rank."""
    import torch
    _build()
    ndev = torch.cuda.device_count()
    if exchange == "nccl" and ndev < 2:
        pytest.skip("NCCL needs one GPU per rank")
    m, n, rp, ci, v = g.banded(4096 + 32, val_mode=1)
    path = str(tmp_path / "banded.mtx")
    g.write_mtx_fast(path, m, n, rp, ci, v)
    cmd = [os.path.join(ROOT, "cli", "tilespmv_multi"), "-n", "2", "-k", "3", "-x", exchange] + (["-s"] if ndev < 2 else []) + [path]
    env = dict(os.environ, TILESPMV_COMM_SPIN_TIMEOUT_S="25", TILESPMV_COMM_TIMEOUT_S="60")
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp_path, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Check... PASS!" in r.stdout and f"({exchange} exchange)" in r.stdout
