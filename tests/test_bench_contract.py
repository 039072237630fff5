"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU path on the whole config-2
matrix) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["metric"].startswith("fp64 SpMV GFLOP/s") and d["unit"] == "GFLOP/s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert "Laplacian 160^3" in d["config"]["workload"] and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    assert cb["cores"] == 1 and cb["value"] == d["value"] and "109215352 nnz" in cb["sample"]
    assert cb["kind"] == ("reference" if O.ref_available("f64") else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.timeout(300)
def test_reference_arm_follows_the_multi_gpu_workload_and_honours_steps():
    """--gpus 2 / 4 name BASELINE config 3, --gpus 8 config 5 (strong scaling); the CPU arm times a bounded sample of that
    workload and runs exactly --steps timed calls after --warmup untimed ones."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "4", "--steps", "3", "--warmup", "2",
                        "--c3-rows", "400000", "--cpu-sample-rows", "100000"], capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["steps"] == 3 and d["warmup"] == 2 and d["scaling"] == "strong" and d["n_gpus"] == 4
    assert "config 3" in d["config"]["workload"] and "first 100000 rows" in d["cpu_baseline"]["sample"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0",
                        "--c5-rows", "1000000", "--cpu-sample-rows", "50000"], capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert "config 5" in d["config"]["workload"] and d["cpu_baseline"]["nnz_sample"] == 50000 * 20
