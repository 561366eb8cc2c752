"""bench.py contract: the reference arm runs on CPU; the CUDA arm's JSON line carries every required key (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_runs_on_cpu():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-ncell", "8", "--cpu-threads", "2"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "atom-steps/sec LJ argon NVE" and d["unit"] == "atom-steps/s" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4 and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_cuda_arm_json_line():
    d = _run(["--steps", "12", "--warmup", "3", "--ncell", "24", "--ref-ncell", "8", "--ref-ncell-serial", "6", "--cpu-steps", "2",
              "--e2e-steps", "3"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 12 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["gpu_launches"] > 30 and d["value"] > 1e7
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert d["e2e"]["h2d_bytes_per_step"] == 72 * d["config"]["n_atoms"] and d["e2e"]["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_fp64_yardstick_arithmetic():
    """roofline_fp64 of the CUDA arm: issue slots per atom x atoms / launch time against the measured DFMA rate."""
    sys.path.insert(0, ROOT)
    import bench

    r = bench.fp64_roofline(85.34, 4_000_000, 1.16, True)
    assert abs(r["slots_per_atom"] - (35 * 85.34 + 110)) < 1e-9
    assert 0.60 < r["frac"] < 0.70 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12     # ncu: 66 % for this launch
    assert bench.fp64_roofline(85.34, 4_000_000, 1.09, False)["frac"] > r["frac"]
