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
    d = _run(["--steps", "12", "--warmup", "3", "--ncell", "24", "--cpu-ncell", "8", "--ref-ncell-serial", "6", "--cpu-steps", "2",
              "--e2e-steps", "3", "--no-strong"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "roofline_hbm", "roofline_fp64"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 12 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["gpu_launches"] > 30 and d["value"] > 1e7
    rl, rf, rh = d["roofline"], d["roofline_fp64"], d["roofline_hbm"]
    # the binding resource first: the L1TEX data pipe (gather wavefronts per second against one per clock and SM); the FP64 view
    # against the measured DFMA rate and the HBM view beside it
    assert rl["bound"] == "l1tex" and abs(rl["frac"] - rl["achieved"] / rl["peak"]) < 1e-12 and 0 < rl["frac"] < 1.2
    assert {"traffic", "share_of_step", "hbm_frac", "fp64_frac"} <= set(rl)
    assert rf["bound"] == "fp64" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert 60 < rf["mean_neighbours_listed"] < 110 and 40 < rf["mean_neighbours_in_range"] < rf["mean_neighbours_listed"]
    assert abs(rf["algorithmic_flops_per_atom"] - (21 * rf["mean_neighbours_listed"] + 18 * rf["mean_neighbours_in_range"])) < 1e-6 + 20
    assert rh["bound"] == "hbm" and rh["unit"] == "GB/s" and abs(rh["frac"] - rh["achieved"] / rh["peak"]) < 1e-12
    assert d["e2e"]["h2d_bytes_per_step"] == 72 * d["config"]["n_atoms"] and d["e2e"]["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_ignores_torchrun_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must still use the host's cores (round 1's
    N > 1 reference numbers were single-thread runs because of it)."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-ncell", "8"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-1000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["sample_atoms"] == 2048 and d["sample_is_whole_workload"] is False and "2048-atom" in d["cpu_baseline"]["sample"]


def test_reference_arm_times_the_real_workload_when_it_fits():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ncell", "8"])
    assert d["sample_atoms"] == d["config"]["n_atoms"] == 2048 and d["sample_is_whole_workload"] is True


def test_fp64_yardstick_arithmetic():
    """`roofline` of the CUDA arm: algorithmic flops (21 K + 18 K_in + 20) against 2 x the measured DFMA rate,
    and the issued-slot fraction (SASS count x listed pairs) against the DFMA rate."""
    sys.path.insert(0, ROOT)
    import bench

    r = bench.fp64_roofline("k", 85.34, 54.0, 4_000_000, 1.0, True)
    assert abs(r["algorithmic_flops_per_atom"] - (21 * 85.34 + 18 * 54.0 + 20)) < 1e-9
    assert abs(r["issued_fp64_slots_per_atom"] - (bench.FP64_SLOTS_PER_PAIR * 85.34 + bench.FP64_SLOTS_EPILOGUE)) < 1e-9
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and abs(r["frac_of_dfma_issue_rate"] - 2 * r["frac"]) < 1e-12
    assert r["bound"] == "fp64" and 0 < r["issue_frac"] < 1
    assert bench.fp64_roofline("k", 85.34, 54.0, 4_000_000, 0.8, True)["frac"] > r["frac"]


def test_l1tex_yardstick_arithmetic():
    """The primary `roofline`: listed pairs per second x the ncu-calibrated wavefronts per pair, against one data-pipe wavefront
    per clock and SM.  The committed calibration file names the capture it comes from."""
    sys.path.insert(0, ROOT)
    import bench

    cal, traffic, fp64_pct, src = bench.force_calibration("k_force_vv<fused>")
    assert cal and traffic > 1e9 and 0 < fp64_pct < 100 and "profiles/" in src
    assert 0.5 < cal["l1tex_wavefronts_per_listed_pair"] <= 1.0 and cal["l1tex_data_pipe_lsu_wavefronts_pct"] > 50
    assert bench.force_calibration("k_some_other_kernel") == ({}, None, None, None)
    r = bench.l1tex_roofline("k", 85.4, 4_000_000, 1.18, 1965.0, cal)
    assert r["bound"] == "l1tex" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["peak"] - 148 * 1.965) < 1e-9 and abs(r["listed_pairs_per_s"] - 85.4 * 4e6 / 1.18e-3) < 1.0
    assert 0.8 < r["frac"] < 1.0          # the ncu capture this is calibrated on says 90.9 %
