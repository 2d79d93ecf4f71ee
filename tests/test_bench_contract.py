"""bench.py's reference arm runs here (CPU only) and prints the contract's JSON line; the GPU arm's keys are checked on the
GPU box by the same function."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
          "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _line(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_arm_contract():
    d = _line(["--impl", "reference", "--steps", "2", "--warmup", "1", "--no-rbpf"])
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "mppi_trajectory_steps_per_sec" and d["unit"] == "trajectory-steps/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("MPPI K=16384 T=64") and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert 1e5 < d["value"] < 1e8          # a single CPU thread: order 1e6 trajectory-steps/s


@pytest.mark.gpu
def test_gpu_arm_contract(gpu_pkg):
    d = _line(["--steps", "50", "--warmup", "3", "--rbpf-scans", "2", "--c4-steps", "20", "--c5-ticks", "20"])
    assert COMMON <= set(d) and {"roofline", "clocks", "rbpf", "parity_check", "c4", "c5"} <= set(d)
    pc = d["parity_check"]
    assert pc["kernel_variant"] == "fast" and max(pc["controls_rel_err"], pc["plan_rel_err"], pc["states_rel_err"]) < 1e-5
    assert d["c4"]["rollouts_per_gpu"] == 65536 and d["c4"]["kernel_variant"] == "fast" and d["c4"]["us_per_call"] > 0
    assert d["c5"]["particles_total"] == 4096 and d["c5"]["meets_50hz_budget"] is True
    assert d["dtype"] == "f64" and d["gpu_launches"] == 100 and d["n_gpus"] == 1
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["algorithmic_bytes_per_launch"] == 16384 * 64 * 12
    assert d["e2e"]["h2d_bytes_per_step"] == 24 and d["e2e"]["d2h_bytes_per_step"] == 16 and d["e2e"]["value"] <= d["value"] * 1.05
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    r = d["rbpf"]
    assert r["unit"] == "particle-updates/s" and r["roofline"]["kernel"].startswith("rbpf_distance_field") and r["cpu_baseline"]["cores"] == 1
