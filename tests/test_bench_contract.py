"""bench.py keeps the driver's JSON contract (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args, timeout=300):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                         text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = run_bench("--impl", "reference", "--workload", "C2", "--steps", "3", "--warmup", "1")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "decisions/s" and d["value"] > 0 and d["steps"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "requests per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "C2", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_contract():
    d = run_bench("--workload", "C3", "--steps", "20", "--warmup", "3", "--min-seconds", "0.05",
                  "--cpu-seconds", "0.5")
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["value"] > 1e8 and d["gpu_launches"] >= 1 and d["parity_checked"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["kernel"].startswith("lig_pick_queue_kernel") and d["gpu_launches"] == 1   # one merged launch
    assert rf["algorithmic_bytes_per_step"] == int(24 * 65536 + (16 * 512 + 4 * 256 * 16) / 20)
    assert rf["algorithmic_bytes_per_launch"] == rf["algorithmic_bytes_per_step"] * 20
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 16 * 65536 and e["d2h_bytes_per_step"] == 8 * 65536
    assert e["value"] < d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["optimised_cpu"]["value"] > cb["single_thread"]["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["streaming"]["errors"] == 0 and d["streaming"]["latency_us"]["p99"] > 0
