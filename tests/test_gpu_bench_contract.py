"""GPU test of the bench contract: bench.py prints ONE JSON line with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3",
                        *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), r.stdout[-2000:]  # nothing but the JSON line on stdout
    return json.loads(lines[0])


def test_bench_line_has_the_contract_keys():
    d = run_bench("--e2e-steps", "1", "--no-extras")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] == 3 and d["scaling"] == "weak"
    assert d["dtype"] == "f32" and "workload" in d["config"] and d["vs_baseline"] is None
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert 0.3 < rf["frac"] < 1.2, rf  # a fast kernel, but not faster than the memory system
    assert abs(d["value"] - 5 * 4096 * 12 * 65536 / (d["ms_per_step"] * 1e-3) / 1e9) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4096 * 65536 * 8 and e["d2h_bytes_per_step"] == 4096 * 65536 * 8
    assert 0 < e["value"] < d["value"]  # host copies inside the timed region
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0


def test_reference_arm_line():
    d = run_bench("--impl", "reference")
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "GFLOP/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
