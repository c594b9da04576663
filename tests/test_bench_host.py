"""CPU tests (-m "not gpu") of bench.py's host-side logic: the clock sampler's window filtering and the workload
constants the JSON line is computed from."""
import importlib.util
import math
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_reports_only_the_timed_window():
    b = load_bench()
    s = b.ClockSampler(0)
    s.max_mhz = 1965.0
    s.samples = [(1.0, 1200, 0), (2.0, 1965, 0), (2.5, 1950, 0x4), (3.0, 1965, 0), (9.0, 500, 0x8)]
    got = s.stop(1.5, 3.5)
    assert got == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "samples": 3, "samples_total": 5, "reasons": ["sw_power_cap"]}
    # a region shorter than one NVML round trip falls back to the sample closest to it
    s = b.ClockSampler(0)
    s.samples = [(1.0, 1200, 0), (9.0, 500, 0x8)]
    got = s.stop(1.2, 1.3)
    assert got["sm_mhz"] == 1200.0 and got["samples"] == 1 and got["reasons"] == []


def test_clock_sampler_without_a_driver_says_so():
    b = load_bench()
    s = b.ClockSampler(0)
    s.start()
    time.sleep(0.02)
    got = s.stop(0.0, 1e18)
    if got["samples"] == 0:  # no NVML here: reported, never raised
        assert got["sm_mhz"] is None and got["reasons"][0].startswith("unavailable")


def test_workload_constants_match_baseline_config():
    b = load_bench()
    assert b.N_FFT == 4096 and b.BATCH == 1 << 16  # BASELINE.json configs[1]
    assert b.FLOP_PER_STEP == 5.0 * 4096 * 12 * 65536
    assert b.ALGO_BYTES_PER_STEP == 2 * 4096 * 65536 * 8 == 4294967296  # one read + one write
    cfg = b.config(8)
    assert cfg["global_batch"] == 8 * 65536 and "workload" in cfg and "model" not in cfg
    assert math.isclose(b.measured_peaks()[0], b.measured_peaks()[0]) and b.measured_peaks()[0] > 1000
