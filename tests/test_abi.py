"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/genfft_cuda.h declares;
host-side argument checking works without a GPU; no compute is attempted."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import genfft_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "genfft_cuda.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(g.LIB_PATH):
        g.build()
    return g.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(genfft_cuda_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("genfft_cuda_plan_c2c_1d", "genfft_cuda_plan_r2c_1d", "genfft_cuda_plan_c2c_2d", "genfft_cuda_plan_vert",
                 "genfft_cuda_plan_dit", "genfft_cuda_exec_c2c", "genfft_cuda_exec_c2c_dev", "genfft_cuda_exec_r2c",
                 "genfft_cuda_exec_c2c_2d", "genfft_cuda_exec_vert", "genfft_cuda_exec_dit", "genfft_cuda_plan_destroy"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", g.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (genfft_cuda_[a-z0-9_]+)", out))
    declared = set(declared_symbols())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    # the python binding covers the same surface, with no torch types in any signature
    assert set(g.exported_symbols()) == declared
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)


def test_extern_c_linkage_only_plain_types():
    text = open(HEADER).read()
    assert 'extern "C"' in text
    for banned in ("torch", "at::", "std::", "cudaStream_t"):
        body = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        assert banned not in body, f"{banned} leaked into the C ABI"


def test_sm100a_only_binary():
    """The product is sm_100a code: the fat binary holds sm_100a SASS and nothing for other architectures."""
    out = subprocess.run(["cuobjdump", "-lelf", g.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_fails_loudly_or_runs(lib):
    """Without a usable sm_100 device plan creation must fail with a CUDA error -- there is no CPU fallback."""
    if g.device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(g.GenfftCudaError, match="error 3"):
        g.FFT(1024)
    with pytest.raises(g.GenfftCudaError):
        g.RealFFT(1024)
    with pytest.raises(g.GenfftCudaError):
        g.FFT2D(64, 64)


def test_argument_errors_before_any_cuda_call(lib):
    h = ctypes.c_void_p()
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 3, 1, 0, 0) == 1      # not a power of two -> ERR_SIZE
    assert b"unsupported size" in lib.genfft_cuda_last_error_string()
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 1 << 30, 1, 0, 0) == 1  # above the size cap
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 8, 0, 0, 0) == 2      # batch < 1 -> ERR_ARG
    assert lib.genfft_cuda_plan_c2c_1d(None, 0, 8, 1, 0, 0) == 2
    assert lib.genfft_cuda_plan_c2c_2d(ctypes.byref(h), 0, 12, 8) == 1
    assert lib.genfft_cuda_exec_c2c(None, None, None, 0) == 2
    assert lib.genfft_cuda_plan_destroy(None) == 0
    # round 2: distances of a batch must hold a whole transform, checked before any CUDA call
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 256, 4, 100, 0) == 2
    assert b"in_dist" in lib.genfft_cuda_last_error_string()
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 256, 4, 0, -256) == 2
    assert lib.genfft_cuda_plan_r2c_1d(ctypes.byref(h), 0, 256, 4, 1, 0, 100) == 2
    assert lib.genfft_cuda_plan_c2r_1d(ctypes.byref(h), 0, 256, 4, 0, 128) == 2
    assert lib.genfft_cuda_plan_c2c_1d(ctypes.byref(h), 0, 2, 1 << 40, 0, 0) == 1  # volume beyond the 32-bit tile counts
    # the new host-side entry points validate their arguments too
    out = ctypes.c_void_p()
    node = ctypes.c_int()
    assert lib.genfft_cuda_host_alloc(None, 16, 1, ctypes.byref(node)) == 2
    assert lib.genfft_cuda_host_alloc(ctypes.byref(out), 0, 1, ctypes.byref(node)) == 2
    assert lib.genfft_cuda_host_free(ctypes.c_void_p(4096)) == 2           # not one of its blocks
    assert lib.genfft_cuda_separate_2x_real(7, None, None, None, 8) == 2  # bad precision
    assert lib.genfft_cuda_separate_2x_real(0, None, None, None, 8) == 2  # null buffers
    peers = (ctypes.c_void_p * 8)()
    assert lib.genfft_cuda_scatter_cols_dev(0, peers, 3, 0, ctypes.c_void_p(4096), 64, 4, 64, None) == 2  # ranks not 2^k
    assert lib.genfft_cuda_scatter_cols_dev(0, None, 2, 0, ctypes.c_void_p(4096), 64, 4, 64, None) == 2
    assert lib.genfft_cuda_debug_mode_launch_count(99) == 0


def test_host_mirror_semantics_without_gpu():
    empty = g.FFT()
    assert not empty and empty.size() == 0           # default-constructed plan (fft.h:59,107-108)
    assert not g.FFT2D() and g.FFT2D().cols() == 0 and g.FFT2D().rows() == 0
    with pytest.raises(g.GenfftCudaError):
        empty.transform(np.zeros(4, np.complex64), np.zeros(4, np.complex64))


def test_product_never_imports_the_oracle():
    """genfft_b200/ must not reference oracle/ in any way (the oracle is the checker, never the product)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "genfft_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "genfft_ref" not in text and "oracle/" not in text, f


def test_tile_decode_division_is_exact():
    """the kernels decode tile -> (t0, t1, t2) with a magic-number multiply instead of an integer division
    (tile_kernel.cuh: fast_div); it must be exact for every tile index below 2^31 and every divisor a plan can
    produce (tile counts per row / block: powers of two, powers of two plus one for the pair tiles, batch counts)"""
    import random
    import genfft_b200 as g
    f = g.lib().genfft_cuda_debug_fast_div
    rng = random.Random(7)
    divisors = set(range(1, 1100)) | {(1 << k) + o for k in range(1, 31) for o in (-1, 0, 1)} | \
        {rng.randrange(1, 1 << 31) for _ in range(300)}
    divisors = {d for d in divisors if 1 <= d < (1 << 31)}
    top = (1 << 31) - 1
    for d in sorted(divisors):
        xs = {0, 1, d - 1, d, d + 1, top, top - 1, top // d * d, max(0, top // d * d - 1)}
        xs |= {rng.randrange(0, top + 1) for _ in range(40)}
        xs |= {min(top, k * d + o) for k in (2, 3, 1000, 65535, 65536) for o in (-1, 0, 1)}
        for x in xs:
            if 0 <= x <= top:
                assert f(x, d) == x // d, (x, d)


def test_bench_reference_arm_prints_one_json_line(checkers):
    """bench.py --impl reference (the CPU arm the driver runs beside ours) needs no GPU: exactly one line on stdout,
    the contract keys, and an e2e entry without host<->device bytes."""
    import json
    import sys
    if checkers[0] is None:
        pytest.skip("compiled reference not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-1000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None
