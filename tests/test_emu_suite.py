"""The -m gpu parity tests, run WITHOUT a GPU on the kernel-logic emulator (tests/emu).

tests/emu compiles the real kernel and plan sources (genfft_b200/csrc/*.cu, unchanged) with g++ against a fake CUDA
runtime: device memory is host memory, every CTA runs with its threads as fibers, __syncthreads() switches fibers.
What this checks on a CPU-only machine is the LOGIC the GPU tests check -- tile addressing, Stockham stages, fused
twiddles, the fused real-FFT split, pass chains and their ticket order, plan construction, the host-pointer staging
-- against the same oracle and tolerances.  It is test infrastructure: genfft_b200 cannot load the emulator (below),
the emulated run says nothing about performance or the inline-PTX paths, and the parity claim of the product rests on
the -m gpu run on a real B200.
"""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu import backend  # noqa: E402

GPU_TEST_FILES = ["test_gpu_c2c.py", "test_gpu_real_vert_2d.py", "test_gpu_chain.py", "test_gpu_random_sweep.py",
                  "test_gpu_dist_kernels.py", "test_gpu_zz_ranks_in_process.py"]


def test_emulator_builds_and_exports_the_abi():
    backend.build()
    h = ctypes.CDLL(backend.LIB)
    import genfft_b200
    assert not [s for s in genfft_b200.exported_symbols() if not hasattr(h, s)]
    assert hasattr(h, "genfft_emu_fiber_switches")


def test_product_never_loads_the_emulator(monkeypatch):
    """genfft_b200 has no CPU path: pointing its loader at the emulator is refused."""
    backend.build()
    from genfft_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", backend.LIB)
    with pytest.raises(_lib.GenfftCudaError, match="emulator"):
        _lib.lib()
    # and the product library itself carries no emulator code
    real = os.path.join(ROOT, "genfft_b200", "lib", "libgenfft_cuda.so")
    if os.path.exists(real):
        assert not hasattr(ctypes.CDLL(real), "genfft_emu_fiber_switches")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "genfft_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "tests.emu" not in src and "libgenfft_emu" not in src and "from emu" not in src, f


def test_gpu_suite_on_the_emulator():
    backend.build()
    env = dict(os.environ, GENFFT_TEST_BACKEND="emu")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-n", str(min(8, os.cpu_count() or 1)), "-p", "no:cacheprovider"]
    cmd += [os.path.join(ROOT, "tests", f) for f in GPU_TEST_FILES]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-4000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 400, tail


@pytest.mark.parametrize("binary", ["tests/cpp/_build/test_mirror", "oracle/_ref/ref_plugin_test"])
def test_cpp_layers_on_the_emulator(binary):
    """The C++ class mirror (include/genfft_cuda/fft.h) and the reference's own classes with the CUDA factories
    (include/genfft_cuda/backend.h), i.e. the programs of tests/test_gpu_cpp.py, with the emulator's C ABI symbols
    interposed in front of libgenfft_cuda.so's."""
    path = os.path.join(ROOT, binary)
    if not os.path.exists(path):
        pytest.skip(f"{binary} was not prebuilt")
    backend.build()
    r = subprocess.run([path], capture_output=True, text=True, timeout=600, env=dict(os.environ, LD_PRELOAD=backend.LIB))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout


def test_abi_fuzz_on_the_emulator():
    """A short seeded run of tools/emu_fuzz.py: random sizes / batches / distances / strides / in-place calls through
    every public class on host and "device" pointers, against numpy, with sentinels around the outputs and guard pages
    behind every plan-owned allocation."""
    backend.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz.py"), "11", "250"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "cases ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def _dist_worker(rank, world, port, q):
    """One rank of a gloo group running the PRODUCT's CudaSlabEngine (its real kernels, on the emulator) under the
    product's DistFFT2D / DistFFT1D orchestration with the packed ("nccl") transport: real kernels, real collectives,
    two address spaces."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import numpy as np
    import torch
    import torch.distributed as dist
    from emu import backend as be
    be.install()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genfft_b200.dist import DistFFT1D, DistFFT2D, four_step_shape
        worst = 0.0
        for dt, tol in ((np.float32, 2e-5), (np.float64, 2e-13)):
            cd = np.complex64 if dt == np.float32 else np.complex128
            for w, h in ((64, 32), (512, 256)):
                rng = np.random.default_rng(w + h)
                full = (rng.uniform(-1, 1, (h, w)) + 1j * rng.uniform(-1, 1, (h, w))).astype(cd)
                hl, wp = h // world, w // world
                for transposed in (False, True):
                    for inv in (False, True):
                        plan = DistFFT2D(w, h, dt, transport="nccl", transposed_out=transposed)
                        got = plan.transform(torch.from_numpy(full[rank * hl:(rank + 1) * hl].copy()), inv).numpy()
                        want = np.fft.ifft2(full.astype(np.complex128)) * (w * h) if inv else np.fft.fft2(full.astype(np.complex128))
                        want = want[:, rank * wp:(rank + 1) * wp] if transposed else want[rank * hl:(rank + 1) * hl]
                        worst = max(worst, float(np.linalg.norm(got - want) / np.linalg.norm(want)) / tol)
                        plan.close()
            for n in (1 << 10, 1 << 15):
                rng = np.random.default_rng(n)
                full = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
                hh, ww = four_step_shape(n, world)
                for transposed in (False, True):
                    plan = DistFFT1D(n, dt, transport="nccl", transposed_out=transposed)
                    got = plan.transform(torch.from_numpy(full[rank * n // world:(rank + 1) * n // world].copy())).numpy()
                    want = np.fft.fft(full.astype(np.complex128))
                    want = want.reshape(ww, hh).T[rank * hh // world:(rank + 1) * hh // world] if transposed \
                        else want[rank * n // world:(rank + 1) * n // world]
                    worst = max(worst, float(np.linalg.norm(got.reshape(want.shape) - want) / np.linalg.norm(want)) / tol)
                    plan.close()
        q.put((rank, worst))
    finally:
        dist.destroy_process_group()


def test_distributed_transforms_real_kernels_over_gloo():
    import socket

    import torch.multiprocessing as mp
    backend.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert len(res) == world and max(res.values()) < 1.0, res  # rel-L2 in units of the tolerance


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_thread_order_independence_on_the_emulator(order):
    """The emulator runs the threads of a CTA one after another between two barriers; a read that lacks a
    __syncthreads() after another thread's write sees stale data when the reader runs first.  Forward order (the other
    tests) exposes that for writers with the higher index, reverse order for the lower, and a different rotation and
    direction per barrier phase mixes both -- a racecheck for the kernels' shared-memory exchanges without a GPU.
    (A slice of the suite here; the whole emulated suite passes in all three orders.)"""
    backend.build()
    env = dict(os.environ, GENFFT_TEST_BACKEND="emu", GENFFT_EMU_ORDER=order)
    sweep = os.path.join(ROOT, "tests", "test_gpu_random_sweep.py")
    nodes = [sweep + "::test_c2c_random[0]", sweep + "::test_r2c_c2r_random[1]", sweep + "::test_vert_and_2d_random[1]",
             os.path.join(ROOT, "tests", "test_gpu_chain.py") + "::test_vert_chain_is_bit_identical"]
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-n", "4", "-p", "no:cacheprovider"] + nodes
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 6, r.stdout[-2000:]
