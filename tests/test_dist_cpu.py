"""CPU tests (-m "not gpu") of the multi-GPU host logic with the gloo backend, world_size 2 and 4.

The slab orchestration (genfft_b200/dist.py: partitioning, per-destination packing order, the two
all_to_all_single calls, the unpack) is exercised with a CPU engine injected in place of the CUDA one -- a
test double built on numpy's FFT; the product engine is CUDA-only.  Results are compared with the full 2D
transform, for natural-order and transposed output, forward and inverse.
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_batch_partitions_exactly():
    from genfft_b200.dist import shard_batch
    for batch in (1, 7, 256, 65536):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class CpuEngine:
    """Test double with the interface of genfft_b200.dist.CudaSlabEngine (nccl transport part)."""

    def __init__(self, width, height, world):
        self.w, self.h, self.p = width, height, world
        self.hl, self.wp = height // world, width // world
        self.cdtype = torch.complex128

    def empty(self, *shape):
        return torch.zeros(shape, dtype=self.cdtype)

    def rows_pack(self, slab, send, inv):
        x = slab.numpy()
        y = np.fft.ifft(x, axis=1) * self.w if inv else np.fft.fft(x, axis=1)
        send.copy_(torch.from_numpy(np.ascontiguousarray(y.reshape(self.hl, self.p, self.wp).transpose(1, 0, 2))))

    def cols(self, out, block, inv):
        x = block.numpy()
        out.copy_(torch.from_numpy(np.fft.ifft(x, axis=0) * self.h if inv else np.fft.fft(x, axis=0)))

    def unpack(self, out, recv):
        out.copy_(recv.permute(1, 0, 2).reshape(self.hl, self.w))

    # the extra local steps of the distributed four-step 1D transform
    def pack_cols(self, send, slab):
        send.copy_(slab.reshape(self.hl, self.p, self.wp).permute(1, 0, 2))

    def twiddle(self, slab, row0, inv):
        kr = np.arange(row0, row0 + self.hl)[:, None]
        c = np.arange(self.w)[None, :]
        tw = np.exp((2j if inv else -2j) * np.pi * (kr * c) / (self.w * self.h))
        slab.copy_(torch.from_numpy(slab.numpy() * tw))

    def rows(self, out, slab, inv):
        x = slab.numpy()
        out.copy_(torch.from_numpy(np.fft.ifft(x, axis=1) * self.w if inv else np.fft.fft(x, axis=1)))

    def transpose(self, out, block):
        out.copy_(block.t())


def _worker(rank, world, port, w, h, transposed, inv, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genfft_b200.dist import DistFFT2D
        rng = np.random.default_rng(5)
        full = rng.uniform(-1, 1, (h, w)) + 1j * rng.uniform(-1, 1, (h, w))
        hl, wp = h // world, w // world
        plan = DistFFT2D(w, h, np.float64, transport="nccl", transposed_out=transposed, engine=CpuEngine(w, h, world))
        got = plan.transform(torch.from_numpy(full[rank * hl:(rank + 1) * hl].copy()), inv).numpy()
        want = np.fft.ifft2(full) * (w * h) if inv else np.fft.fft2(full)
        want = want[:, rank * wp:(rank + 1) * wp] if transposed else want[rank * hl:(rank + 1) * hl]
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("transposed,inv", [(False, False), (True, False), (False, True)])
def test_slab_2d_orchestration_gloo(world, transposed, inv):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 32, 16, transposed, inv, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    errs = dict(q.get(timeout=10) for _ in range(world))
    assert len(errs) == world and max(errs.values()) < 1e-12, errs


def _worker_1d(rank, world, port, n, transposed, inv, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genfft_b200.dist import DistFFT1D, four_step_shape
        rng = np.random.default_rng(11)
        full = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        h, w = four_step_shape(n, world)
        plan = DistFFT1D(n, np.float64, transport="nccl", transposed_out=transposed, engine=CpuEngine(w, h, world))
        shard = torch.from_numpy(full[rank * n // world:(rank + 1) * n // world].copy())
        got = plan.transform(shard, inv).numpy()
        want = np.fft.ifft(full) * n if inv else np.fft.fft(full)
        if transposed:  # Z[kr][kc] = X[kr + H*kc], this rank's rows kr
            want = want.reshape(w, h).T[rank * h // world:(rank + 1) * h // world]
        else:
            want = want[rank * n // world:(rank + 1) * n // world]
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 64), (4, 512)])
@pytest.mark.parametrize("transposed,inv", [(False, False), (True, False), (False, True)])
def test_four_step_1d_orchestration_gloo(world, n, transposed, inv):
    """DistFFT1D: the three global transposes, the twiddle's row offset and the output order, against numpy's fft of
    the whole sequence (local steps by the numpy test double)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_1d, args=(r, world, port, n, transposed, inv, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    errs = dict(q.get(timeout=10) for _ in range(world))
    assert len(errs) == world and max(errs.values()) < 1e-12, errs


def test_four_step_shape():
    from genfft_b200.dist import four_step_shape
    assert four_step_shape(1 << 24, 8) == (4096, 4096)
    assert four_step_shape(1 << 25, 8) == (4096, 8192)
    assert four_step_shape(64, 2) == (8, 8)
    with pytest.raises(ValueError):
        four_step_shape(16, 8)  # H = 4 rows cannot be split over 8 ranks
    with pytest.raises(ValueError):
        four_step_shape(48, 2)
