"""The distributed transforms' LOCAL kernels with several ranks played by one process on one device.

CudaSlabEngine's peer-storing passes (dist_rows / dist_cols with out_peers: the fused all-to-all of DistFFT2D's p2p
transport, genfft_b200/dist.py) only need the peers' buffer addresses.  Here all P ranks' buffers live on the same
device and the ranks run one after another, in the order DistFFT2D.transform / DistFFT1D.transform issue the phases,
so the scatter addressing for P = 2, 4, 8 is checked on a single GPU (and, through tests/test_emu_suite.py, on the
kernel-logic emulator without any GPU).  The multi-process path proper is tests/test_gpu_dist.py.
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from genfft_b200.dist import CudaSlabEngine, four_step_shape  # noqa: E402

TCPX = {np.float32: torch.complex64, np.float64: torch.complex128}
NCPX = {np.float32: np.complex64, np.float64: np.complex128}


def rand_c(rng, shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(NCPX[dt])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("w,h", [(8, 8), (64, 16), (256, 512), (1024, 64), (32768, 8)])
@pytest.mark.parametrize("inv", [False, True])
def test_fft2d_slabs_p2p(comparand, dt, world, w, h, inv):
    if w % world or h % world:
        pytest.skip("shape not divisible by the number of ranks")
    rng = np.random.default_rng(w * 31 + h)
    x = rand_c(rng, (h, w), dt)
    hl, wp = h // world, w // world
    eng = [CudaSlabEngine(w, h, world, dt) for _ in range(world)]
    slabs = [torch.from_numpy(x[r * hl:(r + 1) * hl]).cuda() for r in range(world)]
    blocks = [torch.full((h, wp), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
    outs = [torch.full((hl, w), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
    for r in range(world):  # rows + transpose 1: every rank stores into every peer's (H x W/P) block
        eng[r].rows_to_peers(slabs[r], [b.data_ptr() for b in blocks], r, inv)
    torch.cuda.synchronize()
    for r in range(world):  # columns + transpose 2: back to row slabs, natural order
        eng[r].cols_to_peers(blocks[r].data_ptr(), [o.data_ptr() for o in outs], r, inv)
    torch.cuda.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    want = comparand.fft2d(x, inv)  # genFFT's own FFT2D::transform<inv> (fft.h:213-241)
    assert oracle.rel_l2(got, want) <= oracle.tolerance(w * h, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("w,h", [(64, 16), (512, 256)])
def test_fft2d_slabs_packed_transport(comparand, dt, world, w, h):
    """The NCCL transport's local kernels: rows_pack -> (all-to-all played by a host permutation) -> cols -> unpack."""
    rng = np.random.default_rng(w + h)
    x = rand_c(rng, (h, w), dt)
    hl, wp = h // world, w // world
    eng = [CudaSlabEngine(w, h, world, dt) for _ in range(world)]
    sends = []
    for r in range(world):
        send = eng[r].empty(world, hl, wp)
        eng[r].rows_pack(torch.from_numpy(x[r * hl:(r + 1) * hl]).cuda(), send, False)
        sends.append(send)
    torch.cuda.synchronize()
    outs2 = []
    for g in range(world):  # all_to_all_single: rank g receives block g of every rank
        block = torch.cat([sends[r][g] for r in range(world)], dim=0).contiguous()
        bo = eng[g].empty(h, wp)
        eng[g].cols(bo, block, False)
        outs2.append(bo)
    torch.cuda.synchronize()
    got_t = np.concatenate([o.cpu().numpy() for o in outs2], axis=1)  # transposed-output mode: (H x W) by column blocks
    want = comparand.fft2d(x)
    assert oracle.rel_l2(got_t, want) <= oracle.tolerance(w * h, dt)
    for r in range(world):  # second all-to-all + unpack into the natural-order row slab
        recv = torch.stack([outs2[g][r * hl:(r + 1) * hl] for g in range(world)], dim=0).contiguous()
        out = eng[r].empty(hl, w)
        eng[r].unpack(out, recv)
        torch.cuda.synchronize()
        assert oracle.rel_l2(out.cpu().numpy(), want[r * hl:(r + 1) * hl]) <= oracle.tolerance(w * h, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("lg", [6, 11, 16, 24])
@pytest.mark.parametrize("inv", [False, True])
def test_four_step_1d_p2p(comparand, dt, world, lg, inv):
    """DistFFT1D's p2p phases (strided peer copy, column pass + scatter, twiddle, row pass + scatter, transpose)."""
    n = 1 << lg
    if lg >= 24 and os.environ.get("GENFFT_TEST_BACKEND") == "emu" and (dt == np.float64 or inv or world != 4):
        pytest.skip("on the emulator the large case runs once (float, forward, four ranks)")
    h, w = four_step_shape(n, world)
    hl, wp = h // world, w // world
    rng = np.random.default_rng(lg)
    x = rand_c(rng, n, dt)
    eng = [CudaSlabEngine(w, h, world, dt) for _ in range(world)]
    slabs = [torch.from_numpy(x[r * hl * w:(r + 1) * hl * w].reshape(hl, w)).cuda() for r in range(world)]
    blocks = [torch.full((h, wp), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
    mids = [torch.full((hl, w), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
    blocks2 = [torch.full((h, wp), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
    for r in range(world):  # transpose 1: row slabs -> column blocks
        eng[r].cols_blocks_to_peers(slabs[r], [b.data_ptr() for b in blocks], r)
    torch.cuda.synchronize()
    fused = []
    for r in range(world):  # length-H column transforms, scattered back to row slabs (rows = kr), W_n^(kr c) fused
        fused.append(eng[r].cols_to_peers(blocks[r].data_ptr(), [m.data_ptr() for m in mids], r, inv, twiddle_n=n))
    torch.cuda.synchronize()
    assert len(set(fused)) == 1 and (fused[0] or h <= 2048 or world == 1)  # multi-pass columns carry the twiddle
    for r in range(world):  # W_n^(kr c) where the stores did not carry it, then length-W row transforms -> Z[kr][kc]
        if not fused[r]:
            eng[r].twiddle(mids[r], r * hl, inv)
        eng[r].rows_to_peers(mids[r], [b.data_ptr() for b in blocks2], r, inv)
    torch.cuda.synchronize()
    want = comparand.c2c(x, inv)  # genFFT's own FFT::transform<inv> (fft.h:80-85)
    z = np.concatenate([b.cpu().numpy() for b in blocks2], axis=1)  # Z[kr][kc] = X[kr + H kc]
    assert oracle.rel_l2(z, want.reshape(w, h).T) <= oracle.tolerance(n, dt)
    outs = []
    for r in range(world):  # natural order: rank r holds X[r n/P : (r+1) n/P] = rows kc of Z^T
        out = eng[r].empty(wp, h)
        eng[r].transpose(out, blocks2[r])
        outs.append(out)
    torch.cuda.synchronize()
    got = np.concatenate([o.cpu().numpy().reshape(-1) for o in outs])
    assert oracle.rel_l2(got, want) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("inv", [False, True])
def test_peer_modes_run_and_match_the_generic_store(monkeypatch, world, inv):
    """The last pass of a multi-pass distributed transform runs in the compile-time peer modes (M_PEER2/4/8,
    tile_kernel.cuh) -- not in the run-time generic mode -- and stores bit-identical results."""
    from genfft_b200 import lib
    w, h, dt = 32768, 8, np.float32
    mode = {2: 8, 4: 9, 8: 10}[world]
    x = rand_c(np.random.default_rng(world), (h, w), dt)
    hl, wp = h // world, w // world

    def run():
        eng = [CudaSlabEngine(w, h, world, dt) for _ in range(world)]
        blocks = [torch.full((h, wp), float("nan"), dtype=TCPX[dt], device="cuda") for _ in range(world)]
        for r in range(world):
            eng[r].rows_to_peers(torch.from_numpy(x[r * hl:(r + 1) * hl]).cuda(), [b.data_ptr() for b in blocks], r, inv)
        torch.cuda.synchronize()
        return torch.cat(blocks, dim=1)

    n0 = lib().genfft_cuda_debug_mode_launch_count(mode)
    fast = run()
    assert lib().genfft_cuda_debug_mode_launch_count(mode) - n0 == world
    monkeypatch.setenv("GENFFT_CUDA_PEER_MODES", "0")
    g0 = lib().genfft_cuda_debug_mode_launch_count(0)
    generic = run()
    assert lib().genfft_cuda_debug_mode_launch_count(0) - g0 == world
    assert torch.equal(fast, generic)
