"""Single-GPU tests of the local building blocks of the distributed four-step 1D transform (twiddle2d, transpose)
and of DistFFT1D itself on a one-rank process group (every global transpose degenerates to a local copy, so the
whole p2p path -- peer buffers, flag barrier, column pass, twiddle, row pass, final transpose -- runs on one GPU)."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402
from genfft_b200._lib import check  # noqa: E402

TCPX = {np.float32: torch.complex64, np.float64: torch.complex128}


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("rows,cols,row0,lgn", [(4, 8, 0, 5), (16, 64, 48, 12), (33, 100, 7, 13), (8, 4096, 4088, 24),
                                                 (2, 1 << 15, (1 << 15) - 2, 30)])
@pytest.mark.parametrize("inv", [False, True])
def test_twiddle2d(dt, rows, cols, row0, lgn, inv):
    n = 1 << lgn
    rng = np.random.default_rng(rows + cols)
    x = (rng.uniform(-1, 1, (rows, cols + 3)) + 1j * rng.uniform(-1, 1, (rows, cols + 3))).astype(np.complex128)
    d = torch.from_numpy(x).to(TCPX[dt]).cuda()
    check(g.lib().genfft_cuda_twiddle2d_dev(g.F32 if dt == np.float32 else g.F64, d.data_ptr(), cols + 3, rows, cols,
                                            row0, n, int(inv), None))
    e = (np.arange(row0, row0 + rows, dtype=np.int64)[:, None] * np.arange(cols, dtype=np.int64)[None, :]) % n
    ang = 2 * np.pi * (e.astype(np.float64) / n)  # e/n is exact enough: e < 2^30 here
    tw = np.cos(ang) + (1j if inv else -1j) * np.sin(ang)
    want = x.astype(np.complex64 if dt == np.float32 else np.complex128).astype(np.complex128)
    want[:, :cols] *= tw
    got = d.cpu().numpy().astype(np.complex128)
    assert np.max(np.abs(got - want)) <= (6e-7 if dt == np.float32 else 4e-15)
    assert np.array_equal(got[:, cols:], want[:, cols:])  # padding untouched


def test_twiddle2d_rejects_exponents_beyond_n():
    d = torch.zeros((4, 8), dtype=torch.complex64, device="cuda")
    assert g.lib().genfft_cuda_twiddle2d_dev(g.F32, d.data_ptr(), 8, 4, 8, 1, 32, 0, None) != 0  # (1 + 4) * 8 > 32
    assert g.lib().genfft_cuda_twiddle2d_dev(g.F32, d.data_ptr(), 8, 4, 8, 0, 48, 0, None) != 0  # not a power of two


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("rows,cols", [(1, 1), (32, 32), (5, 77), (300, 129), (4096, 512)])
def test_transpose(dt, rows, cols):
    x = torch.randn(rows, cols + 2, dtype=TCPX[dt], device="cuda")
    out = torch.full((cols, rows + 1), 7.0, dtype=TCPX[dt], device="cuda")
    check(g.lib().genfft_cuda_transpose_dev(g.F32 if dt == np.float32 else g.F64, out.data_ptr(), rows + 1,
                                            x.data_ptr(), cols + 2, rows, cols, None))
    assert torch.equal(out[:, :rows], x[:, :cols].t())
    assert bool((out[:, rows:] == 7.0).all())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def one_rank_group():
    import torch.distributed as dist
    if dist.is_initialized():
        yield
        return
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("lg", [3, 6, 11, 16, 21, 24])
@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_dist_fft1d_on_one_rank(one_rank_group, comparand, dt, lg, transport):
    """The four-step path end to end against genFFT's CPU output of the whole sequence."""
    import oracle
    from genfft_b200.dist import DistFFT1D, four_step_shape
    n = 1 << lg
    if dt == np.float64 and lg > 21:
        pytest.skip("kept small in double to bound CPU time")
    rng = np.random.default_rng(lg)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64 if dt == np.float32 else np.complex128)
    want = comparand.c2c(x)
    h, w = four_step_shape(n, 1)
    for transposed in (False, True):
        plan = DistFFT1D(n, dt, transport=transport, transposed_out=transposed)
        d = torch.from_numpy(x).cuda()
        got = plan.transform(d).cpu().numpy()
        ref = want.reshape(w, h).T if transposed else want
        assert oracle.rel_l2(got, ref) <= oracle.tolerance(n, dt), (lg, transposed)
        if not transposed:
            back = plan.transform(torch.from_numpy(want).cuda(), True).cpu().numpy()
            assert oracle.rel_l2(back, x.astype(np.complex128) * n) <= oracle.tolerance(n, dt)
        plan.close()
