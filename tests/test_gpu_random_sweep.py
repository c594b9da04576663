"""GPU parity sweep over seeded random shapes (sizes, batches, distances, strides, directions) against a
double-precision numpy FFT: catches addressing bugs in combinations the structured tests do not enumerate.
The tolerance is north_star's rel-L2 bound; the reference-output parity proper lives in the other test files."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402

CPX = {np.float32: np.complex64, np.float64: np.complex128}
TC = {np.float32: torch.complex64, np.float64: torch.complex128}
TR = {np.float32: torch.float32, np.float64: torch.float64}


def cases(seed, count, max_lg):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        lg = int(rng.integers(0, max_lg + 1))
        n = 1 << lg
        batch = int(rng.integers(1, max(2, min(200, (1 << 21) // n))))
        dt = np.float32 if rng.random() < 0.6 else np.float64
        yield rng, n, batch, dt


@pytest.mark.parametrize("seed", range(6))
def test_c2c_random(seed):
    for rng, n, batch, dt in cases(seed, 12, 19):
        in_dist = n + int(rng.integers(0, 3)) * 2
        out_dist = n + int(rng.integers(0, 3))
        inv = bool(rng.integers(0, 2))
        buf = (rng.uniform(-1, 1, batch * in_dist) + 1j * rng.uniform(-1, 1, batch * in_dist)).astype(CPX[dt])
        plan = g.FFT(n, dt, batch=batch, in_dist=in_dist, out_dist=out_dist)
        d_out = torch.zeros(batch * out_dist, dtype=TC[dt], device="cuda")
        plan.transform(d_out, torch.from_numpy(buf).cuda(), inv)
        got = d_out.cpu().numpy().reshape(batch, out_dist)[:, :n]
        x = buf.reshape(batch, in_dist)[:, :n].astype(np.complex128)
        want = np.fft.ifft(x, axis=1) * n if inv else np.fft.fft(x, axis=1)
        assert oracle.rel_l2(got, want) <= oracle.tolerance(n, dt), (n, batch, dt.__name__, inv, plan.describe())


@pytest.mark.parametrize("seed", range(4))
def test_r2c_c2r_random(seed):
    for rng, n, batch, dt in cases(100 + seed, 10, 19):
        if n < 2:
            continue
        half = bool(rng.integers(0, 2))
        lim = n // 2 + 1 if half else n
        out_dist = lim + int(rng.integers(0, 4))
        x = rng.uniform(-1, 1, (batch, n)).astype(dt)
        plan = g.RealFFT(n, dt, half=half, batch=batch, out_dist=out_dist)
        d_out = torch.full((batch, out_dist), 9 + 9j, dtype=TC[dt], device="cuda")
        plan.forward(d_out, torch.from_numpy(x).cuda())
        got = d_out.cpu().numpy()
        want = np.fft.fft(x.astype(np.float64), axis=1)[:, :lim]
        assert oracle.rel_l2(got[:, :lim], want) <= oracle.tolerance(n, dt), (n, batch, half, plan.describe())
        assert np.all(got[:, lim:] == 9 + 9j)
        if half:
            inv = g.InverseRealFFT(n, dt, batch=batch, in_dist=out_dist)
            back = torch.empty((batch, n), dtype=TR[dt], device="cuda")
            inv.inverse(back, d_out)
            assert oracle.rel_l2(back.cpu().numpy(), x.astype(np.float64) * n) <= oracle.tolerance(n, dt) * 2


@pytest.mark.parametrize("seed", range(4))
def test_vert_and_2d_random(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(8):
        dt = np.float32 if rng.random() < 0.6 else np.float64
        h = 1 << int(rng.integers(0, 14))
        cols = int(rng.integers(1, max(2, min(300, (1 << 20) // h))))
        stride_in, stride_out = cols + int(rng.integers(0, 5)), cols + int(rng.integers(0, 5))
        inv = bool(rng.integers(0, 2))
        x = (rng.uniform(-1, 1, (h, stride_in)) + 1j * rng.uniform(-1, 1, (h, stride_in))).astype(CPX[dt])
        d_out = torch.full((h, stride_out), 3 - 1j, dtype=TC[dt], device="cuda")
        plan = g.FFTVert(h, dt)
        plan.transform(d_out, torch.from_numpy(x).cuda(), cols, out_stride=stride_out, in_stride=stride_in, inv=inv)
        got = d_out.cpu().numpy()
        x64 = x[:, :cols].astype(np.complex128)
        want = np.fft.ifft(x64, axis=0) * h if inv else np.fft.fft(x64, axis=0)
        assert oracle.rel_l2(got[:, :cols], want) <= oracle.tolerance(h, dt), (h, cols, plan.describe())
        assert np.all(got[:, cols:] == 3 - 1j)
    for _ in range(5):
        dt = np.float32 if rng.random() < 0.6 else np.float64
        lw, lh = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        if lw + lh > 20:
            lh = 20 - lw
        w, h = 1 << lw, 1 << lh
        inv = bool(rng.integers(0, 2))
        x = (rng.uniform(-1, 1, (h, w)) + 1j * rng.uniform(-1, 1, (h, w))).astype(CPX[dt])
        d_out = torch.empty((h, w), dtype=TC[dt], device="cuda")
        plan = g.FFT2D(w, h, dt)
        plan.transform(d_out, torch.from_numpy(x).cuda(), inv=inv)
        x64 = x.astype(np.complex128)
        want = np.fft.ifft2(x64) * (w * h) if inv else np.fft.fft2(x64)
        assert oracle.rel_l2(d_out.cpu().numpy(), want) <= oracle.tolerance(w * h, dt), (w, h, plan.describe())
