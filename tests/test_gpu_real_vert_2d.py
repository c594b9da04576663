"""Parity of the real-input, DIT, vertical and 2D CUDA paths against genFFT's CPU output.

Structure follows the reference's TestRealFFT_Pow2 / TestDIT_Pow2 / TestFFTVert_Pow2
(test/fft_test_impl.h:60-131) and their size lists (test/test_real_fft.cpp:42-67, test/test_dit.cpp:45-67,
test/test_dispatch.cpp:84-126).  FFT2D has no test in the reference; it is checked against the compiled
reference itself and numpy's fft2.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402

CPX = {np.float32: np.complex64, np.float64: np.complex128}
TCPX = {np.float32: torch.complex64, np.float64: torch.complex128}
FILL = 43 + 21j  # the reference's corruption sentinel (test/fft_test_impl.h:88)


def rand_cpx(rng, shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(CPX[dt])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("half", [True, False])
@pytest.mark.parametrize("lg", [0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 13, 14, 16, 18, 20, 22])
def test_real_fft_vs_reference(comparand, checkers, dt, half, lg):
    n = 1 << lg
    ref = checkers[0]
    x = ref.dummy_real(n, dt) if ref is not None else np.random.default_rng(lg).uniform(-1, 1, n).astype(dt)
    want = comparand.r2c(x, half, fill=FILL)
    plan = g.RealFFT(n, dt, half=half)
    d_out = torch.full((max(n, 1),), FILL, dtype=TCPX[dt], device="cuda")
    plan.forward(d_out, torch.from_numpy(x).cuda(), half)
    got = d_out.cpu().numpy()
    limit = 1 if n == 1 else (n // 2 + 1 if half else n)
    assert oracle.rel_l2(got[:limit], want[:limit]) <= oracle.tolerance(n, dt), plan.describe()
    eps = (1e-5 + n * 1e-8) if dt == np.float32 else (1e-8 + n * 1e-12)
    assert np.max(np.abs(got[:limit] - want[:limit])) <= eps
    # "Corruption detected": nothing may be written past n/2+1 when half (test/fft_test_impl.h:102-105)
    assert np.all(got[limit:] == FILL)
    # host-pointer path
    h_out = np.full(max(n, 1), FILL, dtype=CPX[dt])
    plan.forward(h_out, x)
    assert oracle.rel_l2(h_out[:limit], want[:limit]) <= oracle.tolerance(n, dt)
    assert np.all(h_out[limit:] == FILL)


@pytest.mark.parametrize("n,batch", [(8, 5), (1024, 7), (1 << 14, 3), (1 << 16, 2)])
def test_real_fft_batched(comparand, n, batch):
    x = np.random.default_rng(n).uniform(-1, 1, (batch, n)).astype(np.float32)
    plan = g.RealFFT(n, np.float32, half=True, batch=batch)
    d_out = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
    plan.forward(d_out, torch.from_numpy(x).cuda())
    got = d_out.cpu().numpy()
    for b in range(batch):
        want = comparand.r2c(x[b], True)[: n // 2 + 1]
        assert oracle.rel_l2(got[b], want) <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("n,batch,half", [(64, 3, True), (4096, 70000, True), (1 << 15, 5, False), (1 << 17, 3, True),
                                          (1 << 17, 2, False), (1 << 20, 2, True)])
def test_real_fft_unfused_split_path(comparand, monkeypatch, n, batch, half):
    """The stand-alone split kernel (GENFFT_CUDA_FUSED_DIT=0) stays correct, incl. batch > 65535 grid rows."""
    monkeypatch.setenv("GENFFT_CUDA_FUSED_DIT", "0")
    x = np.random.default_rng(n).uniform(-1, 1, (batch, n)).astype(np.float32)
    lim = n // 2 + 1 if half else n
    plan = g.RealFFT(n, np.float32, half=half, batch=batch)
    d_out = torch.empty((batch, lim), dtype=torch.complex64, device="cuda")
    plan.forward(d_out, torch.from_numpy(x).cuda())
    got = d_out.cpu().numpy()
    for b in sorted({0, batch // 2, batch - 1}):
        assert oracle.rel_l2(got[b], comparand.r2c(x[b], half)[:lim]) <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("lg", [1, 2, 3, 4, 6, 10, 14, 18, 20])
def test_dit_vs_reference(comparand, dt, in_place, lg):
    """DIT::apply of the n/2-point FFT equals the n-point FFT of the real signal (TestDIT_Pow2)."""
    n = 1 << lg
    x = np.random.default_rng(lg).uniform(-1, 1, n).astype(dt)
    z = comparand.c2c(x.view(CPX[dt])) if n >= 4 else x.view(CPX[dt]).copy()
    want = comparand.dit(z, n, half=False, in_place=in_place)
    plan = g.DIT(n, dt)
    d_in = torch.zeros(n, dtype=TCPX[dt], device="cuda")
    d_in[: n // 2] = torch.from_numpy(z).cuda()
    d_out = d_in if in_place else torch.zeros_like(d_in)
    plan.apply(d_out, d_in, False)
    got = d_out.cpu().numpy()
    assert oracle.rel_l2(got, want) <= oracle.tolerance(n, dt)
    full = comparand.c2c(x.astype(CPX[dt]))
    assert oracle.rel_l2(got, full) <= oracle.tolerance(n, dt)


# the reference's own (nfft, cols) list: test/test_dispatch.cpp:96-120
VERT_CASES = [(1, 1), (2, 1), (4, 1), (2, 3), (4, 7), (8, 33), (16, 47), (32, 63), (64, 5), (128, 767), (256, 999),
              (512, 1023), (1024, 31), (4096, 17), (8192, 3), (16384, 2), (65536, 31)]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,cols", VERT_CASES)
def test_vert_vs_reference(comparand, dt, n, cols):
    if dt == np.float64 and n * cols > (1 << 21):
        pytest.skip("kept small in double to bound CPU time")
    x = rand_cpx(np.random.default_rng(n + cols), (n, cols), dt)
    plan = g.FFTVert(n, dt)
    for inv in (False, True):
        want = comparand.vert(x, inv)
        d_out = torch.empty((n, cols), dtype=TCPX[dt], device="cuda")
        plan.transform(d_out, torch.from_numpy(x).cuda(), cols, inv=inv)
        assert oracle.rel_l2(d_out.cpu().numpy(), want) <= oracle.tolerance(n, dt), plan.describe()


def test_vert_strided_and_host(comparand):
    n, cols, in_stride, out_stride = 256, 37, 40, 64
    x = rand_cpx(np.random.default_rng(2), (n, in_stride), np.float32)
    want = comparand.vert(x, False, cols=cols)[:, :cols]
    plan = g.FFTVert(n, np.float32)
    d_out = torch.full((n, out_stride), FILL, dtype=torch.complex64, device="cuda")
    plan.transform(d_out, torch.from_numpy(x).cuda(), cols, out_stride=out_stride, in_stride=in_stride)
    got = d_out.cpu().numpy()
    assert oracle.rel_l2(got[:, :cols], want) <= oracle.tolerance(n, np.float32)
    assert np.all(got[:, cols:] == FILL)
    h_out = np.full((n, out_stride), FILL, dtype=np.complex64)
    plan.transform(h_out, x, cols, out_stride=out_stride, in_stride=in_stride)
    assert oracle.rel_l2(h_out[:, :cols], want) <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("n,cols", [(8, 5), (256, 33), (8192, 4)])
def test_vert_no_scramble(comparand, n, cols):
    """FFTVert::transform_no_scramble (fft.h:132-136): rows arrive bit-reversed, in place."""
    x = rand_cpx(np.random.default_rng(n), (n, cols), np.float32)
    want = comparand.vert(x)
    bits = n.bit_length() - 1
    perm = np.array([int(format(i, f"0{bits}b")[::-1], 2) if bits else 0 for i in range(n)])
    scr = np.empty_like(x)
    scr[perm] = x  # scramble_rows: out row bitrev(r) = in row r
    d = torch.from_numpy(scr).cuda()
    g.FFTVert(n, np.float32).transform_no_scramble(d, cols, cols)
    assert oracle.rel_l2(d.cpu().numpy(), want) <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (4, 8), (8, 4), (64, 32), (16, 1024), (1024, 16), (512, 512),
                                 (2048, 256), (256, 8192), (32768, 8)])
def test_fft2d_vs_reference(comparand, dt, w, h):
    x = rand_cpx(np.random.default_rng(w * 3 + h), (h, w), dt)
    plan = g.FFT2D(w, h, dt)
    assert plan.cols() == w and plan.rows() == h
    for inv in (False, True):
        want = comparand.fft2d(x, inv)
        d_out = torch.empty((h, w), dtype=TCPX[dt], device="cuda")
        plan.transform(d_out, torch.from_numpy(x).cuda(), inv=inv)
        got = d_out.cpu().numpy()
        assert oracle.rel_l2(got, want) <= oracle.tolerance(w * h, dt), plan.describe()
    np_want = np.fft.fft2(x.astype(np.complex128))
    d_out = torch.empty((h, w), dtype=TCPX[dt], device="cuda")
    plan.transform(d_out, torch.from_numpy(x).cuda())
    assert oracle.rel_l2(d_out.cpu().numpy(), np_want) <= oracle.tolerance(w * h, dt)


def test_fft2d_strides_and_host(comparand):
    w, h, in_stride, out_stride = 64, 16, 70, 96
    buf = rand_cpx(np.random.default_rng(3), (h, in_stride), np.float32)
    want = comparand.fft2d(np.ascontiguousarray(buf[:, :w]))
    plan = g.FFT2D(w, h, np.float32)
    d_out = torch.full((h, out_stride), FILL, dtype=torch.complex64, device="cuda")
    plan.transform(d_out, torch.from_numpy(buf).cuda(), out_stride=out_stride, in_stride=in_stride)
    got = d_out.cpu().numpy()
    assert oracle.rel_l2(got[:, :w], want) <= oracle.tolerance(w * h, np.float32)
    assert np.all(got[:, w:] == FILL)
    h_out = np.full((h, out_stride), FILL, dtype=np.complex64)
    plan.transform(h_out, buf, out_stride=out_stride, in_stride=in_stride)
    assert oracle.rel_l2(h_out[:, :w], want) <= oracle.tolerance(w * h, np.float32)
    with pytest.raises(g.GenfftCudaError):
        plan.transform(h_out, h_out)  # out != in (fft.h:209)


def test_c4_real_2pow22_batch_properties(comparand):
    """BASELINE config C4 (R2C fp32 N=2^22, batch 256) at full size: sampled transforms vs the reference,
    DC / Nyquist bins are real, and Parseval over the whole batch."""
    n, batch = 1 << 22, 256
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((batch, n), generator=gen, device="cuda") * 2 - 1
    out = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
    plan = g.RealFFT(n, np.float32, half=True, batch=batch)
    plan.forward(out, x)
    for b in (0, 100, batch - 1):
        want = comparand.r2c(x[b].cpu().numpy(), True)[: n // 2 + 1]
        assert oracle.rel_l2(out[b].cpu().numpy(), want) <= oracle.tolerance(n, np.float32), plan.describe()
    assert out[:, 0].imag.abs().max().item() == 0 and out[:, -1].imag.abs().max().item() == 0
    ex = (x.double() ** 2).sum(1)
    p = out.abs().double() ** 2
    ey = (2 * p.sum(1) - p[:, 0] - p[:, -1]) / n
    assert ((ey / ex - 1).abs().max().item()) < 1e-5


def test_c5_shape_2d_properties():
    """BASELINE config C5's row/column length (32768) on one GPU at reduced height x full width and
    full height x reduced width: analytic plane-wave response and round trip."""
    for w, h in ((32768, 64), (64, 32768)):
        plan = g.FFT2D(w, h, np.float32)
        kx, ky = 12345 % w, 17 % h
        xs = torch.arange(w, device="cuda", dtype=torch.float64)
        ys = torch.arange(h, device="cuda", dtype=torch.float64)
        phase = 2 * np.pi * (kx * xs[None, :] / w + ky * ys[:, None] / h)
        x = torch.polar(torch.ones_like(phase), phase).to(torch.complex64)
        y = torch.empty_like(x)
        plan.transform(y, x)
        peak = y[ky, kx].item()
        assert abs(peak - w * h) / (w * h) < 1e-4
        y[ky, kx] = 0
        assert y.abs().max().item() / (w * h) < 1e-4
        gen = torch.Generator(device="cuda").manual_seed(3)
        r = torch.view_as_complex(torch.rand((h, w, 2), generator=gen, device="cuda") * 2 - 1)
        f = torch.empty_like(r)
        b = torch.empty_like(r)
        plan.transform(f, r)
        plan.transform(b, f, inv=True)
        assert (b / (w * h) - r).abs().max().item() < 1e-4


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 8, 256, 4096, 1 << 15])
def test_transform_interleave_and_separate(comparand, dt, n):
    """FFT::transform_interleave + separate_2x_real_FFT (fft.h:100-105, FFTReal.h:35-66): two real spectra from
    one complex transform."""
    rng = np.random.default_rng(n)
    a, b = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, n).astype(dt)
    want_a, want_b = comparand.two_real(a, b)
    plan = g.FFT(n, dt)
    z = torch.empty(n, dtype=TCPX[dt], device="cuda")
    plan.transform_interleave(z, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    fa, fb = torch.empty_like(z), torch.empty_like(z)
    g.separate_2x_real_FFT(fa, fb, z, n)
    assert oracle.rel_l2(fa.cpu().numpy(), want_a) <= oracle.tolerance(n, dt)
    assert oracle.rel_l2(fb.cpu().numpy(), want_b) <= oracle.tolerance(n, dt)
    g.separate_2x_real_FFT(z, fb, z, n)  # out1 aliases in, as the reference allows
    assert oracle.rel_l2(z.cpu().numpy(), want_a) <= oracle.tolerance(n, dt)
    h_z = np.empty(n, dtype=CPX[dt])
    plan.transform_interleave(h_z, a, b)  # host pointers
    assert oracle.rel_l2(h_z, np.fft.fft(a.astype(np.float64) + 1j * b.astype(np.float64))) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("w,h", [(2, 2), (4, 2), (8, 8), (64, 16), (16, 256), (1024, 64), (256, 4096), (8192, 8)])
def test_real_fft2d_vs_reference(checkers, dt, w, h):
    """RealFFT2D::forward (FFTReal.h:83-104): untested in the reference; pinned against the compiled reference
    where present and against numpy's fft2 of the real image."""
    x = np.random.default_rng(w + 7 * h).uniform(-1, 1, (h, w)).astype(dt)
    plan = g.RealFFT2D(w, h, dt)
    assert plan.cols() == w and plan.rows() == h
    d_out = torch.full((h, w), FILL, dtype=TCPX[dt], device="cuda")
    plan.forward(d_out, torch.from_numpy(x).cuda())
    got = d_out.cpu().numpy()
    assert oracle.rel_l2(got, np.fft.fft2(x.astype(np.float64))) <= oracle.tolerance(w * h, dt)
    ref = checkers[0]
    if ref is not None and w >= 4 and h >= 2:
        assert oracle.rel_l2(got, ref.real_fft2d(x)) <= oracle.tolerance(w * h, dt)
    h_out = np.empty((h, w + 3), dtype=CPX[dt])
    xin = np.zeros((h, w + 2), dtype=dt)
    xin[:, :w] = x
    plan.forward(h_out, xin, out_stride=w + 3, in_stride=w + 2)  # host pointers, padded strides
    assert oracle.rel_l2(h_out[:, :w], got) <= 1e-6


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("w,h", [(2, 1), (2, 2), (4, 2), (8, 8), (64, 16), (16, 256), (1024, 64), (256, 4096),
                                 (32768, 4), (16384, 16)])
def test_real_fft2d_forward_2x(checkers, monkeypatch, fused, dt, w, h):
    """RealFFT2D::forward_2x (FFTReal.h:106-118): the spectrum of in1 + i*in2.  Untested in the reference; pinned
    against the compiled reference (equal strides -- its row recursion mixes the strides up otherwise, :178) and
    against numpy's fft2.  Both device paths: real/imaginary parts read from the two images by the first pass
    (GENFFT_CUDA_2X_FUSED=1) and the interleaving copy."""
    monkeypatch.setenv("GENFFT_CUDA_2X_FUSED", fused)
    rng = np.random.default_rng(3 * w + h)
    a = rng.uniform(-1, 1, (h, w)).astype(dt)
    b = rng.uniform(-1, 1, (h, w)).astype(dt)
    want = np.fft.fft2(a.astype(np.float64) + 1j * b.astype(np.float64))
    plan = g.RealFFT2D(w, h, dt)
    d_out = torch.full((h, w), FILL, dtype=TCPX[dt], device="cuda")
    plan.forward_2x(d_out, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    got = d_out.cpu().numpy()
    assert oracle.rel_l2(got, want) <= oracle.tolerance(w * h, dt)
    ref = checkers[0]
    if ref is not None:
        assert oracle.rel_l2(got, ref.real_fft2d_2x(a, b)) <= oracle.tolerance(w * h, dt)
    # the two real spectra come apart with the Hermitian split, as after transform_interleave
    fa = np.fft.fft2(a.astype(np.float64))
    gm = np.conj(np.roll(np.roll(got[::-1, ::-1], 1, axis=0), 1, axis=1))  # conj(G[-ky, -kx])
    assert oracle.rel_l2((got + gm) / 2, fa) <= oracle.tolerance(w * h, dt)
    # unequal, padded strides on device pointers (always correct here; the reference's typo makes it unusable)
    pa = torch.zeros((h, w + 2), dtype=torch.from_numpy(a).dtype, device="cuda")
    pb = torch.zeros((h, w + 6), dtype=pa.dtype, device="cuda")
    pa[:, :w] = torch.from_numpy(a).cuda()
    pb[:, :w] = torch.from_numpy(b).cuda()
    d_pad = torch.full((h, w + 1), FILL, dtype=TCPX[dt], device="cuda")
    plan.forward_2x(d_pad, pa, pb, out_stride=w + 1, in_stride1=w + 2, in_stride2=w + 6)
    got_pad = d_pad.cpu().numpy()
    assert oracle.rel_l2(got_pad[:, :w], got) <= 1e-6
    assert np.all(got_pad[:, w:] == FILL)
    # host pointers
    h_out = np.empty((h, w), dtype=CPX[dt])
    plan.forward_2x(h_out, a, b)
    assert oracle.rel_l2(h_out, got) <= 1e-6


@pytest.mark.parametrize("tag,dt", [("f32", np.float32), ("f64", np.float64)])
def test_real_fft2d_vs_committed_reference_outputs(tag, dt):
    """RealFFT2D::forward / forward_2x against tests/golden/genfft_golden_real2d.npz (outputs of the reference itself,
    generated by tests/golden/make_golden.py): the pin that needs neither /root/reference nor its prebuilt library."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "genfft_golden_real2d.npz"))
    for key in [k for k in gold.files if k.startswith(f"real2d_{tag}_") and k.endswith("_in1")]:
        w, h = (int(v) for v in key.split("_")[2].split("x"))
        a, b = gold[key], gold[key.replace("_in1", "_in2")]
        plan = g.RealFFT2D(w, h, dt)
        out = torch.empty((h, w), dtype=TCPX[dt], device="cuda")
        plan.forward(out, torch.from_numpy(a).cuda())
        assert oracle.rel_l2(out.cpu().numpy(), gold[key.replace("_in1", "_forward")]) <= oracle.tolerance(w * h, dt), key
        plan.forward_2x(out, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        assert oracle.rel_l2(out.cpu().numpy(), gold[key.replace("_in1", "_forward_2x")]) <= oracle.tolerance(w * h, dt), key


def test_real_fft2d_forward_2x_errors():
    plan = g.RealFFT2D(8, 4, np.float32)
    a = torch.zeros((4, 8), device="cuda")
    out = torch.zeros((4, 8), dtype=torch.complex64, device="cuda")
    with pytest.raises(g.GenfftCudaError):
        plan.forward_2x(out, a, a, in_stride1=4)  # stride smaller than the width
    with pytest.raises(ValueError):
        plan.forward_2x(out, a, torch.zeros((2, 8), device="cuda"))  # second image too small
    wrong = g.FFT2D(8, 4)  # a c2c_2d plan is not an r2c_2d plan
    assert g.lib().genfft_cuda_exec_r2c_2d_2x_dev(wrong._h, out.data_ptr(), 8, a.data_ptr(), 8, a.data_ptr(), 8, None) != 0
    assert g.lib().genfft_cuda_exec_r2c_2d_2x_dev(plan._h, out.data_ptr(), 8, a.data_ptr(), 8, None, 8, None) != 0


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,batch", [(2, 3), (4, 1), (8, 2), (64, 5), (4096, 3), (1 << 15, 2), (1 << 17, 2), (1 << 20, 1)])
def test_half_spectrum_inverse(dt, n, batch):
    """The inverse real FFT (absent in the reference): inverse(forward(x)) == n * x, and agreement with numpy's
    irfft on an arbitrary Hermitian half spectrum."""
    rng = np.random.default_rng(n + batch)
    x = rng.uniform(-1, 1, (batch, n)).astype(dt)
    fwd = g.RealFFT(n, dt, half=True, batch=batch)
    inv = g.InverseRealFFT(n, dt, batch=batch)
    spec = torch.empty((batch, n // 2 + 1), dtype=TCPX[dt], device="cuda")
    fwd.forward(spec, torch.from_numpy(x).cuda())
    back = torch.empty((batch, n), dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
    inv.inverse(back, spec)
    eps = (1e-5 + n * 1e-8) if dt == np.float32 else (1e-8 + n * 1e-12)
    assert np.max(np.abs(back.cpu().numpy() / n - x)) <= eps
    s = (rng.uniform(-1, 1, (batch, n // 2 + 1)) + 1j * rng.uniform(-1, 1, (batch, n // 2 + 1))).astype(CPX[dt])
    s[:, 0] = s[:, 0].real
    s[:, -1] = s[:, -1].real
    want = np.fft.irfft(s.astype(np.complex128), n=n, axis=1) * n
    inv.inverse(back, torch.from_numpy(s).cuda())
    assert oracle.rel_l2(back.cpu().numpy(), want) <= oracle.tolerance(n, dt)
    h_out = np.empty((batch, n), dtype=dt)
    inv.inverse(h_out, s)  # host pointers
    assert oracle.rel_l2(h_out, want) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_half_spectrum_inverse_host_leaves_gaps_alone(dt):
    """Host-pointer c2r with out_dist > n writes the n real points of every transform and nothing between them."""
    n, batch, out_dist = 256, 5, 262
    rng = np.random.default_rng(2)
    s = (rng.uniform(-1, 1, (batch, n // 2 + 1)) + 1j * rng.uniform(-1, 1, (batch, n // 2 + 1))).astype(CPX[dt])
    s[:, 0] = s[:, 0].real
    s[:, -1] = s[:, -1].real
    out = np.full((batch, out_dist), 42.5, dtype=dt)
    g.InverseRealFFT(n, dt, batch=batch, out_dist=out_dist).inverse(out, s)
    want = np.fft.irfft(s.astype(np.complex128), n=n, axis=1) * n
    assert oracle.rel_l2(out[:, :n], want) <= oracle.tolerance(n, dt)
    assert np.all(out[:, n:] == 42.5)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_separate_on_host_pointers(comparand, dt):
    """separate_2x_real_FFT on host arrays goes through the C ABI (genfft_cuda_separate_2x_real), aliasing allowed."""
    n = 512
    rng = np.random.default_rng(8)
    a, b = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, n).astype(dt)
    z = np.empty(n, CPX[dt])
    g.FFT(n, dt).transform_interleave(z, a, b)
    fa, fb = np.empty_like(z), np.empty_like(z)
    g.separate_2x_real_FFT(fa, fb, z, n)
    ra, rb = comparand.two_real(a, b)
    assert oracle.rel_l2(fa, ra) <= oracle.tolerance(n, dt) and oracle.rel_l2(fb, rb) <= oracle.tolerance(n, dt)
    g.separate_2x_real_FFT(z, fb, z, n)
    assert np.array_equal(z, fa)
