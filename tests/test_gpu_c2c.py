"""Parity of the CUDA 1D complex path (through the C ABI) against genFFT's CPU output.

Mirrors the structure of the reference's TestFFT_Pow2 (test/fft_test_impl.h:35-58) and its size loops
(test/test_dispatch.cpp:45-65, test/test_fft.cpp:41-66): forward vs reference, inverse vs reference,
round trip; tolerance is north_star's rel-L2 <= 1e-6*log2N (float) / 1e-14*log2N (double), and the
reference's own absolute FFT_Eps (test/test_util.h:62-72) is checked as well.
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402

CPX = {np.float32: np.complex64, np.float64: np.complex128}


def fft_eps(n, dt):
    return (1e-5 + n * 1e-8) if dt == np.float32 else (1e-8 + n * 1e-12)


def rand_cpx(rng, shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(CPX[dt])


def gpu_c2c(x, inv=False, batch=1):
    n = x.shape[-1]
    plan = g.FFT(n, x.real.dtype, batch=batch)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty_like(d_in)
    plan.transform(d_out, d_in, inv)
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), plan


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("lg", list(range(0, 21)))
def test_c2c_pow2_vs_reference(comparand, checkers, dt, lg):
    n = 1 << lg
    ref = checkers[0]
    x = ref.dummy_complex(n, dt) if ref is not None else rand_cpx(np.random.default_rng(lg), n, dt)
    want_f = comparand.c2c(x, False)
    want_i = comparand.c2c(want_f, True)
    got_f, plan = gpu_c2c(x, False)
    assert plan.size() == n and bool(plan)
    tol = oracle.tolerance(n, dt)
    assert oracle.rel_l2(got_f, want_f) <= tol, plan.describe()
    got_i, _ = gpu_c2c(want_f, True)
    assert oracle.rel_l2(got_i, want_i) <= tol
    # the reference's own acceptance: absolute eps per element, and the round trip inv(fwd(x))/n == x
    eps = fft_eps(n, dt)
    assert np.max(np.abs(got_f - want_f)) <= eps
    back, _ = gpu_c2c(got_f, True)
    assert np.max(np.abs(back / n - x)) <= eps


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_reference_test_harness_acceptance(checkers, dt):
    """The reference's own acceptance run on the CUDA path: TestFFT_Pow2 and TestRealFFT_Pow2
    (test/fft_test_impl.h:35-58,84-106) compare against reference_impl::FFT_pow2 (test/fft_ref_impl.h:90-96) with
    the absolute FFT_Eps (test/test_util.h:62-72) over the size loops of test/test_fft.cpp:41-66 and
    test/test_real_fft.cpp:42-67.  FFT_pow2 accumulates its angles, so in double it is only usable as a comparand
    up to 2^12; float runs the reference's full loop."""
    ref = checkers[0]
    if ref is None:
        pytest.skip("compiled reference (with its test harness) not present")
    for lg in range(1, 17 if dt == np.float32 else 13):
        n = 1 << lg
        eps = fft_eps(n, dt)
        x = ref.dummy_complex(n, dt)
        want = ref.testref_fft_pow2(x)
        got, _ = gpu_c2c(x)
        assert np.max(np.abs(got.real - want.real)) <= eps and np.max(np.abs(got.imag - want.imag)) <= eps, n
        back, _ = gpu_c2c(got, True)
        assert np.max(np.abs(back.real / n - x.real)) <= eps and np.max(np.abs(back.imag / n - x.imag)) <= eps, n
        r = ref.dummy_real(n, dt)
        want_r = ref.testref_fft_pow2(r.astype(x.dtype))
        for half in (True, False):
            lim = n // 2 + 1 if half else n
            out = torch.full((n,), 43 + 21j, dtype=torch.complex64 if dt == np.float32 else torch.complex128,
                             device="cuda")
            g.RealFFT(n, dt, half=half).forward(out, torch.from_numpy(r).cuda(), half)
            got_r = out.cpu().numpy()
            assert np.max(np.abs(got_r[:lim] - want_r[:lim])) <= eps * 1.5, (n, half)
            assert np.all(got_r[lim:] == 43 + 21j), "Corruption detected"  # test/fft_test_impl.h:102-105


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,batch", [(2, 5), (16, 300), (64, 33), (256, 17), (1024, 9), (4096, 64), (4096, 7),
                                     (8192, 3), (1 << 15, 3), (1 << 17, 2)])
def test_c2c_batched(comparand, dt, n, batch):
    rng = np.random.default_rng(n + batch)
    x = rand_cpx(rng, (batch, n), dt)
    for inv in (False, True):
        want = comparand.c2c_batch(x, inv)
        got, plan = gpu_c2c(x, inv, batch=batch)
        assert oracle.rel_l2(got, want) <= oracle.tolerance(n, dt), plan.describe()
        for b in (0, batch - 1):
            assert oracle.rel_l2(got[b], want[b]) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_c2c_batched_with_dist(comparand, dt):
    n, batch, in_dist, out_dist = 512, 6, 520, 600
    rng = np.random.default_rng(1)
    buf = rand_cpx(rng, batch * in_dist, dt)
    plan = g.FFT(n, dt, batch=batch, in_dist=in_dist, out_dist=out_dist)
    d_in = torch.from_numpy(buf).cuda()
    d_out = torch.full((batch * out_dist,), 7 + 5j, dtype=d_in.dtype, device="cuda")
    plan.transform(d_out, d_in)
    got = d_out.cpu().numpy().reshape(batch, out_dist)
    for b in range(batch):
        want = comparand.c2c(buf[b * in_dist:b * in_dist + n].copy())
        assert oracle.rel_l2(got[b, :n], want) <= oracle.tolerance(n, dt)
        assert np.all(got[b, n:] == 7 + 5j)  # gaps are not touched


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 8, 1024, 4096, 1 << 16])
def test_c2c_host_pointers(comparand, dt, n):
    """The literal drop-in: host buffers in, host buffers out (FFT<T>::transform on std::complex<T>*)."""
    rng = np.random.default_rng(n)
    x = rand_cpx(rng, n, dt)
    out = np.empty_like(x)
    plan = g.FFT(n, dt)
    plan.forward(out, x)
    assert oracle.rel_l2(out, comparand.c2c(x)) <= oracle.tolerance(n, dt)
    back = np.empty_like(x)
    plan.inverse(back, out)
    assert np.max(np.abs(back / n - x)) <= fft_eps(n, dt)


def test_c2c_host_pointers_batched_chunked(comparand, monkeypatch):
    monkeypatch.setenv("GENFFT_CUDA_HOST_CHUNK_MB", "1")  # force several pipelined chunks
    n, batch = 4096, 100
    x = rand_cpx(np.random.default_rng(5), (batch, n), np.float32)
    out = np.empty_like(x)
    g.FFT(n, np.float32, batch=batch).forward(out, x)
    assert oracle.rel_l2(out, comparand.c2c_batch(x)) <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 4, 64, 4096, 1 << 15])
def test_transform_no_scramble(comparand, dt, n):
    """FFT<T>::transform_no_scramble (fft.h:69-73): bit-reversed input, in place, natural output."""
    rng = np.random.default_rng(n + 3)
    x = rand_cpx(rng, n, dt)
    want = comparand.c2c_no_scramble(x)
    plan = g.FFT(n, dt)
    d = torch.from_numpy(x.copy()).cuda()
    plan.transform_no_scramble(d)
    assert oracle.rel_l2(d.cpu().numpy(), want) <= oracle.tolerance(n, dt)
    h = x.copy()
    plan.transform_no_scramble(h, inv=True)
    assert oracle.rel_l2(h, comparand.c2c_no_scramble(x, True)) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 16, 1024, 1 << 15])
def test_transform_real(comparand, dt, n):
    """FFT<T>::transform_real (fft.h:90-94)."""
    x = np.random.default_rng(n).uniform(-1, 1, n).astype(dt)
    plan = g.FFT(n, dt)
    d_out = torch.empty(n, dtype=torch.complex64 if dt == np.float32 else torch.complex128, device="cuda")
    plan.transform_real(d_out, torch.from_numpy(x).cuda())
    assert oracle.rel_l2(d_out.cpu().numpy(), comparand.transform_real(x)) <= oracle.tolerance(n, dt)


def test_in_place_device(comparand):
    for n in (4096, 1 << 16):
        x = rand_cpx(np.random.default_rng(n), n, np.float32)
        d = torch.from_numpy(x).cuda()
        g.FFT(n, np.float32).transform(d, d)
        assert oracle.rel_l2(d.cpu().numpy(), comparand.c2c(x)) <= oracle.tolerance(n, np.float32)


def test_errors():
    with pytest.raises(g.GenfftCudaError):
        g.FFT(3)  # not a power of two (reference: assert(!"unsupported size"))
    with pytest.raises(g.GenfftCudaError):
        g.FFT(0)
    empty = g.FFT()
    assert not empty and empty.size() == 0
    with pytest.raises(g.GenfftCudaError):
        empty.transform(np.zeros(4, np.complex64), np.zeros(4, np.complex64))
    plan = g.FFT(8)
    x = np.zeros(8, np.complex64)
    with pytest.raises(g.GenfftCudaError):
        plan.transform(x, x)  # out != in on the host path, like the reference
    with pytest.raises(ValueError):
        plan.transform(np.zeros(4, np.complex64), x)
    with pytest.raises(TypeError):
        plan.transform(np.zeros(8, np.complex128), np.zeros(8, np.complex128))


@pytest.mark.parametrize("n,dt", [(1 << 22, np.float32), (1 << 23, np.float32), (1 << 22, np.float64)])
def test_c2c_large(comparand, n, dt):
    x = rand_cpx(np.random.default_rng(9), n, dt)
    got, plan = gpu_c2c(x)
    assert oracle.rel_l2(got, comparand.c2c(x)) <= oracle.tolerance(n, dt), plan.describe()


def test_c3_2pow24_double(checkers):
    """BASELINE config C3: N = 2^24 double, against the reference's own kernels via the factory hook."""
    ref = checkers[0]
    if ref is None:
        pytest.skip("compiled reference not present")
    n = 1 << 24
    x = rand_cpx(np.random.default_rng(24), n, np.float64)
    got, plan = gpu_c2c(x)
    assert oracle.rel_l2(got, ref.c2c(x)) <= oracle.tolerance(n, np.float64), plan.describe()
    back, _ = gpu_c2c(got, True)
    assert np.max(np.abs(back / n - x)) <= 1e-8 + n * 1e-12


def test_c2_full_size_properties(comparand):
    """BASELINE config C2 at full size (N=4096 x 2^16, fp32): sampled rows vs the reference, round trip,
    linearity and Parseval over the whole batch."""
    n, batch = 4096, 1 << 16
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.view_as_complex(torch.rand((batch, n, 2), generator=gen, device="cuda") * 2 - 1)
    plan = g.FFT(n, np.float32, batch=batch)
    y = torch.empty_like(x)
    plan.transform(y, x)
    rows = [0, 1, 4095, 32768, batch - 1]
    xs = x[rows].cpu().numpy()
    ys = y[rows].cpu().numpy()
    for r in range(len(rows)):
        assert oracle.rel_l2(ys[r], comparand.c2c(xs[r])) <= oracle.tolerance(n, np.float32)
    # Parseval: sum |X|^2 = n * sum |x|^2
    ex = (x.abs().double() ** 2).sum().item()
    ey = (y.abs().double() ** 2).sum().item()
    assert abs(ey / (n * ex) - 1) < 1e-5
    # round trip
    z = torch.empty_like(x)
    plan.transform(z, y, True)
    assert (z / n - x).abs().max().item() <= fft_eps(n, np.float32)
    # linearity: F(x + 2*roll(x)) == F(x) + 2*F(roll(x)) along the batch axis
    x2 = x + 2 * torch.roll(x, 1, 0)
    y2 = torch.empty_like(x)
    plan.transform(y2, x2)
    want = y + 2 * torch.roll(y, 1, 0)
    rel = ((y2 - want).abs().double() ** 2).sum().sqrt().item() / (want.abs().double() ** 2).sum().sqrt().item()
    assert rel <= oracle.tolerance(n, np.float32)


@pytest.mark.parametrize("lg,dt", [(25, np.float32), (27, np.float32), (25, np.float64)])
def test_sizes_beyond_the_reference_maximum(lg, dt):
    """N = 2^25 ... 2^27 (the reference stops at 2^23): a spectral line lands in the right bin with the right
    amplitude, Parseval holds, and the unscaled inverse returns n * x."""
    n = 1 << lg
    cd = torch.complex64 if dt == np.float32 else torch.complex128
    plan = g.FFT(n, dt)
    k0 = 123457 % n
    t = torch.arange(n, device="cuda", dtype=torch.float64)
    gen = torch.Generator(device="cuda").manual_seed(lg)
    noise = torch.view_as_complex(torch.rand((n, 2), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1)
    phase = 2 * np.pi * ((k0 * t) % n) / n
    x = (torch.polar(torch.ones_like(phase), phase) + 1e-3 * noise).to(cd)
    del t, phase, noise
    y = torch.empty_like(x)
    plan.forward(y, x)
    peak = y[k0].item()
    assert abs(peak - n) / n < 1e-4
    ex = (x.abs().double() ** 2).sum().item()
    ey = (y.abs().double() ** 2).sum().item()
    assert abs(ey / (n * ex) - 1) < (1e-5 if dt == np.float32 else 1e-12)
    z = torch.empty_like(x)
    plan.inverse(z, y)
    err = (z / n - x).abs().max().item()
    assert err < (1e-4 if dt == np.float32 else 1e-12), plan.describe()


@pytest.mark.parametrize("dt,n,batch", [(np.float32, 1024, 1), (np.float32, 4096, 700), (np.float32, 16, 5),
                                        (np.float64, 256, 1000), (np.float64, 8192, 3)])
def test_cached_launch_matches_the_general_driver(dt, n, batch):
    """Single-pass plans resolve their launch at plan creation (FastPath, plan.h); an execution under a changed
    GENFFT_CUDA_* environment goes through the general pass driver instead.  Both must launch the same kernel on the
    same grid: bit-identical output and one launch each -- forward, inverse, in place, and from an input that is only
    8-byte aligned (no TMA prefetch)."""
    import os
    if os.environ.get("GENFFT_TEST_BACKEND") == "emu":
        batch = min(batch, 40)  # the emulator has 3 SMs: 12 tiles already take the TMA-prefetch mode
    tdt = torch.complex64 if dt == np.float32 else torch.complex128
    rng = np.random.default_rng(n + batch)
    x = torch.from_numpy(rand_cpx(rng, batch * n + 1, dt)).cuda()
    plan = g.FFT(n, dt, batch=batch)

    def run(off, inv, in_place=False):
        buf = torch.empty(batch * n + 1, dtype=tdt, device="cuda")
        src = buf[off:off + batch * n]  # off = 1: the input is only 8-byte aligned in float
        src.copy_(x[:batch * n])
        assert src.data_ptr() % 16 == (8 * off if dt == np.float32 else 0)
        dst = src if in_place else torch.empty(batch * n, dtype=tdt, device="cuda")
        n0 = g.launch_count()
        plan.transform(dst, src, inv)
        torch.cuda.synchronize()
        assert g.launch_count() - n0 == 1
        return dst.clone()

    for off in (0, 1) if dt == np.float32 else (0,):
        for inv in (False, True):
            for in_place in (False, True):
                fast = run(off, inv, in_place)
                os.environ["GENFFT_CUDA_UNRELATED_KNOB"] = "1"  # any change of the knob environment bypasses the cache
                try:
                    general = run(off, inv, in_place)
                finally:
                    del os.environ["GENFFT_CUDA_UNRELATED_KNOB"]
                assert torch.equal(fast, general)
    want = np.fft.fft(x[:batch * n].cpu().numpy().reshape(batch, n).astype(np.complex128), axis=1)
    assert oracle.rel_l2(run(0, False).cpu().numpy().reshape(batch, n), want) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_host_path_leaves_the_gaps_between_transforms_alone(dt):
    """Batched host-pointer execution with out_dist > n: the caller's memory between two outputs is not the
    library's to write (found by tools/emu_fuzz.py: the staged chunk used to be copied back as one linear range)."""
    n, batch, in_dist, out_dist = 256, 13, 258, 261
    rng = np.random.default_rng(5)
    x = rand_cpx(rng, batch * in_dist, dt)
    out = np.full(batch * out_dist, 7.5 - 3.25j, CPX[dt])
    g.FFT(n, dt, batch=batch, in_dist=in_dist, out_dist=out_dist).transform(out, x)
    got = out.reshape(batch, out_dist)
    want = np.fft.fft(x.reshape(batch, in_dist)[:, :n].astype(np.complex128), axis=1)
    assert oracle.rel_l2(got[:, :n], want) <= oracle.tolerance(n, dt)
    assert np.all(got[:, n:] == 7.5 - 3.25j)
    r = rng.uniform(-1, 1, (batch, n)).astype(dt)
    rout = np.full((batch, n // 2 + 4), 7.5 - 3.25j, CPX[dt])
    g.RealFFT(n, dt, half=True, batch=batch, out_dist=n // 2 + 4).forward(rout, r)
    assert oracle.rel_l2(rout[:, :n // 2 + 1], np.fft.rfft(r.astype(np.float64), axis=1)) <= oracle.tolerance(n, dt)
    assert np.all(rout[:, n // 2 + 1:] == 7.5 - 3.25j)


def test_c_loop_pair_timing_hook():
    """genfft_cuda_debug_time_c2c_pairs (bench.py's C1 figure from a C loop): runs the pairs and leaves the round trip."""
    import ctypes
    n = 1024
    x = torch.from_numpy(rand_cpx(np.random.default_rng(1), n, np.float32)).cuda()
    y, z = torch.empty_like(x), torch.empty_like(x)
    plan = g.FFT(n, np.float32)
    us = ctypes.c_double(0.0)
    st = torch.cuda.current_stream().cuda_stream
    n0 = g.launch_count()
    assert g.lib().genfft_cuda_debug_time_c2c_pairs(plan._h, z.data_ptr(), y.data_ptr(), x.data_ptr(), 20, st, ctypes.byref(us)) == 0
    torch.cuda.synchronize()
    assert g.launch_count() - n0 == 2 * (20 + 2) and us.value > 0
    assert float((z / n - x).abs().max()) < 1e-5


@pytest.mark.no_emu  # needs real streams
def test_chained_plan_on_two_streams_at_once(comparand):
    """One chained (two-pass) plan executed on two streams concurrently: every stream has its own ticket / group
    counters, so neither execution resets the other's mid-kernel (the counters used to be one per plan)."""
    n, batch = 1 << 16, 64
    rng = np.random.default_rng(11)
    xa, xb = rand_cpx(rng, (batch, n), np.float32), rand_cpx(rng, (batch, n), np.float32)
    plan = g.FFT(n, np.float32, batch=batch)
    da, db = torch.from_numpy(xa).cuda(), torch.from_numpy(xb).cuda()
    oa, ob = torch.empty_like(da), torch.empty_like(db)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(6):
        with torch.cuda.stream(sa):
            plan.forward(oa, da)
        with torch.cuda.stream(sb):
            plan.forward(ob, db)
    torch.cuda.synchronize()
    want_a = np.fft.fft(xa[:4].astype(np.complex128), axis=1)
    assert oracle.rel_l2(oa[:4].cpu().numpy(), want_a) <= oracle.tolerance(n, np.float32)
    ref_a, ref_b = torch.empty_like(da), torch.empty_like(db)
    plan.forward(ref_a, da)
    plan.forward(ref_b, db)
    torch.cuda.synchronize()
    assert torch.equal(oa, ref_a) and torch.equal(ob, ref_b)


def test_host_pointers_pinned_chunked_two_pass(comparand, monkeypatch):
    """Host-pointer execution of a chained two-pass plan from PINNED memory in several chunks: the chunks alternate
    between two streams and really overlap, each with its own chain counters."""
    monkeypatch.setenv("GENFFT_CUDA_HOST_CHUNK_MB", "1")
    n, batch = 1 << 16, 24
    x = torch.from_numpy(rand_cpx(np.random.default_rng(3), (batch, n), np.float32))
    out = torch.empty_like(x)
    if torch.cuda.is_available() and os.environ.get("GENFFT_TEST_BACKEND") != "emu":
        x, out = x.pin_memory(), out.pin_memory()
    plan = g.FFT(n, np.float32, batch=batch)
    for rep in range(3):
        plan.forward(out, x)
    want = np.fft.fft(x.numpy().astype(np.complex128), axis=1)
    assert oracle.rel_l2(out.numpy(), want) <= oracle.tolerance(n, np.float32)
    assert oracle.rel_l2(out.numpy()[:2], comparand.c2c_batch(x.numpy()[:2])) <= oracle.tolerance(n, np.float32)


def test_host_pointer_calls_from_several_threads_on_one_plan(comparand):
    """The reference's impl objects are stateless and safe to call concurrently; the host-pointer path serialises per
    plan instead of corrupting its staging buffers."""
    import threading
    n, batch = 4096, 8
    plan = g.FFT(n, np.float32, batch=batch)
    rng = np.random.default_rng(9)
    xs = [rand_cpx(rng, (batch, n), np.float32) for _ in range(4)]
    outs = [np.empty_like(x) for x in xs]
    errs = []

    def work(k):
        try:
            for _ in range(5):
                plan.forward(outs[k], xs[k])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for x, o in zip(xs, outs):
        assert oracle.rel_l2(o, np.fft.fft(x.astype(np.complex128), axis=1)) <= oracle.tolerance(n, np.float32)


def test_distances_smaller_than_a_transform_are_rejected():
    for kw in (dict(in_dist=100), dict(out_dist=255), dict(in_dist=-256)):
        with pytest.raises(g.GenfftCudaError):
            g.FFT(256, np.float32, batch=4, **kw)
    with pytest.raises(g.GenfftCudaError):
        g.RealFFT(256, np.float32, half=True, batch=4, out_dist=100)
    with pytest.raises(g.GenfftCudaError):
        g.InverseRealFFT(256, np.float32, batch=4, out_dist=128)
    g.FFT(256, np.float32, batch=1, in_dist=1)  # a single transform has no distance to violate
    with pytest.raises(ValueError):
        g.DIT(64, np.float32).apply(torch.zeros(8, dtype=torch.complex64, device="cuda"),
                                    torch.zeros(32, dtype=torch.complex64, device="cuda"), False)
    with pytest.raises(ValueError):
        g.FFTVert(64, np.float32).transform_no_scramble(torch.zeros(64, dtype=torch.complex64, device="cuda"), 4, 4)


def test_plan_is_bound_to_its_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    plan = g.FFT(1024, np.float32)
    x = torch.zeros(1024, dtype=torch.complex64, device="cuda:1")
    with torch.cuda.device(1):
        with pytest.raises(g.GenfftCudaError):
            plan.forward(torch.empty_like(x), x)


def test_host_buffers_near_the_gpu(comparand):
    """genfft_cuda_host_alloc (page-locked, bound to the GPU's NUMA node when the host has several) as the caller's
    buffers of the host-pointer path."""
    from genfft_b200.hostmem import PinnedNearGpu
    n, batch = 4096, 16
    bx, by = PinnedNearGpu((batch, n), np.complex64), PinnedNearGpu((batch, n), np.complex64)
    assert bx.numa_node >= -1 and bx.array.shape == (batch, n)
    bx.array[...] = rand_cpx(np.random.default_rng(4), (batch, n), np.float32)
    g.FFT(n, np.float32, batch=batch).forward(by.array, bx.array)
    assert oracle.rel_l2(by.array, comparand.c2c_batch(bx.array)) <= oracle.tolerance(n, np.float32)
    ptr = bx.ptr
    bx.close()
    by.close()
    assert g.lib().genfft_cuda_host_free(ptr) != 0  # not (or no longer) one of its blocks: an error, not a crash
