"""CPU tests (-m "not gpu"): pin the oracle before trusting it.

1. The C restatement (oracle/genfft_oracle.c) is BIT-EXACT against the committed outputs of the reference's
   generic scalar back-end (tests/golden/, generated from /root/reference by tests/golden/make_golden.py) and
   within float rounding of the reference's best-ISA dispatch back-end.
2. When the compiled reference is present (this container, or its prebuilt .so on the GPU box) the same is
   re-checked live, plus the reference's own acceptance tests restated: reference_impl::FFT_pow2 vs naive DFT
   (test/test_reference.cpp:31-65) and genFFT vs FFT_pow2 within FFT_Eps (test/fft_test_impl.h:35-58,
   test/test_util.h:62-72).
"""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "genfft_golden.npz")
DT = {"f32": np.float32, "f64": np.float64}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def port(checkers):
    return checkers[1]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_golden_file_is_complete(gold):
    assert len(gold.files) == 258


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_c2c_bit_exact_vs_reference_generic(gold, port, tag):
    for key in [k for k in gold.files if k.startswith(f"c2c_{tag}_") and k.endswith("_in")]:
        n = int(key.split("_")[2])
        x = gold[key]
        for inv in (0, 1):
            got = port.c2c(x, bool(inv))
            assert np.array_equal(bits(got), bits(gold[f"c2c_{tag}_{n}_gen_{inv}"])), (key, inv)
            assert oracle.rel_l2(got, gold[f"c2c_{tag}_{n}_disp_{inv}"]) <= oracle.tolerance(n, DT[tag]) / 10


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_r2c_bit_exact_vs_reference_generic(gold, port, tag):
    for key in [k for k in gold.files if k.startswith(f"r2c_{tag}_") and k.endswith("_in")]:
        n = int(key.split("_")[2])
        x = gold[key]
        for half in (0, 1):
            lim = 1 if n == 1 else (n // 2 + 1 if half else n)
            got = port.r2c(x, bool(half), fill=43 + 21j)
            assert np.array_equal(bits(got[:lim]), bits(gold[f"r2c_{tag}_{n}_gen_{half}"])), (key, half)
            assert np.all(got[lim:] == 43 + 21j)  # nothing written past n/2+1 (test/fft_test_impl.h:102-105)
            assert oracle.rel_l2(got[:lim], gold[f"r2c_{tag}_{n}_disp_{half}"]) <= oracle.tolerance(n, DT[tag]) / 10


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_vert_and_2d_vs_golden(gold, port, tag):
    for key in [k for k in gold.files if k.startswith(f"vert_{tag}_") and k.endswith("_in")]:
        shape = key.split("_")[2]
        got = port.vert(gold[key])
        assert np.array_equal(bits(got), bits(gold[f"vert_{tag}_{shape}_gen"])), key
        assert oracle.rel_l2(got, gold[f"vert_{tag}_{shape}_disp"]) <= oracle.tolerance(got.shape[0], DT[tag]) / 10
    for key in [k for k in gold.files if k.startswith(f"fft2d_{tag}_") and k.endswith("_in")]:
        shape = key.split("_")[2]
        x = gold[key]
        for inv in (0, 1):
            got = port.fft2d(x, bool(inv))
            assert oracle.rel_l2(got, gold[f"fft2d_{tag}_{shape}_disp_{inv}"]) <= oracle.tolerance(x.size, DT[tag]) / 10
        ref64 = np.fft.fft2(x.astype(np.complex128))
        assert oracle.rel_l2(port.fft2d(x), ref64) <= oracle.tolerance(x.size, DT[tag])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_port_against_numpy_conventions(port, dt):
    """Sign / scaling / layout conventions: forward = numpy fft, inverse = ifft * n (unscaled)."""
    rng = np.random.default_rng(0)
    cd = np.complex64 if dt == np.float32 else np.complex128
    for n in (2, 8, 64, 2048):
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
        x64 = x.astype(np.complex128)
        assert oracle.rel_l2(port.c2c(x), np.fft.fft(x64)) <= oracle.tolerance(n, dt)
        assert oracle.rel_l2(port.c2c(x, True), np.fft.ifft(x64) * n) <= oracle.tolerance(n, dt)
        r = rng.uniform(-1, 1, n).astype(dt)
        assert oracle.rel_l2(port.r2c(r, True)[: n // 2 + 1], np.fft.rfft(r.astype(np.float64))) <= oracle.tolerance(n, dt)
        assert oracle.rel_l2(port.r2c(r, False), np.fft.fft(r.astype(np.float64))) <= oracle.tolerance(n, dt)
        assert oracle.rel_l2(port.transform_real(r), np.fft.fft(r.astype(np.float64))) <= oracle.tolerance(n, dt)
        r2 = rng.uniform(-1, 1, n).astype(dt)
        a, b = port.two_real(r, r2)
        assert oracle.rel_l2(a, np.fft.fft(r.astype(np.float64))) <= oracle.tolerance(n, dt)
        assert oracle.rel_l2(b, np.fft.fft(r2.astype(np.float64))) <= oracle.tolerance(n, dt)
    # transform_no_scramble: natural-order result from bit-reversed input
    n = 256
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
    perm = np.array([int(format(i, "08b")[::-1], 2) for i in range(n)])
    scr = np.empty_like(x)
    scr[perm] = x
    assert oracle.rel_l2(port.c2c_no_scramble(scr), np.fft.fft(x.astype(np.complex128))) <= oracle.tolerance(n, dt)


def test_port_rejects_bad_sizes(port):
    with pytest.raises(ValueError):
        port.c2c(np.zeros(3, np.complex64))
    with pytest.raises(ValueError):
        port.r2c(np.zeros(12, np.float32))


# ---- live checks against the compiled reference (skipped when its .so did not travel) --------------------

@pytest.fixture(scope="module")
def ref(checkers):
    if checkers[0] is None:
        pytest.skip("oracle/_ref/libgenfft_ref.so not present")
    return checkers[0]


def test_reference_build_matches_golden(gold, ref):
    """The golden fixtures are reproducible from the reference build at hand (same ISA path or not, within
    rounding; the generic back-end bit for bit)."""
    gen = oracle.RefGeneric()
    for n in (8, 256, 4096):
        x = gold[f"c2c_f32_{n}_in"]
        assert np.array_equal(x, ref.dummy_complex(n, np.float32))  # DummyData is deterministic
        assert np.array_equal(bits(gen.c2c(x)), bits(gold[f"c2c_f32_{n}_gen_0"]))
        assert oracle.rel_l2(ref.c2c(x), gold[f"c2c_f32_{n}_disp_0"]) <= 1e-6


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_port_bit_exact_vs_live_generic_reference(port, ref, dt):
    gen = oracle.RefGeneric()
    rng = np.random.default_rng(11)
    cd = np.complex64 if dt == np.float32 else np.complex128
    for n in (1, 2, 4, 8, 32, 512, 8192, 1 << 16):
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
        for inv in (False, True):
            assert np.array_equal(bits(port.c2c(x, inv)), bits(gen.c2c(x, inv))), (n, inv)
        r = rng.uniform(-1, 1, n).astype(dt)
        for half in (False, True):
            assert np.array_equal(bits(port.r2c(r, half)), bits(gen.r2c(r, half))), (n, half)
    x = (rng.uniform(-1, 1, (128, 37)) + 1j * rng.uniform(-1, 1, (128, 37))).astype(cd)
    assert np.array_equal(bits(port.vert(x)), bits(gen.vert(x)))
    assert np.array_equal(bits(port.vert(x, True)), bits(gen.vert(x, True)))


def test_reference_in_test_comparand_vs_naive_dft(ref):
    """test/test_reference.cpp:31-65: FFT_pow2 against the O(n^2) DFT in double, abs eps 1e-10 + 1e-11 n."""
    for n in (2, 4, 8, 16, 32, 64, 128, 256, 512):
        x = ref.dummy_complex(n, np.float64)
        for inv in (False, True):
            a, b = ref.testref_fft_pow2(x, inv), ref.testref_dft(x, inv)
            assert np.max(np.abs(a - b)) <= 1e-10 + 1e-11 * n


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_reference_passes_its_own_acceptance(ref, port, dt):
    """TestFFT_Pow2 / TestRealFFT_Pow2 restated (test/fft_test_impl.h:35-58,84-106): genFFT (dispatch) and
    the restatement both stay within FFT_Eps of reference_impl::FFT_pow2."""
    for lg in range(1, 17):
        n = 1 << lg
        eps = (1e-5 + n * 1e-8) if dt == np.float32 else (1e-8 + n * 1e-12)
        x = ref.dummy_complex(n, dt)
        want = ref.testref_fft_pow2(x)
        for impl in (ref, port):
            got = impl.c2c(x)
            assert np.max(np.abs(got.real - want.real)) <= eps and np.max(np.abs(got.imag - want.imag)) <= eps
            back = impl.c2c(got, True)
            assert np.max(np.abs((back.astype(np.complex128) / n).real - x.real)) <= eps
        r = ref.dummy_real(n, dt)
        want_r = ref.testref_fft_pow2(r.astype(x.dtype))
        for half in (True, False):
            lim = n // 2 + 1 if half else n
            for impl in (ref, port):
                got = impl.r2c(r, half, fill=43 + 21j)
                assert np.max(np.abs(got[:lim] - want_r[:lim])) <= eps * 1.5
                assert np.all(got[lim:] == 43 + 21j)


def test_reference_2pow24_hook(ref):
    """The factory hook that lifts the reference past its 2^23 switch (oracle/ref_native_big.cpp) agrees with a
    double-precision numpy FFT; this is the C3 comparand."""
    n = 1 << 24
    rng = np.random.default_rng(3)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))
    assert oracle.rel_l2(ref.c2c(x), np.fft.fft(x)) <= 1e-14


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_real_fft2d_restatement_and_reference_match_numpy(checkers, dt):
    """RealFFT2D::forward and forward_2x (FFTReal.h:83-118) have no test in the reference: pin the restatement and,
    where present, the compiled reference (the comparands of the GPU parity tests) against numpy's fft2 and against
    each other (the reference's RealFFT2D has no factory parameter, so it always runs the dispatch back-end: equal
    within rounding, identical where no twiddle is involved)."""
    ref, port = checkers
    for w, h in [(1, 1), (2, 1), (2, 2), (4, 2), (8, 8), (64, 16), (16, 128), (512, 32)]:
        rng = np.random.default_rng(w * 31 + h)
        a = rng.uniform(-1, 1, (h, w)).astype(dt)
        b = rng.uniform(-1, 1, (h, w)).astype(dt)
        tol = oracle.tolerance(w * h, dt) / 10
        want = np.fft.fft2(a.astype(np.float64))
        want2 = np.fft.fft2(a.astype(np.float64) + 1j * b.astype(np.float64))
        for impl in (port, ref):
            if impl is None:
                continue
            assert oracle.rel_l2(impl.real_fft2d(a), want) <= tol, (w, h)
            assert oracle.rel_l2(impl.real_fft2d_2x(a, b), want2) <= tol, (w, h)
        if ref is not None:
            assert oracle.rel_l2(port.real_fft2d(a), ref.real_fft2d(a)) <= tol
            assert oracle.rel_l2(port.real_fft2d_2x(a, b), ref.real_fft2d_2x(a, b)) <= tol
            if w * h <= 8:  # butterflies without multiplications: the two agree bit for bit
                assert np.array_equal(bits(port.real_fft2d(a)), bits(ref.real_fft2d(a)))


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_real_fft2d_vs_reference_fixtures(port, tag):
    """The restatement of RealFFT2D::forward / forward_2x against the committed outputs of the reference (dispatch
    back-end -- RealFFT2D has no factory parameter, so no generic-back-end output exists to be bit-exact with)."""
    gold2 = np.load(os.path.join(os.path.dirname(GOLD), "genfft_golden_real2d.npz"))
    keys = [k for k in gold2.files if k.startswith(f"real2d_{tag}_") and k.endswith("_in1")]
    assert len(keys) == 6
    for key in keys:
        w, h = (int(v) for v in key.split("_")[2].split("x"))
        a, b = gold2[key], gold2[key.replace("_in1", "_in2")]
        tol = oracle.tolerance(w * h, DT[tag]) / 10
        assert oracle.rel_l2(port.real_fft2d(a), gold2[key.replace("_in1", "_forward")]) <= tol, key
        assert oracle.rel_l2(port.real_fft2d_2x(a, b), gold2[key.replace("_in1", "_forward_2x")]) <= tol, key
        assert oracle.rel_l2(gold2[key.replace("_in1", "_forward")], np.fft.fft2(a.astype(np.float64))) <= tol
