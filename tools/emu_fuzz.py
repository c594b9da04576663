"""Randomised sweep of the C ABI on the kernel-logic emulator (tests/emu): sizes, batches, distances, strides, odd
widths, in-place calls, all entry points, checked against numpy in double precision.  Device allocations of the emulated
library sit against guard pages, so an out-of-bounds access of plan-owned memory faults; user buffers are padded with
sentinels that must survive.   usage: python tools/emu_fuzz.py [seed] [cases]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("GENFFT_TEST_BACKEND", "emu")
from emu import backend
backend.install()
import torch
import genfft_b200 as g

CPX = {np.float32: np.complex64, np.float64: np.complex128}
TOL = {np.float32: 2e-5, np.float64: 2e-13}
BIG = int(os.environ.get("GENFFT_FUZZ_BIG", "0"))  # extra bits of size: 3 reaches the three-pass plans (slow)
SENT = 7.5 - 3.25j


def rc(rng, shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(CPX[dt])


def rel(a, b):
    d = np.linalg.norm(np.asarray(a, np.complex128).ravel() - np.asarray(b, np.complex128).ravel())
    return d / max(np.linalg.norm(np.asarray(b, np.complex128).ravel()), 1e-300)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def case_c2c(rng, dt):
    n = 1 << int(rng.integers(0, 18 + BIG)); batch = int(rng.integers(1, max(2, min(64, (1 << (18 + BIG)) // n) + 1)))
    ind, outd = n + int(rng.integers(0, 4)), n + int(rng.integers(0, 4)); inv = bool(rng.integers(0, 2))
    buf = rc(rng, batch * ind + 2, dt); out = np.full(batch * outd + 2, SENT, CPX[dt])
    off = int(rng.integers(0, 2)) if dt == np.float32 else 0
    plan = g.FFT(n, dt, batch=batch, in_dist=ind, out_dist=outd)
    dev = bool(rng.integers(0, 2))
    src, dst = buf[off:off + batch * ind], out[:batch * outd]
    if dev: plan.transform(t(dst), t(src), inv)
    else: plan.transform(dst, src, inv)
    x = src.reshape(batch, ind)[:, :n].astype(np.complex128)
    want = np.fft.ifft(x, axis=1) * n if inv else np.fft.fft(x, axis=1)
    got = dst.reshape(batch, outd)
    assert rel(got[:, :n], want) < TOL[dt] * max(1, np.log2(max(n, 2))), ("c2c", n, batch, ind, outd, inv, dev, plan.describe())
    assert np.all(got[:, n:] == SENT) and np.all(out[batch * outd:] == SENT), ("c2c pad", n, batch)
    if ind == outd and rng.random() < 0.5:  # in place on the device path
        w = src.copy(); plan.transform(t(w), t(w), inv)
        assert rel(w.reshape(batch, ind)[:, :n], want) < TOL[dt] * max(1, np.log2(max(n, 2))), ("c2c in place", n, batch)


def case_r2c(rng, dt):
    n = 1 << int(rng.integers(0, 19 + BIG)); batch = int(rng.integers(1, max(2, min(40, (1 << (18 + BIG)) // n) + 1)))
    half = bool(rng.integers(0, 2)); lim = 1 if n == 1 else (n // 2 + 1 if half else n)
    outd = lim + int(rng.integers(0, 3)); ind = n + 2 * int(rng.integers(0, 3)) if n >= 2 else n + int(rng.integers(0, 3))
    x = rng.uniform(-1, 1, (batch, ind)).astype(dt); out = np.full((batch, outd), SENT, CPX[dt])
    plan = g.RealFFT(n, dt, half=half, batch=batch, in_dist=ind, out_dist=outd)
    if rng.integers(0, 2): plan.forward(t(out), t(x))
    else: plan.forward(out, x)
    want = np.fft.fft(x[:, :n].astype(np.float64), axis=1)[:, :lim]
    assert rel(out[:, :lim], want) < TOL[dt] * max(1, np.log2(max(n, 2))), ("r2c", n, batch, half, ind, outd, plan.describe())
    assert np.all(out[:, lim:] == SENT), ("r2c pad", n, batch, half)
    if half and n >= 2:
        inv = g.InverseRealFFT(n, dt, batch=batch, in_dist=outd, out_dist=ind if ind % 2 == 0 else n)
        od = ind if ind % 2 == 0 else n
        back = np.full((batch, od), 5.5, dt)
        inv.inverse(t(back), t(out))
        assert rel(back[:, :n], x[:, :n].astype(np.float64) * n) < 2 * TOL[dt] * max(1, np.log2(n)), ("c2r", n, batch)
        assert np.all(back[:, n:] == 5.5), ("c2r pad", n)


def case_vert(rng, dt):
    h = 1 << int(rng.integers(0, 13)); cols = int(rng.integers(1, max(2, min(200, (1 << 17) // h))))
    si, so = cols + int(rng.integers(0, 4)), cols + int(rng.integers(0, 4)); inv = bool(rng.integers(0, 2))
    x = rc(rng, (h, si), dt); out = np.full((h, so), SENT, CPX[dt])
    plan = g.FFTVert(h, dt)
    if rng.integers(0, 2): plan.transform(t(out), t(x), cols, out_stride=so, in_stride=si, inv=inv)
    else: plan.transform(out, x, cols, out_stride=so, in_stride=si, inv=inv)
    x64 = x[:, :cols].astype(np.complex128)
    want = np.fft.ifft(x64, axis=0) * h if inv else np.fft.fft(x64, axis=0)
    assert rel(out[:, :cols], want) < TOL[dt] * max(1, np.log2(max(h, 2))), ("vert", h, cols, si, so, inv, plan.describe())
    assert np.all(out[:, cols:] == SENT), ("vert pad", h, cols)


def case_2d(rng, dt):
    lw, lh = int(rng.integers(0, 11)), int(rng.integers(0, 11))
    w, h = 1 << lw, 1 << lh; si, so = w + int(rng.integers(0, 3)), w + int(rng.integers(0, 3)); inv = bool(rng.integers(0, 2))
    x = rc(rng, (h, si), dt); out = np.full((h, so), SENT, CPX[dt])
    plan = g.FFT2D(w, h, dt)
    if rng.integers(0, 2): plan.transform(t(out), t(x), out_stride=so, in_stride=si, inv=inv)
    else: plan.transform(out, x, out_stride=so, in_stride=si, inv=inv)
    x64 = x[:, :w].astype(np.complex128)
    want = np.fft.ifft2(x64) * (w * h) if inv else np.fft.fft2(x64)
    assert rel(out[:, :w], want) < TOL[dt] * max(1, lw + lh), ("2d", w, h, si, so, inv, plan.describe())
    assert np.all(out[:, w:] == SENT), ("2d pad", w, h)


def case_real2d(rng, dt):
    lw, lh = int(rng.integers(1, 10)), int(rng.integers(0, 10))
    w, h = 1 << lw, 1 << lh; so = w + int(rng.integers(0, 3)); si = w + 2 * int(rng.integers(0, 2))
    a = rng.uniform(-1, 1, (h, si)).astype(dt); b = rng.uniform(-1, 1, (h, si)).astype(dt)
    plan = g.RealFFT2D(w, h, dt); out = np.full((h, so), SENT, CPX[dt])
    if rng.integers(0, 2):
        plan.forward(t(out), t(a), out_stride=so, in_stride=si)
        want = np.fft.fft2(a[:, :w].astype(np.float64))
    else:
        plan.forward_2x(t(out), t(a), t(b), out_stride=so, in_stride1=si, in_stride2=si)
        want = np.fft.fft2(a[:, :w].astype(np.float64) + 1j * b[:, :w].astype(np.float64))
    assert rel(out[:, :w], want) < TOL[dt] * max(1, lw + lh), ("real2d", w, h, si, so, plan.describe())
    assert np.all(out[:, w:] == SENT), ("real2d pad", w, h)


def bitrev_perm(n):
    lg = max(n.bit_length() - 1, 0)
    idx = np.arange(n)
    rev = np.zeros(n, dtype=np.int64)
    for b in range(lg):
        rev |= ((idx >> b) & 1) << (lg - 1 - b)
    return rev


def case_misc_1d(rng, dt):
    n = 1 << int(rng.integers(1, 16)); lg = n.bit_length() - 1
    plan = g.FFT(n, dt)
    tol = TOL[dt] * max(1, lg)
    which = int(rng.integers(0, 4))
    dev = bool(rng.integers(0, 2))
    w = (lambda a: t(a)) if dev else (lambda a: a)
    if which == 0:  # transform_no_scramble: bit-reversed input, in place (fft.h:69-73)
        x = rc(rng, n, dt); inv = bool(rng.integers(0, 2))
        buf = x[bitrev_perm(n)].copy()
        plan.transform_no_scramble(w(buf), inv)
        want = np.fft.ifft(x.astype(np.complex128)) * n if inv else np.fft.fft(x.astype(np.complex128))
        assert rel(buf, want) < tol, ("no_scramble", n, inv, dev)
    elif which == 1:  # transform_real (fft.h:90-94)
        r = rng.uniform(-1, 1, n).astype(dt); out = np.full(n + 1, SENT, CPX[dt])
        plan.transform_real(w(out[:n]), w(r))
        assert rel(out[:n], np.fft.fft(r.astype(np.float64))) < tol and out[n] == SENT, ("real_in", n, dev)
    elif which == 2:  # transform_interleave + separate_2x_real_FFT (fft.h:100-105, FFTReal.h:35-66)
        a = rng.uniform(-1, 1, n).astype(dt); b = rng.uniform(-1, 1, n).astype(dt)
        z = np.zeros(n, CPX[dt])
        plan.transform_interleave(w(z), w(a), w(b))
        assert rel(z, np.fft.fft(a.astype(np.float64) + 1j * b.astype(np.float64))) < tol, ("interleave", n, dev)
        fa, fb = np.zeros(n, CPX[dt]), np.zeros(n, CPX[dt])
        g.separate_2x_real_FFT(t(fa), t(fb), t(z), n)
        assert rel(fa, np.fft.fft(a.astype(np.float64))) < 2 * tol and rel(fb, np.fft.fft(b.astype(np.float64))) < 2 * tol, ("separate", n)
    else:  # DIT<T>::apply on the packed half-size transform, out may alias in (fft.h:173-196)
        r = rng.uniform(-1, 1, n).astype(dt); half = bool(rng.integers(0, 2))
        zc = np.fft.fft(r.astype(np.float64)[0::2] + 1j * r.astype(np.float64)[1::2]).astype(CPX[dt])
        lim = n // 2 + 1 if half else n
        buf = np.full(max(lim, n // 2) + 1, SENT, CPX[dt]); buf[:n // 2] = zc
        alias = bool(rng.integers(0, 2))
        out = buf if alias else np.full(lim + 1, SENT, CPX[dt])
        g.DIT(n, dt).apply(w(out), w(buf), half)
        assert rel(out[:lim], np.fft.fft(r.astype(np.float64))[:lim]) < 2 * tol, ("dit", n, half, alias, dev)
        assert out[lim] == SENT if not alias else True, ("dit pad", n, half)


def case_vert_no_scramble(rng, dt):
    h = 1 << int(rng.integers(1, 11)); cols = int(rng.integers(1, 70)); stride = cols + int(rng.integers(0, 4))
    inv = bool(rng.integers(0, 2))
    x = rc(rng, (h, stride), dt)
    buf = x[bitrev_perm(h)].copy()
    buf[:, cols:] = SENT
    plan = g.FFTVert(h, dt)
    if rng.integers(0, 2): plan.transform_no_scramble(t(buf), stride, cols, inv)
    else: plan.transform_no_scramble(buf, stride, cols, inv)
    x64 = x[:, :cols].astype(np.complex128)
    want = np.fft.ifft(x64, axis=0) * h if inv else np.fft.fft(x64, axis=0)
    assert rel(buf[:, :cols], want) < TOL[dt] * max(1, np.log2(h)), ("vert_no_scramble", h, cols, stride, inv)
    assert np.all(buf[:, cols:] == SENT), ("vert_no_scramble pad", h, cols)


CASES = [case_c2c, case_c2c, case_r2c, case_r2c, case_vert, case_2d, case_real2d, case_misc_1d, case_misc_1d, case_vert_no_scramble]

if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rng = np.random.default_rng(seed)
    t0 = time.time()
    for k in range(count):
        fn = CASES[int(rng.integers(0, len(CASES)))]
        dt = np.float32 if rng.random() < 0.6 else np.float64
        fn(rng, dt)
    print(f"emu_fuzz seed {seed}: {count} cases ok in {time.time() - t0:.1f} s")
