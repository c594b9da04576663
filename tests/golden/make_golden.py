#!/usr/bin/env python
"""Generates tests/golden/genfft_golden.npz from the UNMODIFIED reference (run in the build container,
where /root/reference exists):  python tests/golden/make_golden.py

Inputs are the reference's own test data (DummyData: default-seeded std::mt19937_64,
uniform_real_distribution<T>(-1,1), test/test_util.h:36-57); outputs come from the reference compiled in
place (oracle/_ref/libgenfft_ref.so): the dispatch back-end ("disp", the best CPU ISA: the parity target)
and the generic scalar back-end ("gen", the bit-exact pin of the C restatement).  The reference has no
stored golden vectors of its own; these fixtures play that role on the GPU box,
where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

SIZES_1D = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 4096]
SIZES_R2C = [1, 2, 4, 8, 16, 64, 256, 1024, 4096]
VERT = [(2, 3), (4, 7), (8, 33), (64, 5), (256, 9)]
TWO_D = [(4, 8), (64, 32), (128, 16)]  # (width, height)
REAL_2D = [(2, 2), (4, 2), (8, 8), (64, 32), (128, 16), (16, 64)]  # (width, height)


def main():
    oracle.build("all")
    ref, gen = oracle.Ref(), oracle.RefGeneric()
    out = {}
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for n in SIZES_1D:
            x = ref.dummy_complex(n, dt)
            out[f"c2c_{tag}_{n}_in"] = x
            for inv in (0, 1):
                out[f"c2c_{tag}_{n}_disp_{inv}"] = ref.c2c(x, bool(inv))
                out[f"c2c_{tag}_{n}_gen_{inv}"] = gen.c2c(x, bool(inv))
        for n in SIZES_R2C:
            x = ref.dummy_real(n, dt)
            out[f"r2c_{tag}_{n}_in"] = x
            for half in (0, 1):
                lim = 1 if n == 1 else (n // 2 + 1 if half else n)
                out[f"r2c_{tag}_{n}_disp_{half}"] = ref.r2c(x, bool(half))[:lim]
                out[f"r2c_{tag}_{n}_gen_{half}"] = gen.r2c(x, bool(half))[:lim]
        for n, cols in VERT:
            x = ref.dummy_complex(n * cols, dt).reshape(n, cols)
            out[f"vert_{tag}_{n}x{cols}_in"] = x
            out[f"vert_{tag}_{n}x{cols}_disp"] = ref.vert(x)
            out[f"vert_{tag}_{n}x{cols}_gen"] = gen.vert(x)
        for w, h in TWO_D:
            x = ref.dummy_complex(w * h, dt).reshape(h, w)
            out[f"fft2d_{tag}_{w}x{h}_in"] = x
            for inv in (0, 1):
                out[f"fft2d_{tag}_{w}x{h}_disp_{inv}"] = ref.fft2d(x, bool(inv))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "genfft_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB; {ref.describe()}")
    # RealFFT2D::forward / forward_2x (FFTReal.h:83-118), added later: a second file so that the first stays as it was
    out = {}
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for w, h in REAL_2D:
            ab = ref.dummy_real(2 * w * h, dt)
            a, b = ab[: w * h].reshape(h, w), ab[w * h:].reshape(h, w)
            out[f"real2d_{tag}_{w}x{h}_in1"] = a
            out[f"real2d_{tag}_{w}x{h}_in2"] = b
            out[f"real2d_{tag}_{w}x{h}_forward"] = ref.real_fft2d(a)
            out[f"real2d_{tag}_{w}x{h}_forward_2x"] = ref.real_fft2d_2x(a, b)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "genfft_golden_real2d.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
