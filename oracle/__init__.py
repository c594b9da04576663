"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracle.

Two checkers live here:

* ``port``  -- the plain-C restatement of genFFT's algorithm (``genfft_oracle.c``), built by
  ``make -C oracle oracle`` into ``oracle/_build/libgenfft_oracle.so``.
* ``ref``   -- the UNMODIFIED reference compiled from ``/root/reference`` where it lies by
  ``make -C oracle ref`` into ``oracle/_ref/libgenfft_ref.so`` (git-ignored; travels to the GPU box
  as a prebuilt file because ``/root/reference`` does not exist there).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.  ``genfft_b200`` never does: the product has no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libgenfft_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgenfft_ref.so")
REFERENCE_ROOT = "/root/reference"

_c = ctypes
_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CPX = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}


def build(which: str = "all", quiet: bool = True) -> None:
    """Compile the checkers.  ``ref`` is only (re)built when the reference sources are present."""
    targets = []
    if which in ("all", "oracle"):
        targets.append("oracle")
    if which in ("all", "ref") and os.path.isdir(REFERENCE_ROOT):
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-C", HERE, "-j8", *targets], check=True,
                       stdout=subprocess.DEVNULL if quiet else None,
                       stderr=subprocess.DEVNULL if quiet else None)


def _real_dtype(a: np.ndarray) -> np.dtype:
    return np.dtype(np.float32) if a.dtype in (np.float32, np.complex64) else np.dtype(np.float64)


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_c.c_void_p)


class _Lib:
    prefix = ""
    path = ""
    max_log2 = 62

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = _c.CDLL(self.path)

    def _fn(self, name, dtype, restype=_c.c_int, argtypes=None):
        f = getattr(self.lib, f"{self.prefix}{name}_{_SUF[np.dtype(dtype)]}")
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f

    # --- 1D C2C: FFT<T>::transform<inv> (fft.h:80-85) ---
    def c2c(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x)
        rd = _real_dtype(x)
        out = np.empty_like(x)
        f = self._fn("c2c", rd, argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int])
        rc = f(_ptr(out), _ptr(x), x.shape[-1], int(inverse))
        if rc:
            raise ValueError(f"c2c rejected n={x.shape[-1]}")
        return out

    def c2c_batch(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x)
        return np.stack([self.c2c(row, inverse) for row in x.reshape(-1, x.shape[-1])]).reshape(x.shape)

    # --- FFT<T>::transform_no_scramble<inv> (fft.h:69-73): in place on bit-reversed input ---
    def c2c_no_scramble(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        y = np.array(x, copy=True, order="C")
        f = self._fn("c2c_no_scramble", _real_dtype(y), argtypes=[_c.c_void_p, _c.c_int, _c.c_int])
        if f(_ptr(y), y.shape[-1], int(inverse)):
            raise ValueError("c2c_no_scramble rejected size")
        return y

    # --- RealFFT<T>::forward(out, in, half) (FFTReal.h:204-213) ---
    def r2c(self, x: np.ndarray, half: bool = True, fill=None) -> np.ndarray:
        """Returns the full n-element output buffer; with half=True only the first n/2+1 are written."""
        x = np.ascontiguousarray(x)
        n = x.shape[-1]
        out = np.full(max(n, 1), fill if fill is not None else 0, dtype=_CPX[x.dtype])
        f = self._fn("r2c", x.dtype, argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int])
        if f(_ptr(out), _ptr(x), n, int(half)):
            raise ValueError("r2c rejected size")
        return out

    # --- DIT<T>::apply(out, in, half) (fft.h:173-196) ---
    def dit(self, z: np.ndarray, n: int, half: bool = False, in_place: bool = False) -> np.ndarray:
        rd = _real_dtype(z)
        buf = np.zeros(max(n, 1), dtype=_CPX[rd])
        buf[: z.shape[0]] = z
        f = self._fn("dit", rd, argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int])
        if in_place:
            if f(_ptr(buf), _ptr(buf), n, int(half)):
                raise ValueError("dit rejected size")
            return buf
        out = np.zeros(max(n, 1), dtype=_CPX[rd])
        if f(_ptr(out), _ptr(buf), n, int(half)):
            raise ValueError("dit rejected size")
        return out

    # --- FFTVert<T>::transform<inv>(out, os, in, is, cols) (fft.h:145-150) ---
    def vert(self, x: np.ndarray, inverse: bool = False, cols: int | None = None) -> np.ndarray:
        """x: (n, stride) complex array; transforms the first `cols` columns along axis 0."""
        x = np.ascontiguousarray(x)
        n, stride = x.shape
        cols = stride if cols is None else cols
        out = np.zeros_like(x)
        f = self._fn("vert", _real_dtype(x),
                     argtypes=[_c.c_void_p, _c.c_long, _c.c_void_p, _c.c_long, _c.c_int, _c.c_int, _c.c_int])
        if f(_ptr(out), stride, _ptr(x), stride, n, cols, int(inverse)):
            raise ValueError("vert rejected size")
        return out

    # --- FFT2D<T>::transform<inv> (fft.h:213-218); x is (height, width) ---
    def fft2d(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x)
        h, w = x.shape
        out = np.empty_like(x)
        f = self._fn("fft2d", _real_dtype(x),
                     argtypes=[_c.c_void_p, _c.c_long, _c.c_void_p, _c.c_long, _c.c_int, _c.c_int, _c.c_int])
        if f(_ptr(out), w, _ptr(x), w, w, h, int(inverse)):
            raise ValueError("fft2d rejected size")
        return out

    # --- FFT<T>::transform_real (fft.h:90-94) ---
    def transform_real(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x)
        out = np.empty(x.shape[-1], dtype=_CPX[x.dtype])
        f = self._fn("transform_real", x.dtype, argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int])
        if f(_ptr(out), _ptr(x), x.shape[-1]):
            raise ValueError("transform_real rejected size")
        return out

    # --- FFT<T>::transform_interleave + separate_2x_real_FFT (fft.h:100-105, FFTReal.h:35-66) ---
    def two_real(self, a: np.ndarray, b: np.ndarray):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        n = a.shape[-1]
        o1 = np.empty(n, dtype=_CPX[a.dtype])
        o2 = np.empty(n, dtype=_CPX[a.dtype])
        f = self._fn("two_real", a.dtype, argtypes=[_c.c_void_p] * 4 + [_c.c_int])
        if f(_ptr(o1), _ptr(o2), _ptr(a), _ptr(b), n):
            raise ValueError("two_real rejected size")
        return o1, o2


    # --- RealFFT2D<T> (FFTReal.h:71-184) ---
    def real_fft2d(self, x: np.ndarray) -> np.ndarray:
        """RealFFT2D<T>::forward (FFTReal.h:83-104); x is (height, width) real."""
        x = np.ascontiguousarray(x)
        h, w = x.shape
        out = np.zeros((h, w), dtype=_CPX[x.dtype])
        f = self._fn("real_fft2d", x.dtype,
                     argtypes=[_c.c_void_p, _c.c_long, _c.c_void_p, _c.c_long, _c.c_int, _c.c_int])
        if f(_ptr(out), w, _ptr(x), w, w, h):
            raise ValueError("real_fft2d rejected size")
        return out

    def real_fft2d_2x(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """RealFFT2D<T>::forward_2x (FFTReal.h:114-118) with equal input strides; a, b are (height, width) real."""
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        h, w = a.shape
        out = np.zeros((h, w), dtype=_CPX[a.dtype])
        f = self._fn("real_fft2d_2x", a.dtype,
                     argtypes=[_c.c_void_p, _c.c_long, _c.c_void_p, _c.c_void_p, _c.c_long, _c.c_int, _c.c_int])
        if f(_ptr(out), w, _ptr(a), _ptr(b), w, w, h):
            raise ValueError("real_fft2d_2x rejected size")
        return out


class Port(_Lib):
    """The plain-C restatement (genfft_oracle.c)."""
    prefix = "oracle_"
    path = PORT_SO


class Ref(_Lib):
    """The real reference, dispatch back-end (best ISA the host CPU has)."""
    prefix = "genfft_ref_"
    path = REF_SO

    def describe(self) -> str:
        self.lib.genfft_ref_describe.restype = _c.c_char_p
        return self.lib.genfft_ref_describe().decode()

    def hardware_threads(self) -> int:
        return int(self.lib.genfft_ref_hardware_threads())

    def c2c(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        n = x.shape[-1]
        if n > (1 << 23):  # beyond the reference's size switch: factory hook (ref_native_big.cpp)
            x = np.ascontiguousarray(x)
            out = np.empty_like(x)
            f = self._fn("big_c2c", _real_dtype(x), argtypes=[_c.c_void_p, _c.c_void_p, _c.c_long, _c.c_int])
            if f(_ptr(out), _ptr(x), n, int(inverse)):
                raise ValueError(f"reference cannot run n={n}")
            return out
        return super().c2c(x, inverse)

    def c2c_rows(self, x: np.ndarray, inverse: bool = False, threads: int = 0) -> np.ndarray:
        """FFT<T>::transform on every row of a 2-D host array, one plan, `threads` workers (0 = all host threads)."""
        x = np.ascontiguousarray(x)
        count, n = x.shape
        out = np.empty_like(x)
        f = self._fn("c2c_rows", _real_dtype(x),
                     argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_long, _c.c_int, _c.c_int])
        if f(_ptr(out), _ptr(x), n, count, threads or self.hardware_threads(), int(inverse)):
            raise ValueError(f"c2c_rows rejected n={n}")
        return out

    def dummy_complex(self, n: int, dtype=np.float32, real: bool = False) -> np.ndarray:
        """DummyData(std::vector<std::complex<T>>&, real) -- test/test_util.h:36-47."""
        out = np.empty(n, dtype=_CPX[np.dtype(dtype)])
        f = self._fn("dummy_complex", dtype, restype=None, argtypes=[_c.c_void_p, _c.c_long, _c.c_int])
        f(_ptr(out), n, int(real))
        return out

    def dummy_real(self, n: int, dtype=np.float32) -> np.ndarray:
        """DummyData(std::vector<T>&) -- test/test_util.h:49-57."""
        out = np.empty(n, dtype=dtype)
        f = self._fn("dummy_real", dtype, restype=None, argtypes=[_c.c_void_p, _c.c_long])
        f(_ptr(out), n)
        return out

    def testref_fft_pow2(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        """reference_impl::FFT_pow2 -- test/fft_ref_impl.h:90-96 (the reference tests' comparand)."""
        x = np.ascontiguousarray(x)
        out = np.empty_like(x)
        f = self._fn("testref_fft_pow2", _real_dtype(x), argtypes=[_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int])
        f(_ptr(out), _ptr(x), x.shape[-1], int(inverse))
        return out

    def testref_dft(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        """reference_impl::DFT -- test/fft_ref_impl.h:120-135 (naive O(n^2), double accumulation)."""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(x)
        f = self.lib.genfft_ref_testref_dft_f64
        f.restype = _c.c_int
        f.argtypes = [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int]
        f(_ptr(out), _ptr(x), x.shape[-1], int(inverse))
        return out

    # ---- CPU-baseline timing (restated test/fft_bench.cpp loops; seconds inside transform calls) ----
    def bench_c2c(self, n: int, count: int, threads: int, dtype=np.float32, fwd_only: bool = True) -> float:
        if n > (1 << 23):
            f = self.lib.genfft_ref_bench_big_c2c_f64
            f.restype = _c.c_double
            f.argtypes = [_c.c_long, _c.c_long]
            return float(f(n, count))
        f = self._fn("bench_c2c", dtype, restype=_c.c_double,
                     argtypes=[_c.c_int, _c.c_long, _c.c_int, _c.c_int])
        return float(f(n, count, threads, int(fwd_only)))

    def bench_c2c_array(self, n: int, count: int, threads: int, warm: int = 1, reps: int = 1) -> float:
        """mean seconds per sweep over `count` distinct transforms, host array in -> host array out (DRAM-streaming)"""
        f = self.lib.genfft_ref_bench_c2c_array_f32
        f.restype = _c.c_double
        f.argtypes = [_c.c_int, _c.c_long, _c.c_int, _c.c_int, _c.c_int]
        return float(f(n, count, threads, warm, reps))

    def bench_r2c(self, n: int, count: int, threads: int) -> float:
        f = self.lib.genfft_ref_bench_r2c_f32
        f.restype = _c.c_double
        f.argtypes = [_c.c_int, _c.c_long, _c.c_int]
        return float(f(n, count, threads))

    def bench_fft2d(self, w: int, h: int, count: int) -> float:
        f = self.lib.genfft_ref_bench_fft2d_f32
        f.restype = _c.c_double
        f.argtypes = [_c.c_int, _c.c_int, _c.c_long]
        return float(f(w, h, count))


class RefGeneric(_Lib):
    """The real reference, generic scalar back-end (no SIMD, no FMA): the bit-exact pin for Port."""
    prefix = "genfft_ref_generic_"
    path = REF_SO


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_port() -> bool:
    return os.path.exists(PORT_SO)


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 in double -- the parity measure north_star states."""
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def tolerance(n_total: int, dtype) -> float:
    """north_star: rel-L2 <= 1e-6*log2(N) (float), 1e-14*log2(N) (double)."""
    lg = max(1.0, float(np.log2(max(2, n_total))))
    return (1e-6 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else 1e-14) * lg
