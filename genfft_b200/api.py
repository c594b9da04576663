"""Host-side mirror of genFFT's public classes on top of libgenfft_cuda.

Same names, argument meaning and error behaviour as the reference's C++ classes
(``/root/reference/include/genFFT/fft.h`` and ``FFTReal.h``): ``FFT``, ``FFTVert``, ``DIT``, ``FFT2D``,
``RealFFT``; ``transform(out, in, inv)`` with an unscaled inverse, ``out != in`` where the reference
requires it, strides in complex elements, ``size()`` and truthiness of a default-constructed object.
Buffers are either CUDA ``torch`` tensors (device-pointer path, asynchronous on the current stream) or
host buffers -- numpy arrays / CPU tensors -- (host-pointer path, the literal CPU-caller drop-in).
Torch is used for device memory and streams only; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, GenfftCudaError, check, lib

try:  # torch is plumbing (device memory, streams); host buffers work without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _precision(dtype) -> int:
    if torch is not None and isinstance(dtype, torch.dtype):
        if dtype in (torch.float32, torch.complex64):
            return F32
        if dtype in (torch.float64, torch.complex128):
            return F64
        raise TypeError(f"unsupported dtype {dtype}")
    d = np.dtype(dtype)
    if d in (np.dtype(np.float32), np.dtype(np.complex64)):
        return F32
    if d in (np.dtype(np.float64), np.dtype(np.complex128)):
        return F64
    raise TypeError(f"unsupported dtype {dtype}")


_TORCH_DTYPES = {} if torch is None else {torch.float32: (F32, 1), torch.complex64: (F32, 2),
                                          torch.float64: (F64, 1), torch.complex128: (F64, 2)}


class _Buf:
    """Resolved view of a user buffer: raw pointer, residency, scalar precision, length in scalars."""

    __slots__ = ("ptr", "cuda", "precision", "nscalars", "keep")

    def __init__(self, x, writable: bool = False):
        self.keep = x
        if torch is not None and isinstance(x, torch.Tensor):
            if not x.is_contiguous():
                raise ValueError("buffers must be contiguous")
            info = _TORCH_DTYPES.get(x.dtype)  # (precision, scalars per element); a dict lookup: this is the C1 hot path
            if info is None:
                raise TypeError(f"unsupported dtype {x.dtype}")
            self.ptr = x.data_ptr()
            self.cuda = x.is_cuda
            self.precision = info[0]
            self.nscalars = x.numel() * info[1]
        elif isinstance(x, np.ndarray):
            if not x.flags["C_CONTIGUOUS"]:
                raise ValueError("buffers must be C-contiguous")
            if writable and not x.flags["WRITEABLE"]:
                raise ValueError("output buffer is read-only")
            self.ptr = x.ctypes.data
            self.cuda = False
            self.precision = _precision(x.dtype)
            self.nscalars = x.size * (2 if np.iscomplexobj(x) else 1)
        else:
            raise TypeError("expected a torch.Tensor or numpy.ndarray")


_raw_stream = getattr(getattr(torch, "_C", None), "_cuda_getCurrentRawStream", None) if torch is not None else None


def _stream() -> int:
    """The current CUDA stream's handle (launches are asynchronous on it, like any torch op)."""
    if torch is None:
        return 0
    if _raw_stream is not None:  # ~0.3 us instead of ~1.5 us for building a torch.cuda.Stream object per call
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _pair(out, inp, precision):
    o, i = _Buf(out, True), _Buf(inp)
    if o.cuda != i.cuda:
        raise ValueError("out and in must both be device tensors or both be host buffers")
    if o.precision != precision or i.precision != precision:
        raise TypeError("buffer dtype does not match the plan's precision")
    return o, i


class _Plan:
    def __init__(self):
        self._h = C.c_void_p(None)
        self._n = 0

    def __del__(self):
        try:
            if self._h:
                lib().genfft_cuda_plan_destroy(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def __bool__(self):  # explicit operator bool() (fft.h:108)
        return bool(self._h)

    def size(self) -> int:  # fft.h:107
        return self._n

    @property
    def num_passes(self) -> int:
        return lib().genfft_cuda_plan_num_passes(self._h) if self._h else 0

    def describe(self) -> str:
        buf = C.create_string_buffer(1024)
        check(lib().genfft_cuda_plan_describe(self._h, buf, 1024))
        return buf.value.decode()

    def _need(self):
        if not self._h:
            raise GenfftCudaError("empty (default-constructed) plan")


class FFT(_Plan):
    """genfft::FFT<T> (fft.h:54-113) plus a batch dimension.

    ``FFT(n, dtype)`` is the reference's ``FFT<T>(n)``; ``batch`` transforms are laid out ``in_dist`` /
    ``out_dist`` complex elements apart (default ``n``)."""

    def __init__(self, n: int | None = None, dtype=np.float32, batch: int = 1, in_dist: int = 0, out_dist: int = 0):
        super().__init__()
        if n is None:
            return
        self.precision = _precision(dtype)
        self.batch = batch
        self.in_dist = in_dist or n
        self.out_dist = out_dist or n
        check(lib().genfft_cuda_plan_c2c_1d(C.byref(self._h), self.precision, n, batch, in_dist, out_dist))
        self._n = n

    def _check_len(self, b: _Buf, dist: int, complex_elems: bool = True):
        need = ((self.batch - 1) * dist + self._n) * (2 if complex_elems else 1)
        if b.nscalars < need:
            raise ValueError(f"buffer too small: {b.nscalars} scalars, need {need}")

    def transform(self, out, inp, inv: bool = False):
        """FFT<T>::transform<inv>(out, in) (fft.h:80-85).  Unscaled inverse."""
        self._need()
        o, i = _pair(out, inp, self.precision)
        self._check_len(o, self.out_dist)
        self._check_len(i, self.in_dist)
        if o.cuda:
            check(lib().genfft_cuda_exec_c2c_dev(self._h, o.ptr, i.ptr, int(inv), _stream()))
        else:
            check(lib().genfft_cuda_exec_c2c(self._h, o.ptr, i.ptr, int(inv)))
        return out

    def forward(self, out, inp):  # README.txt:30
        return self.transform(out, inp, False)

    def inverse(self, out, inp):  # README.txt:31 -- no scaling
        return self.transform(out, inp, True)

    def transform_no_scramble(self, inout, inv: bool = False):
        """FFT<T>::transform_no_scramble<inv>(inout) (fft.h:69-73): bit-reversed input, in place."""
        self._need()
        b = _Buf(inout, True)
        if b.precision != self.precision:
            raise TypeError("buffer dtype does not match the plan's precision")
        self._check_len(b, self.in_dist)
        if b.cuda:
            check(lib().genfft_cuda_exec_c2c_no_scramble_dev(self._h, b.ptr, int(inv), _stream()))
        else:
            check(lib().genfft_cuda_exec_c2c_no_scramble(self._h, b.ptr, int(inv)))
        return inout

    def transform_real(self, out, inp):
        """FFT<T>::transform_real(out, in) (fft.h:90-94): n real inputs, n complex bins."""
        self._need()
        o, i = _pair(out, inp, self.precision)
        self._check_len(o, self.out_dist)
        self._check_len(i, self.in_dist, complex_elems=False)
        if o.cuda:
            check(lib().genfft_cuda_exec_c2c_real_in_dev(self._h, o.ptr, i.ptr, _stream()))
        else:
            check(lib().genfft_cuda_exec_c2c_real_in(self._h, o.ptr, i.ptr))
        return out


    def transform_interleave(self, out, in1, in2):
        """FFT<T>::transform_interleave(out, in1, in2) (fft.h:100-105): transform of in1 + i*in2; separate the two
        real spectra with separate_2x_real_FFT."""
        self._need()
        o, i1 = _pair(out, in1, self.precision)
        i2 = _Buf(in2)
        if i2.cuda != o.cuda or i2.precision != self.precision:
            raise ValueError("in2 must match in1")
        self._check_len(o, self.out_dist)
        self._check_len(i1, self.in_dist, complex_elems=False)
        self._check_len(i2, self.in_dist, complex_elems=False)
        if o.cuda:
            check(lib().genfft_cuda_exec_c2c_interleave_dev(self._h, o.ptr, i1.ptr, i2.ptr, _stream()))
        else:
            check(lib().genfft_cuda_exec_c2c_interleave(self._h, o.ptr, i1.ptr, i2.ptr))
        return out


def separate_2x_real_FFT(out1, out2, inp, n: int):
    """genfft::separate_2x_real_FFT(out1, out2, in, N) (FFTReal.h:35-66); device tensors or host arrays (all three on
    the same side); out1/out2 may alias in."""
    o1, o2, i = _Buf(out1, True), _Buf(out2, True), _Buf(inp)
    if not (o1.cuda == o2.cuda == i.cuda):
        raise ValueError("separate_2x_real_FFT: the three buffers must all be device tensors or all host arrays")
    if not (o1.precision == o2.precision == i.precision):
        raise TypeError("separate_2x_real_FFT: buffers of different precision")
    if min(o1.nscalars, o2.nscalars, i.nscalars) < 2 * n:
        raise ValueError("buffers must hold n complex elements")
    if i.cuda:
        check(lib().genfft_cuda_separate_2x_real_dev(i.precision, o1.ptr, o2.ptr, i.ptr, n, _stream()))
    else:
        check(lib().genfft_cuda_separate_2x_real(i.precision, o1.ptr, o2.ptr, i.ptr, n))
    return out1, out2


class RealFFT(_Plan):
    """genfft::RealFFT<T> (FFTReal.h:186-221) plus a batch dimension; ``half`` is fixed per plan."""

    def __init__(self, n: int | None = None, dtype=np.float32, half: bool = True, batch: int = 1,
                 in_dist: int = 0, out_dist: int = 0):
        super().__init__()
        if n is None:
            return
        self.precision = _precision(dtype)
        self.half = bool(half)
        self.batch = batch
        self.in_dist = in_dist or n
        self.out_len = 1 if n == 1 else (n // 2 + 1 if half else n)
        self.out_dist = out_dist or self.out_len
        check(lib().genfft_cuda_plan_r2c_1d(C.byref(self._h), self.precision, n, batch, int(half), in_dist, out_dist))
        self._n = n

    def forward(self, out, inp, half: bool | None = None):
        """RealFFT<T>::forward(out, in, half) (FFTReal.h:204-213).  With half=True exactly n/2+1 bins are written."""
        self._need()
        if half is not None and bool(half) != self.half:
            raise ValueError("`half` is fixed at plan creation")
        o, i = _pair(out, inp, self.precision)
        if o.nscalars < 2 * ((self.batch - 1) * self.out_dist + self.out_len):
            raise ValueError("output buffer too small")
        if i.nscalars < (self.batch - 1) * self.in_dist + self._n:
            raise ValueError("input buffer too small")
        if o.cuda:
            check(lib().genfft_cuda_exec_r2c_dev(self._h, o.ptr, i.ptr, _stream()))
        else:
            check(lib().genfft_cuda_exec_r2c(self._h, o.ptr, i.ptr))
        return out


class InverseRealFFT(_Plan):
    """Half-spectrum inverse of RealFFT (an addition: the reference has none, README.txt:51-52).  ``inverse(out, inp)``
    maps n/2+1 bins to n real points, unscaled like every genFFT inverse: inverse(forward(x)) == n * x."""

    def __init__(self, n: int | None = None, dtype=np.float32, batch: int = 1, in_dist: int = 0, out_dist: int = 0):
        super().__init__()
        if n is None:
            return
        self.precision = _precision(dtype)
        self.batch = batch
        self.in_dist = in_dist or n // 2 + 1
        self.out_dist = out_dist or n
        check(lib().genfft_cuda_plan_c2r_1d(C.byref(self._h), self.precision, n, batch, in_dist, out_dist))
        self._n = n

    def inverse(self, out, inp):
        self._need()
        o, i = _pair(out, inp, self.precision)
        if i.nscalars < 2 * ((self.batch - 1) * self.in_dist + self._n // 2 + 1):
            raise ValueError("input buffer too small")
        if o.nscalars < (self.batch - 1) * self.out_dist + self._n:
            raise ValueError("output buffer too small")
        if o.cuda:
            check(lib().genfft_cuda_exec_c2r_dev(self._h, o.ptr, i.ptr, _stream()))
        else:
            check(lib().genfft_cuda_exec_c2r(self._h, o.ptr, i.ptr))
        return out


class FFTVert(_Plan):
    """genfft::FFTVert<T> (fft.h:115-171): n-point FFT along axis 0 of an (n x cols) array."""

    def __init__(self, n: int | None = None, dtype=np.float32):
        super().__init__()
        if n is None:
            return
        self.precision = _precision(dtype)
        check(lib().genfft_cuda_plan_vert(C.byref(self._h), self.precision, n))
        self._n = n

    def transform(self, out, inp, cols: int, out_stride: int | None = None, in_stride: int | None = None,
                  inv: bool = False):
        """FFTVert<T>::transform<inv>(out, out_stride, in, in_stride, cols) (fft.h:145-150); dense overload :158-163."""
        self._need()
        out_stride = cols if out_stride is None else out_stride
        in_stride = cols if in_stride is None else in_stride
        o, i = _pair(out, inp, self.precision)
        if self._n and cols:
            if o.nscalars < 2 * ((self._n - 1) * out_stride + cols) or i.nscalars < 2 * ((self._n - 1) * in_stride + cols):
                raise ValueError("buffer too small for n x cols with the given stride")
        if o.cuda:
            check(lib().genfft_cuda_exec_vert_dev(self._h, o.ptr, out_stride, i.ptr, in_stride, cols, int(inv), _stream()))
        else:
            check(lib().genfft_cuda_exec_vert(self._h, o.ptr, out_stride, i.ptr, in_stride, cols, int(inv)))
        return out

    def transform_no_scramble(self, data, stride: int, cols: int, inv: bool = False):
        """FFTVert<T>::transform_no_scramble<inv>(data, stride, cols) (fft.h:132-136)."""
        self._need()
        b = _Buf(data, True)
        if b.precision != self.precision:
            raise TypeError("buffer dtype does not match the plan's precision")
        if self._n and cols and b.nscalars < 2 * ((self._n - 1) * stride + cols):
            raise ValueError("buffer too small for n x cols with the given stride")
        if b.cuda:
            check(lib().genfft_cuda_exec_vert_no_scramble_dev(self._h, b.ptr, stride, cols, int(inv), _stream()))
        else:
            check(lib().genfft_cuda_exec_vert_no_scramble(self._h, b.ptr, stride, cols, int(inv)))
        return data


class DIT(_Plan):
    """genfft::DIT<T> (fft.h:173-196): the real-FFT split of size n."""

    def __init__(self, n: int | None = None, dtype=np.float32):
        super().__init__()
        if n is None:
            return
        self.precision = _precision(dtype)
        check(lib().genfft_cuda_plan_dit(C.byref(self._h), self.precision, n))
        self._n = n

    def apply(self, out, inp, half: bool):
        """DIT<T>::apply(out, in, half) (fft.h:181-189); out may alias in."""
        self._need()
        o, i = _pair(out, inp, self.precision)
        n = self._n
        if i.nscalars < 2 * max(1, n // 2) or o.nscalars < 2 * (1 if n <= 1 else (n // 2 + 1 if half else n)):
            raise ValueError("buffer too small: DIT::apply reads n/2 complex elements and writes n/2+1 (half) or n")
        if o.cuda:
            check(lib().genfft_cuda_exec_dit_dev(self._h, o.ptr, i.ptr, int(half), _stream()))
        else:
            check(lib().genfft_cuda_exec_dit(self._h, o.ptr, i.ptr, int(half)))
        return out


class FFT2D(_Plan):
    """genfft::FFT2D<T>(width, height) (fft.h:198-245).  Note the (width, height) argument order."""

    def __init__(self, width: int | None = None, height: int | None = None, dtype=np.float32):
        super().__init__()
        if width is None:
            return
        self.precision = _precision(dtype)
        check(lib().genfft_cuda_plan_c2c_2d(C.byref(self._h), self.precision, width, height))
        self._w, self._hgt = width, height
        self._n = width * height

    def cols(self) -> int:  # fft.h:221
        return self._w if self._h else 0

    def rows(self) -> int:  # fft.h:223
        return self._hgt if self._h else 0

    def transform(self, out, inp, out_stride: int | None = None, in_stride: int | None = None, inv: bool = False):
        """FFT2D<T>::transform<inv>(out, out_stride, in, in_stride) (fft.h:213-218); out != in."""
        self._need()
        out_stride = self._w if out_stride is None else out_stride
        in_stride = self._w if in_stride is None else in_stride
        o, i = _pair(out, inp, self.precision)
        if o.nscalars < 2 * ((self._hgt - 1) * out_stride + self._w) or i.nscalars < 2 * ((self._hgt - 1) * in_stride + self._w):
            raise ValueError("buffer too small for height x width with the given stride")
        if o.cuda:
            check(lib().genfft_cuda_exec_c2c_2d_dev(self._h, o.ptr, out_stride, i.ptr, in_stride, int(inv), _stream()))
        else:
            check(lib().genfft_cuda_exec_c2c_2d(self._h, o.ptr, out_stride, i.ptr, in_stride, int(inv)))
        return out


class RealFFT2D(_Plan):
    """genfft::RealFFT2D<T>(width, height) (FFTReal.h:71-184): forward transform of a real image to its full
    width x height complex spectrum (the reference's inverse is a TODO, FFTReal.h:127-130)."""

    def __init__(self, width: int | None = None, height: int | None = None, dtype=np.float32):
        super().__init__()
        if width is None:
            return
        self.precision = _precision(dtype)
        check(lib().genfft_cuda_plan_r2c_2d(C.byref(self._h), self.precision, width, height))
        self._w, self._hgt = width, height
        self._n = width * height

    def cols(self) -> int:
        return self._w if self._h else 0

    def rows(self) -> int:
        return self._hgt if self._h else 0

    def forward(self, out, inp, out_stride: int | None = None, in_stride: int | None = None):
        """RealFFT2D<T>::forward(out, out_stride, in, in_stride); out_stride in complex elements, in_stride in scalars."""
        self._need()
        out_stride = self._w if out_stride is None else out_stride
        in_stride = self._w if in_stride is None else in_stride
        o, i = _pair(out, inp, self.precision)
        if o.nscalars < 2 * ((self._hgt - 1) * out_stride + self._w) or i.nscalars < (self._hgt - 1) * in_stride + self._w:
            raise ValueError("buffer too small for height x width with the given stride")
        if o.cuda:
            check(lib().genfft_cuda_exec_r2c_2d_dev(self._h, o.ptr, out_stride, i.ptr, in_stride, _stream()))
        else:
            check(lib().genfft_cuda_exec_r2c_2d(self._h, o.ptr, out_stride, i.ptr, in_stride))
        return out

    def forward_2x(self, out, in1, in2, out_stride: int | None = None, in_stride1: int | None = None,
                   in_stride2: int | None = None):
        """RealFFT2D<T>::forward_2x(out, out_stride, in1, in_stride1, in2, in_stride2) (FFTReal.h:106-118): the
        spectrum of in1 + i*in2, two real images in one complex 2D transform."""
        self._need()
        out_stride = self._w if out_stride is None else out_stride
        in_stride1 = self._w if in_stride1 is None else in_stride1
        in_stride2 = self._w if in_stride2 is None else in_stride2
        o, i1 = _pair(out, in1, self.precision)
        i2 = _Buf(in2)
        if i2.cuda != o.cuda or i2.precision != self.precision:
            raise ValueError("in2 must match in1")
        if (o.nscalars < 2 * ((self._hgt - 1) * out_stride + self._w)
                or i1.nscalars < (self._hgt - 1) * in_stride1 + self._w
                or i2.nscalars < (self._hgt - 1) * in_stride2 + self._w):
            raise ValueError("buffer too small for height x width with the given stride")
        if o.cuda:
            check(lib().genfft_cuda_exec_r2c_2d_2x_dev(self._h, o.ptr, out_stride, i1.ptr, in_stride1, i2.ptr,
                                                       in_stride2, _stream()))
        else:
            check(lib().genfft_cuda_exec_r2c_2d_2x(self._h, o.ptr, out_stride, i1.ptr, in_stride1, i2.ptr, in_stride2))
        return out


def launch_count() -> int:
    return int(lib().genfft_cuda_launch_count())


def device_count() -> int:
    return int(lib().genfft_cuda_device_count())
