"""genfft_b200 -- B200-native FFT engine behind genFFT's API.

The package is a thin host layer over ``lib/libgenfft_cuda.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/genfft_cuda.h``).  It has no CPU or PyTorch compute path.
"""
from ._lib import F32, F64, GenfftCudaError, LIB_PATH, build, exported_symbols, lib
from .api import DIT, FFT, FFT2D, FFTVert, InverseRealFFT, RealFFT, RealFFT2D, device_count, launch_count, separate_2x_real_FFT

__all__ = ["FFT", "FFTVert", "DIT", "FFT2D", "RealFFT", "RealFFT2D", "InverseRealFFT", "separate_2x_real_FFT", "F32", "F64", "GenfftCudaError", "LIB_PATH", "build",
           "exported_symbols", "lib", "device_count", "launch_count"]
