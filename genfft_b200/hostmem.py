"""Host buffers for the host-pointer path (genfft_cuda_exec_*): page-locked memory on the NUMA node of the GPU.

The host-pointer entry points move every byte across PCIe twice (H2D, D2H), so for N = 4096 the transform is ~2 % of
the call and the placement of the CALLER's buffers decides the rate: on a two-socket 8-GPU host a pinned buffer that
lives on the other socket crosses the socket interconnect as well, and with one process per GPU all eight links are
busy at once.  ``PinnedNearGpu`` wraps ``genfft_cuda_host_alloc`` (mmap + mbind to the current device's node +
cudaHostRegister), which needs no CPU of that node in the process's cpuset.  Nothing in the reference corresponds to
this (genFFT never leaves the CPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


class PinnedNearGpu:
    """A page-locked host array placed on the NUMA node of the CURRENT CUDA device (when ``numa_local``)."""

    def __init__(self, shape, dtype, numa_local: bool = True):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p, node = C.c_void_p(), C.c_int(-1)
        check(lib().genfft_cuda_host_alloc(C.byref(p), self.nbytes, int(numa_local), C.byref(node)))
        self.ptr, self.numa_node = p.value, node.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def tensor(self):
        import torch
        return torch.from_numpy(self.array)

    def close(self):
        if self.ptr:
            self.array = None
            lib().genfft_cuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
