"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

Runs the test-suite's "device" code on the kernel-logic emulator (tests/emu/_build/libgenfft_emu.so: the real
genfft_b200/csrc sources compiled by g++ against a fake CUDA runtime, see shim/cuda_runtime.h), so that the kernels'
addressing / twiddle / barrier / plan logic is checked against the oracle on a machine without a GPU.

``install()`` is called by tests/conftest.py only when GENFFT_TEST_BACKEND=emu (tests/test_emu_suite.py starts such
a pytest run in a subprocess).  It
  * swaps the ctypes handle inside genfft_b200._lib for the emulator's (the package itself refuses to load it),
  * makes "cuda" mean host memory for torch: ``.cuda()`` is a host copy, ``device="cuda"`` becomes "cpu",
    ``torch.cuda.synchronize`` is a no-op -- the emulator's device pointers ARE host pointers,
  * marks CPU tensors as device-resident for genfft_b200.api so that the ``*_dev`` entry points (the measured
    path) are the ones exercised; numpy arrays still take the host-pointer entry points (host_exec.cu).
Nothing here is importable from genfft_b200, and -m gpu runs on a GPU box never set the variable.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# GENFFT_EMU_VARIANT=<name>:<extra compiler flags> selects a compile-time variant of the kernel sources, built into
# _build_<name>/ (e.g. "packed:-DGENFFT_PACKED_F32=1": the formulas of the packed-pair variant with scalar arithmetic)
_VARIANT = os.environ.get("GENFFT_EMU_VARIANT", "")
_VNAME, _, _VFLAGS = _VARIANT.partition(":")
BUILD_DIR = os.path.join(HERE, "_build_" + _VNAME if _VNAME else "_build")
LIB = os.path.join(BUILD_DIR, "libgenfft_emu.so")


def sources() -> list[str]:
    return (glob.glob(os.path.join(ROOT, "genfft_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(ROOT, "genfft_b200", "csrc", "*.h"))
            + glob.glob(os.path.join(HERE, "shim", "*.h")) + [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "build.sh"),
                                                               os.path.join(ROOT, "include", "genfft_cuda.h")])


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False) -> None:
    if force or stale():
        env = dict(os.environ, GENFFT_EMU_OUT=BUILD_DIR, GENFFT_EMU_EXTRA=_VFLAGS)
        subprocess.run(["bash", os.path.join(HERE, "build.sh")], check=True, stdout=subprocess.DEVNULL, env=env)


def load() -> C.CDLL:
    build()
    from genfft_b200 import _lib
    handle = C.CDLL(LIB)
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    handle.genfft_emu_fiber_switches.restype = C.c_ulonglong
    return handle


_installed = False


def install() -> C.CDLL:
    """Process-wide switch of the test process to the emulator (see the module docstring)."""
    global _installed
    import torch
    from torch.overrides import TorchFunctionMode

    from genfft_b200 import _lib, api
    handle = load()
    if _installed:
        return handle
    _lib._lib = handle

    def to_cpu(dev):
        return "cpu" if dev is not None and str(dev).startswith("cuda") else dev

    class CudaIsHost(TorchFunctionMode):
        def __torch_function__(self, func, types, args=(), kwargs=None):
            kwargs = dict(kwargs or {})
            if "device" in kwargs:
                kwargs["device"] = to_cpu(kwargs["device"])
            name = getattr(func, "__name__", "")
            if name == "cuda" and args and isinstance(args[0], torch.Tensor):
                return args[0].clone()  # a device copy never aliases its host source
            if name == "to" and len(args) >= 2 and isinstance(args[1], (str, torch.device)):
                args = (args[0], to_cpu(args[1])) + tuple(args[2:])
            return func(*args, **kwargs)

    mode = CudaIsHost()
    mode.__enter__()  # for the life of the test process
    install._mode = mode

    real_generator = torch.Generator
    torch.Generator = lambda device="cpu": real_generator(device=to_cpu(device))

    class _Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1

    orig_init = api._Buf.__init__

    def buf_init(self, x, writable=False):
        orig_init(self, x, writable)
        if isinstance(x, torch.Tensor):
            self.cuda = True  # emulated device memory is host memory

    api._Buf.__init__ = buf_init
    api._stream = lambda: 0
    _installed = True
    return handle
