"""ctypes binding of libgenfft_cuda.so (the C ABI declared in include/genfft_cuda.h).

The library is built in-tree by ``genfft_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is
no fallback: if the shared object is missing this module raises, and every call fails loudly when no
sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GENFFT_CUDA_LIB selects another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("GENFFT_CUDA_LIB") or os.path.join(HERE, "lib", "libgenfft_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "genfft_cuda.h")

F32, F64 = 0, 1


class GenfftCudaError(RuntimeError):
    pass


def build(verbose: bool = False) -> None:
    import subprocess
    subprocess.run(["bash", os.path.join(HERE, "csrc", "build.sh")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None

_i64 = C.c_int64
_vp = C.c_void_p
_plan_p = C.POINTER(C.c_void_p)

_SIGNATURES = {
    "genfft_cuda_last_error_string": (C.c_char_p, []),
    "genfft_cuda_device_count": (C.c_int, []),
    "genfft_cuda_launch_count": (C.c_uint64, []),
    "genfft_cuda_plan_c2c_1d": (C.c_int, [_plan_p, C.c_int, _i64, _i64, _i64, _i64]),
    "genfft_cuda_plan_r2c_1d": (C.c_int, [_plan_p, C.c_int, _i64, _i64, C.c_int, _i64, _i64]),
    "genfft_cuda_plan_c2c_2d": (C.c_int, [_plan_p, C.c_int, _i64, _i64]),
    "genfft_cuda_plan_c2r_1d": (C.c_int, [_plan_p, C.c_int, _i64, _i64, _i64, _i64]),
    "genfft_cuda_exec_c2r_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "genfft_cuda_exec_c2r": (C.c_int, [_vp, _vp, _vp]),
    "genfft_cuda_plan_vert": (C.c_int, [_plan_p, C.c_int, _i64]),
    "genfft_cuda_plan_dit": (C.c_int, [_plan_p, C.c_int, _i64]),
    "genfft_cuda_plan_destroy": (C.c_int, [_vp]),
    "genfft_cuda_plan_set_grid_fraction": (C.c_int, [_vp, C.c_double, C.c_double]),
    "genfft_cuda_plan_size": (_i64, [_vp]),
    "genfft_cuda_plan_num_passes": (C.c_int, [_vp]),
    "genfft_cuda_plan_scratch_bytes": (C.c_size_t, [_vp]),
    "genfft_cuda_plan_describe": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "genfft_cuda_exec_c2c_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "genfft_cuda_exec_c2c_no_scramble_dev": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "genfft_cuda_exec_c2c_real_in_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "genfft_cuda_exec_r2c_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "genfft_cuda_exec_c2c_interleave_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "genfft_cuda_separate_2x_real_dev": (C.c_int, [C.c_int, _vp, _vp, _vp, _i64, _vp]),
    "genfft_cuda_separate_2x_real": (C.c_int, [C.c_int, _vp, _vp, _vp, _i64]),
    "genfft_cuda_plan_r2c_2d": (C.c_int, [_plan_p, C.c_int, _i64, _i64]),
    "genfft_cuda_exec_r2c_2d_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "genfft_cuda_exec_c2c_interleave": (C.c_int, [_vp, _vp, _vp, _vp]),
    "genfft_cuda_exec_r2c_2d_2x_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "genfft_cuda_exec_r2c_2d_2x": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64]),
    "genfft_cuda_exec_r2c_2d": (C.c_int, [_vp, _vp, _i64, _vp, _i64]),
    "genfft_cuda_exec_c2c_2d_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp]),
    "genfft_cuda_exec_vert_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp]),
    "genfft_cuda_exec_vert_no_scramble_dev": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "genfft_cuda_exec_dit_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "genfft_cuda_exec_c2c": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "genfft_cuda_exec_c2c_no_scramble": (C.c_int, [_vp, _vp, C.c_int]),
    "genfft_cuda_exec_c2c_real_in": (C.c_int, [_vp, _vp, _vp]),
    "genfft_cuda_exec_r2c": (C.c_int, [_vp, _vp, _vp]),
    "genfft_cuda_exec_c2c_2d": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int]),
    "genfft_cuda_exec_vert": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int]),
    "genfft_cuda_exec_vert_no_scramble": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int]),
    "genfft_cuda_exec_dit": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "genfft_cuda_plan_dist_rows": (C.c_int, [_plan_p, C.c_int, _i64, _i64, C.c_int]),
    "genfft_cuda_exec_dist_rows_dev": (C.c_int, [_vp, _vp, C.POINTER(_vp), _i64, _i64, _vp, _i64, C.c_int, _vp]),
    "genfft_cuda_plan_dist_cols": (C.c_int, [_plan_p, C.c_int, _i64, _i64, C.c_int]),
    "genfft_cuda_exec_dist_cols_dev": (C.c_int, [_vp, _vp, C.POINTER(_vp), _i64, _i64, _vp, _i64, C.c_int, _vp]),
    "genfft_cuda_exec_dist_cols_tw_dev": (C.c_int, [_vp, C.POINTER(_vp), _i64, _i64, _vp, _i64, C.c_int, _i64, C.POINTER(C.c_int), _vp]),
    "genfft_cuda_scatter_cols_dev": (C.c_int, [C.c_int, C.POINTER(_vp), C.c_int, _i64, _vp, _i64, _i64, _i64, _vp]),
    "genfft_cuda_copy2d_dev": (C.c_int, [C.c_int, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "genfft_cuda_twiddle2d_dev": (C.c_int, [C.c_int, _vp, _i64, _i64, _i64, _i64, _i64, C.c_int, _vp]),
    "genfft_cuda_transpose_dev": (C.c_int, [C.c_int, _vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "genfft_cuda_debug_fast_div": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "genfft_cuda_debug_mode_launch_count": (C.c_uint64, [C.c_int]),
    "genfft_cuda_debug_time_c2c_pairs": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.POINTER(C.c_double)]),
    "genfft_cuda_peer_barrier_dev": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint32, _vp]),
    "genfft_cuda_memset_dev": (C.c_int, [_vp, C.c_int, C.c_size_t]),
    "genfft_cuda_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t, C.c_int, C.POINTER(C.c_int)]),
    "genfft_cuda_host_free": (C.c_int, [_vp]),
    "genfft_cuda_malloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "genfft_cuda_free": (C.c_int, [_vp]),
    "genfft_cuda_ipc_get_handle": (C.c_int, [_vp, C.c_char_p]),
    "genfft_cuda_ipc_open_handle": (C.c_int, [C.POINTER(_vp), C.c_char_p]),
    "genfft_cuda_ipc_close_handle": (C.c_int, [_vp]),
}


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Loads libgenfft_cuda.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GenfftCudaError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(genfft_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        if hasattr(handle, "genfft_emu_fiber_switches"):
            # tests/emu builds the kernel sources for host fibers to test their logic without a GPU; it is not a backend
            raise GenfftCudaError(f"{LIB_PATH} is the test suite's kernel-logic emulator, not libgenfft_cuda "
                                  "(genfft_b200 has no CPU or PyTorch fallback)")
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().genfft_cuda_last_error_string()
        raise GenfftCudaError(f"libgenfft_cuda error {rc}: {msg.decode() if msg else '?'}")
