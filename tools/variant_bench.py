"""Scratch: A/B timing of compile-time variants of libgenfft_cuda in one process.
usage: python tools/variant_bench.py <lib_dir>[,<lib_dir>...] <config>...   (lib_dir relative to genfft_b200/, e.g. lib,lib_exp_a)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genfft_b200 as g
from genfft_b200 import _lib
sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_bench as q

libs = sys.argv[1].split(",")
which = sys.argv[2:]
for name in libs:
    _lib._lib = None
    _lib.LIB_PATH = os.path.join(ROOT, "genfft_b200", name, "libgenfft_cuda.so")
    if not os.path.exists(_lib.LIB_PATH):
        print(f"== {name}: not built (GENFFT_LIB_OUT=genfft_b200/<dir> GENFFT_NVCC_EXTRA=... bash genfft_b200/csrc/build.sh)", flush=True)
        continue
    print(f"== {name}", flush=True)
    if "c3" in which:
        q.c2c(1 << 24, 1, np.float64)
        q.c2c(1 << 20, 16, np.float64)
        q.c2c(4096, 1 << 15, np.float64)
    if "c3f" in which:
        q.c2c(1 << 24, 1, np.float32)
        q.c2c(1 << 21, 256, np.float32)
    if "c4" in which:
        q.r2c(1 << 22, 256)
    if "c5" in which:
        q.fft2d(32768, 32768)
    if "c2r" in which:
        q.c2r(4096, 1 << 16)
        q.c2r(1 << 22, 64)
    if "small2" in which:  # single-pass sizes, and two batches that fit in L2 (evict-first stores must not hurt them)
        for n in (256, 1024, 16384):
            q.c2c(n, (1 << 28) // n)
        q.c2c(4096, 1024)
        q.c2c(4096, 256)
        q.r2c(4096, 1 << 16)
    if "c2" in which:
        q.c2c(4096, 1 << 16)
    torch.cuda.empty_cache()
