#!/bin/bash
# Builds the compile-time variants of libgenfft_cuda that are waiting for an A/B measurement into
# genfft_b200/lib_exp_<name>/ (git-ignored, travels to the GPU box).  Run HERE (no GPU needed), then on the box:
#   python tools/variant_bench.py lib,lib_exp_packed c2 c3f c4 c5
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/.."
# packed single-precision adds (FADD2 on register pairs, radix.cuh): 14-20 % fewer instructions per tile, bit-identical
GENFFT_LIB_OUT="$ROOT/genfft_b200/lib_exp_packed" GENFFT_NVCC_EXTRA="-DGENFFT_PACKED_F32=1" bash "$ROOT/genfft_b200/csrc/build.sh"
# ... and with the two-instruction packed complex multiply (more shuffles, spills in the twiddled column passes)
GENFFT_LIB_OUT="$ROOT/genfft_b200/lib_exp_packed2" GENFFT_NVCC_EXTRA="-DGENFFT_PACKED_F32=2" bash "$ROOT/genfft_b200/csrc/build.sh"
# packed adds + tile-major inter-pass twiddle table (immediate offsets instead of 15 computed addresses per thread)
GENFFT_LIB_OUT="$ROOT/genfft_b200/lib_exp_packed_tw" GENFFT_NVCC_EXTRA="-DGENFFT_PACKED_F32=1 -DGENFFT_TWB_TILED=1" bash "$ROOT/genfft_b200/csrc/build.sh"
# half-spectrum inverse with its pre-process fused into the first pass's load (one pass over HBM less)
GENFFT_LIB_OUT="$ROOT/genfft_b200/lib_exp_c2r" GENFFT_NVCC_EXTRA="-DGENFFT_FUSED_C2R=1" bash "$ROOT/genfft_b200/csrc/build.sh"
