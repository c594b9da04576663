#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-200; }
WHICH="c3 c4 2d"
run A=0
run GENFFT_CUDA_WIDE_C_F32=16
run GENFFT_CUDA_WIDE_C_F64=8
