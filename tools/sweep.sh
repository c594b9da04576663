#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-300; }
WHICH="c4 2d"
run A=0
run GENFFT_CUDA_MAXLEN_F32=256
run GENFFT_CUDA_WIDE_C_F32=16
