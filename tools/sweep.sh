#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-250; }
WHICH="2dmid"
run A=0
run GENFFT_CUDA_WIDE_SINGLE_F32=512
run GENFFT_CUDA_WIDE_SINGLE_F32=256
WHICH="c3 c4 2d"
run A=0
