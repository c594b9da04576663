#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-110; }
WHICH="c3 c4 2d"
run GENFFT_CUDA_TMA_COLS=1
run GENFFT_CUDA_TMA_COLS=0
