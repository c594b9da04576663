#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-130; }
WHICH="c2"
run GENFFT_CUDA_TMA=0
run GENFFT_CUDA_TMA_TILES=2
run GENFFT_CUDA_TMA_TILES=4
run GENFFT_CUDA_TMA_TILES=8
run GENFFT_CUDA_TMA_TILES=16
