#!/bin/bash
# L2-resident pass chains: off vs on, group size and lag sweeps (tools/quick_bench.py configs)
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-110; }
WHICH="${WHICH:-c3 2d}"
run GENFFT_CUDA_CHAIN=0
run GENFFT_CUDA_CHAIN=1
for kb in 1024 2048; do run GENFFT_CUDA_CHAIN=1 GENFFT_CUDA_CHAIN_KB=$kb; done
for lag in 4 8 16; do run GENFFT_CUDA_CHAIN=1 GENFFT_CUDA_CHAIN_LAG=$lag; done
run GENFFT_CUDA_CHAIN=1 GENFFT_CUDA_CHAIN_KB=2048 GENFFT_CUDA_CHAIN_LAG=16
