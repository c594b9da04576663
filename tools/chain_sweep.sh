#!/bin/bash
# L2-resident pass chains: off vs on, group size and lag sweeps (tools/quick_bench.py configs)
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-110; }
WHICH="${WHICH:-c3 c4 2d}"
run GENFFT_CUDA_CHAIN=0
run GENFFT_CUDA_CHAIN=1
for kb in 2048 8192; do run GENFFT_CUDA_CHAIN=1 GENFFT_CUDA_CHAIN_KB=$kb; done
for lag in 2 3 6; do run GENFFT_CUDA_CHAIN=1 GENFFT_CUDA_CHAIN_LAG=$lag; done
