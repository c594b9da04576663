#!/bin/bash
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-110; }
WHICH="c3 c4 2d"
run GENFFT_CUDA_L2_GROUP_MB=0
run GENFFT_CUDA_L2_GROUP_MB=16
run GENFFT_CUDA_L2_GROUP_MB=32
run GENFFT_CUDA_L2_GROUP_MB=48
run GENFFT_CUDA_L2_GROUP_MB=64
