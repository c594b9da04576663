#!/bin/bash
# Round 2: C4's last two passes (twiddled pass + fused-split pass) as one L2 chain over WHOLE transforms (16 MiB groups,
# above the default GENFFT_CUDA_CHAIN_MAX_KB of 8192): does the chain pay with the round-2 scheduler?
set -u
mkdir -p gpurun_out
L=gpurun_out/c4_pair_chain.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 60 python tools/quick_bench.py c4 2>&1 | cut -c1-230 >> $L; }
run GENFFT_CUDA_CHAIN=1
run GENFFT_CUDA_CHAIN_MAX_KB=16384
run GENFFT_CUDA_CHAIN_MAX_KB=16384 GENFFT_CUDA_CHAIN_LAG=2
run GENFFT_CUDA_CHAIN_MAX_KB=16384 GENFFT_CUDA_CHAIN_LAG=3
run GENFFT_CUDA_CHAIN_MAX_KB=16384 GENFFT_CUDA_CHAIN_LAG=4
cat $L
