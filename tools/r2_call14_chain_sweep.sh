#!/bin/bash
# Chain group size / lag once more for the single-GPU configs, after the packed adds and the twiddle powers.
set -u
mkdir -p gpurun_out
: > gpurun_out/chain_sweep_r02.log
for kb in 2048 4096 8192; do
  for lag in 0 3 6 10; do
    echo "== CHAIN_KB=$kb CHAIN_LAG=$lag" >> gpurun_out/chain_sweep_r02.log
    GENFFT_CUDA_CHAIN_KB=$kb GENFFT_CUDA_CHAIN_LAG=$lag timeout 200 python tools/variant_bench.py lib c3 c5 2>&1 | grep -E "n=16777216|32768x32768" | cut -c1-75 >> gpurun_out/chain_sweep_r02.log
  done
done
cat gpurun_out/chain_sweep_r02.log
