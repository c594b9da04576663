#!/bin/bash
# Round 2, last experiment call: what is left in the C2 contract kernel (0.96 of the measured HBM peak)?
#   tiles per CTA of the TMA mode (end-of-grid tail vs per-CTA prologue), streaming stores, evict-first TMA loads,
#   a register budget sized for the 3 CTAs/SM the shared memory allows (80 instead of 64 registers);
#   plus the 3-CTAs-per-SM / 80-register budget on the multi-pass f32 kernels (C4, C5).
set -u
mkdir -p gpurun_out
L=gpurun_out/c2_sweep.log
: > $L
timeout 150 python tools/c2_sweep.py lib 2,3,4,5,6,8,10,12,16 4 40 >> $L 2>&1
timeout 100 python tools/c2_sweep.py lib_exp_cs,lib_exp_ef,lib_exp_lb,lib 4,8 4 40 >> $L 2>&1
timeout 150 python tools/variant_bench.py lib,lib_exp_t768,lib,lib_exp_t768 c4 c5 2>&1 | cut -c1-110 >> $L
cat $L
