#!/bin/bash
# Round 2: cache operators of the stand-alone kernels' global accesses (call 22 found streaming stores 4.5 % faster on
# C2 under sustained load).  Variants: cs = TMA-mode stores .cs, wt = .wt, cs2 = every stand-alone store .cs,
# cs3 = cs2 + stand-alone loads .cs.  C2 in the bench's burst regime (one process per measurement) and sustained;
# C4 / C3 / small sizes / an L2-resident batch on the sustained harness.
set -u
mkdir -p gpurun_out
L=gpurun_out/cache_ops.log
: > $L
timeout 120 python tools/burst_c2.py lib,lib_exp_cs,lib_exp_wt,lib_exp_cs2,lib_exp_cs3 3 >> $L 2>&1
timeout 100 python tools/c2_sweep.py lib,lib_exp_cs,lib_exp_wt,lib_exp_cs2,lib_exp_cs3,lib 8 4 40 >> $L 2>&1
timeout 200 python tools/variant_bench.py lib,lib_exp_cs2,lib_exp_cs3,lib c4 c3 small2 2>&1 | cut -c1-120 >> $L
cat $L
