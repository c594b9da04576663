#!/bin/bash
# ncu with source correlation for ONE launch of C5's row chain kernel (per-instruction stall samples).
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fft_chain -s 2 -c 1 -f -o gpurun_out/r02_c5_rows_chain python tools/passes.py c5 > gpurun_out/ncu_c5_chain.log 2>&1
ls -la gpurun_out/r02_c5_rows_chain.ncu-rep
ncu -i gpurun_out/r02_c5_rows_chain.ncu-rep --page source --csv > gpurun_out/r02_c5_rows_chain_source.csv 2>/dev/null; wc -l gpurun_out/r02_c5_rows_chain_source.csv
