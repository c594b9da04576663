#!/bin/bash
# Round 2, profiling call (1 GPU): ncu launch list of the contract bench, full capture of the dominant (C2) kernel, full
# capture of the multi-pass kernels of C3 / C4 / C5.  Numbers printed by runs under ncu are never bench values.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_tile -s 4 -c 1 -f -o gpurun_out/r02_c2 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-extras > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none -k regex:"fft_tile|fft_chain" -c 12 -f -o /tmp/r02_passes python tools/passes.py c3 c4 c5 > gpurun_out/ncu_passes.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_c2.ncu-rep > gpurun_out/r02_c2_kernel_ncu.txt 2>&1
python tools/ncu_summary.py /tmp/r02_passes.ncu-rep > gpurun_out/r02_pass_kernels_ncu.txt 2>&1
ncu -i /tmp/r02_passes.ncu-rep --page raw --csv > gpurun_out/r02_pass_kernels_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_c2.ncu-rep --page raw --csv > gpurun_out/r02_c2_kernel_ncu_full.csv 2>/dev/null
ls -la gpurun_out/ /tmp/*.ncu-rep; du -sh gpurun_out; cat gpurun_out/r02_pass_kernels_ncu.txt | grep -E "=====|time_duration|dram__bytes|issue_active|fp64|warps_active|stalls" | cut -c1-170
(time python bench.py) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 400 gpurun_out/bench_1gpu.json
# C4 / C5 plan-shape knobs once more with the packed arithmetic (two big passes instead of three small ones)
for env in "X=1" "GENFFT_CUDA_MAXLEN_F32=2048" "GENFFT_CUDA_MAXLEN_F32=1024" "GENFFT_CUDA_MAXLEN_F32=256" "GENFFT_CUDA_WIDE_C_F32=16" "GENFFT_CUDA_WIDE_C_F32=32"; do
  echo "== $env" >> gpurun_out/c4_c5_shape_sweep.log
  env $env timeout 200 python tools/variant_bench.py lib c4 c5 2>&1 | grep -v "^==" | cut -c1-220 >> gpurun_out/c4_c5_shape_sweep.log
done
cut -c1-150 gpurun_out/c4_c5_shape_sweep.log
