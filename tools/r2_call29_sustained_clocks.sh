#!/bin/bash
# Round 2: clocks, power and throttle reasons while the C2 kernel runs back to back for seconds (the sustained regime of
# profiles/r02_ab_c2_tiles_cache_operators.log), sampled every 100 ms.
set -u
mkdir -p gpurun_out
L=gpurun_out/sustained_clocks.log
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.sw_thermal_slowdown --format=csv -lms 100 > gpurun_out/sustained_smi.csv 2>&1 &
SMI=$!
timeout 120 python tools/c2_sweep.py lib 8 12 400 > $L 2>&1
sleep 1
kill $SMI
cat $L
# idle head, then the loaded samples
head -3 gpurun_out/sustained_smi.csv
awk -F', ' 'NR>1 {gsub(/ W/,"",$4); if ($4+0 > 400) print}' gpurun_out/sustained_smi.csv | head -60
