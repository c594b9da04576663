#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over one small launch of every kernel mode, incl. round 2's peer-store modes,
# fused four-step twiddle and scatter.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py dist > gpurun_out/sanitizer_${tool}_dist.log 2>&1; tail -4 gpurun_out/sanitizer_${tool}_dist.log
done
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_cases.py > gpurun_out/sanitizer_memcheck_all.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_all.log
