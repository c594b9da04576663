#!/bin/bash
# Round 2: A/B of the running-pointer addressing variant, then the whole -m gpu suite and the bench on the main library.
set -u
mkdir -p gpurun_out
python tools/variant_bench.py lib,lib_exp_runptr,lib,lib_exp_runptr c3 c3f c4 c5 c2 > gpurun_out/variant_runptr.log 2>&1; cut -c1-100 gpurun_out/variant_runptr.log
(time python -m pytest tests -m gpu -q --durations=4 --maxfail=10) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
