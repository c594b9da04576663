#!/bin/bash
# A/B: chain kernel with the counting phase barrier + per-warp release/acquire (lib) vs two CTA-wide barriers per tile
# (lib_exp_old); then the chain and parity tests on the new library.
set -u
mkdir -p gpurun_out
timeout 300 python tools/variant_bench.py lib_exp_old,lib,lib_exp_old,lib c3 c3f c5 > gpurun_out/variant_chainsync.log 2>&1; cut -c1-100 gpurun_out/variant_chainsync.log
(time timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_c2c.py tests/test_gpu_real_vert_2d.py tests/test_gpu_c5_full.py tests/test_gpu_zz_ranks_in_process.py tests/test_gpu_ipc_same_device.py -m gpu -q -x) > gpurun_out/pytest_chainsync.log 2>&1; tail -3 gpurun_out/pytest_chainsync.log
