#!/bin/bash
# A/B: programmatic dependent launch of the tile kernels (GENFFT_CUDA_PDL).
set -u
mkdir -p gpurun_out
: > gpurun_out/pdl_ab.log
for v in 0 1 0 1; do GENFFT_CUDA_PDL=$v timeout 120 python tools/c1_latency.py >> gpurun_out/pdl_ab.log 2>&1; done
for v in 0 1; do echo "== PDL=$v" >> gpurun_out/pdl_ab.log; GENFFT_CUDA_PDL=$v timeout 200 python tools/variant_bench.py lib c3 c4 c5 c2 2>&1 | grep -v "^==" | cut -c1-90 >> gpurun_out/pdl_ab.log; done
cat gpurun_out/pdl_ab.log
(GENFFT_CUDA_PDL=1 timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_bench_contract.py) > gpurun_out/pytest_pdl.log 2>&1; tail -3 gpurun_out/pytest_pdl.log
