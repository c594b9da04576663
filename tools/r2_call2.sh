#!/bin/bash
# Round 2, GPU call 2 (1 GPU): full -m gpu suite, the queued compile-time A/Bs, the direct-twiddle-table A/B, C3 knob
# sweep, the DSMEM exchange microbenchmark, one contract bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=8 --maxfail=10) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/variant_bench.py lib,lib_exp_packed,lib_exp_packed_tw,lib_exp_packed2 c2 c3f c4 c5 > gpurun_out/variant_packed.log 2>&1; cut -c1-110 gpurun_out/variant_packed.log
for lg in 0 21; do
  echo "== GENFFT_CUDA_DIRECT_TW_LOG2=$lg" >> gpurun_out/variant_directtw.log
  GENFFT_CUDA_DIRECT_TW_LOG2=$lg python tools/variant_bench.py lib,lib_exp_packed c3 c3f c4 c5 >> gpurun_out/variant_directtw.log 2>&1
done
python tools/variant_bench.py lib c3 >> gpurun_out/variant_directtw.log 2>&1
cut -c1-110 gpurun_out/variant_directtw.log
python tools/variant_bench.py lib,lib_exp_c2r c2r > gpurun_out/variant_c2r.log 2>&1; cut -c1-110 gpurun_out/variant_c2r.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_c2r/libgenfft_cuda.so python -m pytest tests/test_gpu_real_vert_2d.py tests/test_gpu_random_sweep.py -m gpu -q -x -k 'half_spectrum or r2c_c2r' > gpurun_out/pytest_c2r.log 2>&1; tail -2 gpurun_out/pytest_c2r.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_packed/libgenfft_cuda.so python -m pytest tests/test_gpu_c2c.py tests/test_gpu_real_vert_2d.py tests/test_gpu_chain.py -m gpu -q -x > gpurun_out/pytest_packed.log 2>&1; tail -2 gpurun_out/pytest_packed.log
for env in "GENFFT_CUDA_CHAIN12=1" "GENFFT_CUDA_CHAIN12=0"; do
  echo "== packed $env" >> gpurun_out/chain12.log
  env $env GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_packed/libgenfft_cuda.so timeout 300 python tools/quick_bench.py c4 >> gpurun_out/chain12.log 2>&1
done
cut -c1-110 gpurun_out/chain12.log
# C3 (fp64 2^24) knob sweep
for env in "X=1" "GENFFT_CUDA_P_F64=8" "GENFFT_CUDA_P_F64=8 GENFFT_CUDA_CHAIN=0" "GENFFT_CUDA_CHAIN=0" "GENFFT_CUDA_MAXLEN_F64=4096" "GENFFT_CUDA_MAXLEN_F64=1024" "GENFFT_CUDA_WIDE_C_F64=16" "GENFFT_CUDA_CHAIN_KB=2048" "GENFFT_CUDA_CHAIN_KB=8192"; do
  echo "== $env" >> gpurun_out/c3_sweep.log
  env $env timeout 120 python tools/variant_bench.py lib c3 2>&1 | grep -v "^==" | cut -c1-200 >> gpurun_out/c3_sweep.log
done
cat gpurun_out/c3_sweep.log | cut -c1-130
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dsmem_bench.cu -o /tmp/dsmem_bench && timeout 60 /tmp/dsmem_bench > gpurun_out/dsmem_bench.log 2>&1; cat gpurun_out/dsmem_bench.log
(time python bench.py) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json; echo
