#!/bin/bash
# First GPU call of a round (1 GPU): state of the tree on real hardware in one go.
#   bash tools/build_variants.sh                      # HERE first (no GPU, ~4 min): the variant libraries waiting for an A/B
#   gpurun --timeout 1500 -- bash tools/round_start.sh
# Writes everything under gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
(time python -m pytest tests -m gpu -q --durations=10 --maxfail=10) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
(time python bench.py) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.json; echo
python tools/quick_bench.py c3 c4 2d > gpurun_out/quick_bench.log 2>&1; cut -c1-100 gpurun_out/quick_bench.log
# compile-time variants waiting for an A/B (built beforehand by tools/build_variants.sh; skipped when absent)
if [ -f genfft_b200/lib_exp_packed/libgenfft_cuda.so ]; then
  python tools/variant_bench.py lib,lib_exp_packed,lib_exp_packed_tw,lib_exp_packed2 c2 c3f c4 c5 > gpurun_out/variant_packed.log 2>&1; cut -c1-120 gpurun_out/variant_packed.log
  python tools/variant_bench.py lib,lib_exp_c2r c2r > gpurun_out/variant_c2r.log 2>&1; cut -c1-120 gpurun_out/variant_c2r.log
  GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_c2r/libgenfft_cuda.so python -m pytest tests/test_gpu_real_vert_2d.py tests/test_gpu_random_sweep.py -m gpu -q -x -k 'half_spectrum or r2c_c2r' > gpurun_out/pytest_c2r.log 2>&1; tail -2 gpurun_out/pytest_c2r.log
  GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_packed/libgenfft_cuda.so python -m pytest tests/test_gpu_c2c.py tests/test_gpu_real_vert_2d.py -m gpu -q -x > gpurun_out/pytest_packed.log 2>&1; tail -2 gpurun_out/pytest_packed.log
fi
# paths written without a GPU at hand (off by default), in a process of their own: a trap there must not poison the rest
GENFFT_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q -k first_two > gpurun_out/pytest_experimental.log 2>&1; tail -2 gpurun_out/pytest_experimental.log
GENFFT_CUDA_CHAIN12=1 timeout 300 python tools/quick_bench.py c4 > gpurun_out/quick_bench_chain12.log 2>&1; cut -c1-120 gpurun_out/quick_bench_chain12.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_packed/libgenfft_cuda.so GENFFT_CUDA_CHAIN12=1 timeout 300 python tools/quick_bench.py c4 > gpurun_out/quick_bench_chain12_packed.log 2>&1; cut -c1-120 gpurun_out/quick_bench_chain12_packed.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_packed/libgenfft_cuda.so GENFFT_CUDA_CHAIN12=2 timeout 300 python tools/quick_bench.py c3 > gpurun_out/quick_bench_chain12_2_packed.log 2>&1; cut -c1-120 gpurun_out/quick_bench_chain12_2_packed.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
