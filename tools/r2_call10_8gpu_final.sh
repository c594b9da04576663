#!/bin/bash
# Round 2, last 8-GPU call: the all-to-all SM-store ceiling (2/4/8 GPUs) and the contract bench at 8 GPUs with the final tree.
set -u
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/alltoall_store_bench.cu -o /tmp/a2a_bench && timeout 120 /tmp/a2a_bench > gpurun_out/alltoall_store_bench.log 2>&1; cat gpurun_out/alltoall_store_bench.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 30411 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/bench_8gpu_final.json 2> gpurun_out/bench_8gpu_final.err; tail -c 1700 gpurun_out/bench_8gpu_final.json; echo; tail -3 gpurun_out/bench_8gpu_final.err
