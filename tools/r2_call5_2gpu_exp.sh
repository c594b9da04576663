#!/bin/bash
# Round 2, 2-GPU experiment call: what limits the link-bound peer-storing pass?  (a) store shape microbenchmark,
# (b) knob sweep of the slab 2D transform (natural order only).
set -u
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/peer_store_bench.cu -o /tmp/peer_store_bench && timeout 120 /tmp/peer_store_bench > gpurun_out/peer_store_bench.log 2>&1; cat gpurun_out/peer_store_bench.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29700
: > gpurun_out/c5_dist_2gpu_sweep.jsonl
for env in "X=1" "GENFFT_CUDA_PEER_MODES=0" "GENFFT_CUDA_CHAIN_GRID_PCT=75" "GENFFT_CUDA_CHAIN_GRID_PCT=50" "GENFFT_CUDA_CHAIN_GRID_PCT=50 GENFFT_CUDA_PEER_MODES=0" "GENFFT_CUDA_CHAIN_LAG=4" "GENFFT_CUDA_CHAIN_LAG=8" "GENFFT_CUDA_CHAIN_KB=8192" "GENFFT_CUDA_CHAIN=0" "GENFFT_CUDA_CHAIN=0 GENFFT_CUDA_PEER_MODES=0"; do
  port=$((port+1))
  env $env $TR --master-port $port bench_dist.py --phases --transports p2p --chunks 1 --steps 10 --outputs natural >> gpurun_out/c5_dist_2gpu_sweep.jsonl 2>> gpurun_out/c5_dist_2gpu_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/c5_dist_2gpu_sweep.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); k = {a: b for a, b in d['knobs'].items()}
        print(round(d['ms'], 3), {a: b[1] for a, b in d['phases_ms_rank0_and_max'].items() if 'transpose' in a}, k)
PY
