#!/bin/bash
# Round 2, 2-GPU call: finer sweep of the chain's lag / grid share for the link-bound peer-storing passes.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29800
: > gpurun_out/c5_dist_2gpu_sweep2.jsonl
for env in "GENFFT_CUDA_CHAIN_LAG=8" "GENFFT_CUDA_CHAIN_LAG=8 GENFFT_CUDA_PEER_MODES=0" "GENFFT_CUDA_CHAIN_LAG=6" "GENFFT_CUDA_CHAIN_LAG=12" "GENFFT_CUDA_CHAIN_LAG=16" "GENFFT_CUDA_CHAIN_LAG=24" \
  "GENFFT_CUDA_CHAIN_LAG=8 GENFFT_CUDA_CHAIN_GRID_PCT=75" "GENFFT_CUDA_CHAIN_LAG=12 GENFFT_CUDA_CHAIN_GRID_PCT=75" "GENFFT_CUDA_CHAIN_LAG=16 GENFFT_CUDA_CHAIN_GRID_PCT=75" "GENFFT_CUDA_CHAIN_GRID_PCT=85" "GENFFT_CUDA_CHAIN_GRID_PCT=65" \
  "GENFFT_CUDA_CHAIN_LAG=8 GENFFT_CUDA_CHAIN_KB=2048" "GENFFT_CUDA_CHAIN_LAG=16 GENFFT_CUDA_CHAIN_KB=2048" "GENFFT_CUDA_CHAIN_LAG=12 GENFFT_CUDA_PEER_MODES=0" "GENFFT_CUDA_CHAIN_LAG=8 GENFFT_CUDA_CHAIN_GRID_PCT=75 GENFFT_CUDA_PEER_MODES=0"; do
  port=$((port+1))
  env $env $TR --master-port $port bench_dist.py --phases --transports p2p --chunks 1 --steps 10 --outputs natural >> gpurun_out/c5_dist_2gpu_sweep2.jsonl 2>> gpurun_out/c5_dist_2gpu_sweep2.err
done
python - <<'PY'
import json
for l in open('gpurun_out/c5_dist_2gpu_sweep2.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); k = {a.replace('GENFFT_CUDA_', ''): b for a, b in d['knobs'].items()}
        print(round(d['ms'], 3), {a: b[1] for a, b in d['phases_ms_rank0_and_max'].items() if 'transpose' in a}, k)
PY
