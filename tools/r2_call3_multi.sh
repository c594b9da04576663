#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the multi-process tests, the contract bench with extras.C5_dist, and the
# peer-mode A/B of the slab 2D transform.  usage: bash tools/r2_call3_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(time python -m pytest tests/test_gpu_dist.py -m gpu -q -x) > gpurun_out/pytest_dist_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_dist_${N}gpu.log
(time $TR --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 1500 gpurun_out/bench_${N}gpu.json; echo; tail -5 gpurun_out/bench_${N}gpu.err
$TR --master-port 29612 bench_dist.py --phases --transports p2p --chunks 1 --steps 10 > gpurun_out/c5_dist_${N}gpu_peer_modes.jsonl 2> gpurun_out/c5_dist_${N}gpu.err; cut -c1-400 gpurun_out/c5_dist_${N}gpu_peer_modes.jsonl
GENFFT_CUDA_PEER_MODES=0 $TR --master-port 29613 bench_dist.py --phases --transports p2p --chunks 1 --steps 10 > gpurun_out/c5_dist_${N}gpu_generic_store.jsonl 2>> gpurun_out/c5_dist_${N}gpu.err; cut -c1-400 gpurun_out/c5_dist_${N}gpu_generic_store.jsonl
$TR --master-port 29614 bench_dist.py --one-d 28 --transports p2p --steps 10 > gpurun_out/dist1d_${N}gpu.jsonl 2>> gpurun_out/c5_dist_${N}gpu.err; cut -c1-300 gpurun_out/dist1d_${N}gpu.jsonl
if [ "$N" -ge 4 ]; then
  python tools/pcie_concurrency.py > gpurun_out/pcie_concurrency_${N}gpu.json 2> gpurun_out/pcie_concurrency.err; head -c 3000 gpurun_out/pcie_concurrency_${N}gpu.json
  nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1; nproc >> gpurun_out/topo_${N}gpu.txt; numactl -H >> gpurun_out/topo_${N}gpu.txt 2>&1; cat /sys/devices/system/node/online >> gpurun_out/topo_${N}gpu.txt
fi
