#!/bin/bash
# Round 2, the 8-GPU call: multi-process tests, contract bench at 8 and 4 GPUs (extras.C5_dist with parity), knob A/B of
# the slab 2D transform at 8 GPUs, distributed 1D, PCIe concurrency / host topology.
set -u
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1; nproc >> gpurun_out/topo_8gpu.txt; (numactl -H || lscpu | grep -i numa) >> gpurun_out/topo_8gpu.txt 2>&1; cat /sys/devices/system/node/online >> gpurun_out/topo_8gpu.txt; free -g >> gpurun_out/topo_8gpu.txt
(time $TR8 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -c 1800 gpurun_out/bench_8gpu.json; echo; tail -4 gpurun_out/bench_8gpu.err
(time python -m pytest tests/test_gpu_dist.py -m gpu -q -x) > gpurun_out/pytest_dist_8gpu.log 2>&1; tail -3 gpurun_out/pytest_dist_8gpu.log
port=29920
: > gpurun_out/c5_dist_8gpu_sweep.jsonl
for env in "X=1" "GENFFT_CUDA_PEER_MODES=0" "GENFFT_CUDA_CHAIN_LAG=5" "GENFFT_CUDA_CHAIN_LAG=12" "GENFFT_CUDA_CHAIN_LAG=16" "GENFFT_CUDA_CHAIN_GRID_PCT=75" "GENFFT_CUDA_CHAIN_KB=2048 GENFFT_CUDA_CHAIN_LAG=16"; do
  port=$((port+1))
  env $env $TR8 --master-port $port bench_dist.py --phases --transports p2p --chunks 1 --steps 10 --outputs natural >> gpurun_out/c5_dist_8gpu_sweep.jsonl 2>> gpurun_out/c5_dist_8gpu_sweep.err
done
$TR8 --master-port 29940 bench_dist.py --phases --transports p2p --chunks 1 --steps 10 --outputs transposed >> gpurun_out/c5_dist_8gpu_sweep.jsonl 2>> gpurun_out/c5_dist_8gpu_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/c5_dist_8gpu_sweep.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); k = {a.replace('GENFFT_CUDA_', ''): b for a, b in d['knobs'].items()}
        print(d['output'][:10], round(d['ms'], 3), {a: b[1] for a, b in d['phases_ms_rank0_and_max'].items()}, k)
PY
$TR8 --master-port 29941 bench_dist.py --one-d 28 --transports p2p --steps 10 > gpurun_out/dist1d_8gpu.jsonl 2>> gpurun_out/c5_dist_8gpu_sweep.err; cut -c1-330 gpurun_out/dist1d_8gpu.jsonl
python tools/pcie_concurrency.py > gpurun_out/pcie_concurrency_8gpu.json 2> gpurun_out/pcie_concurrency.err; python -c "
import json; d=json.load(open('gpurun_out/pcie_concurrency_8gpu.json'))
for p,r in d['placements'].items():
    print(p, r['gpu_numa_nodes'])
    for s,v in r['sets'].items(): print('  ', s, {m: v[m]['aggregate_gbs_per_direction'] for m in v})
"
(time $TR4 --master-port 29950 bench.py --gpus 4 --steps 20 --warmup 5) > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; tail -c 1500 gpurun_out/bench_4gpu.json; echo
