#!/bin/bash
# Round 2, 2-GPU call: where the distributed 1D transform spends its time; one-pass vs split rows for the peer-storing pass.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time python -m pytest tests/test_gpu_dist.py -m gpu -q -x) > gpurun_out/pytest_dist_2gpu.log 2>&1; tail -2 gpurun_out/pytest_dist_2gpu.log
port=30100
: > gpurun_out/dist1d_2gpu_phases.jsonl
for env in "X=1" "GENFFT_CUDA_DIST_ROWS_SINGLE=2048" "GENFFT_CUDA_DIST_ROWS_SINGLE=512"; do
  port=$((port+1))
  env $env $TR --master-port $port bench_dist.py --one-d 28 --transports p2p --steps 10 --phases >> gpurun_out/dist1d_2gpu_phases.jsonl 2>> gpurun_out/dist1d_2gpu.err
  port=$((port+1))
  env $env $TR --master-port $port bench_dist.py --size 16384 --phases --transports p2p --chunks 1 --steps 10 --outputs natural >> gpurun_out/dist1d_2gpu_phases.jsonl 2>> gpurun_out/dist1d_2gpu.err
done
port=$((port+1))
$TR --master-port $port bench_dist.py --phases --transports p2p --chunks 1 --steps 10 --outputs natural >> gpurun_out/dist1d_2gpu_phases.jsonl 2>> gpurun_out/dist1d_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/dist1d_2gpu_phases.jsonl'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['workload'][:34], d['output'][:12], round(d['ms'], 3), d.get('phases_ms_max_over_ranks') or {a: b[1] for a, b in d['phases_ms_rank0_and_max'].items()}, {a.replace('GENFFT_CUDA_', ''): b for a, b in d.get('knobs', {}).items()})
PY
tail -3 gpurun_out/dist1d_2gpu.err
