#!/bin/bash
# Round 2, second 8-GPU call: the fused distributed 1D transform (one-pass vs split rows), C5 once more, dist tests.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=30300
: > gpurun_out/dist_8gpu_final.jsonl
for env in "X=1" "GENFFT_CUDA_DIST_ROWS_SINGLE=2048"; do
  port=$((port+1))
  env $env $TR --master-port $port bench_dist.py --one-d 28 --transports p2p --steps 10 --phases >> gpurun_out/dist_8gpu_final.jsonl 2>> gpurun_out/dist_8gpu_final.err
done
port=$((port+1))
$TR --master-port $port bench_dist.py --one-d 30 --transports p2p --steps 10 --phases >> gpurun_out/dist_8gpu_final.jsonl 2>> gpurun_out/dist_8gpu_final.err
port=$((port+1))
$TR --master-port $port bench_dist.py --phases --transports p2p --chunks 1 --steps 10 >> gpurun_out/dist_8gpu_final.jsonl 2>> gpurun_out/dist_8gpu_final.err
python - <<'PY'
import json
for l in open('gpurun_out/dist_8gpu_final.jsonl'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['workload'][:34], d['output'][:12], round(d['ms'], 3), d.get('frac_of_nvlink_roofline_770'), d.get('phases_ms_max_over_ranks') or {a: b[1] for a, b in d['phases_ms_rank0_and_max'].items()}, {a.replace('GENFFT_CUDA_', ''): b for a, b in d.get('knobs', {}).items()})
PY
(time python -m pytest tests/test_gpu_dist.py -m gpu -q -x) > gpurun_out/pytest_dist_8gpu.log 2>&1; tail -2 gpurun_out/pytest_dist_8gpu.log
tail -3 gpurun_out/dist_8gpu_final.err
