#!/bin/bash
# Round 2, last 2-GPU check of the final tree: multi-process tests and the contract bench as the driver launches it.
set -u
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_dist.py tests/test_gpu_chain.py -m gpu -q -x) > gpurun_out/pytest_dist_2gpu.log 2>&1; tail -3 gpurun_out/pytest_dist_2gpu.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 30511 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/bench_2gpu_final.json 2> gpurun_out/bench_2gpu_final.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_2gpu_final.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches')}, d['e2e']['value'], d['clocks'])
c = d['extras']['C5_dist']
print(c['natural_order']); print(c['transposed_output']['ms'], c['parity'], c.get('one_gpu_ms'), c.get('child_wall_s'))
PY
tail -3 gpurun_out/bench_2gpu_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 30512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>/dev/null | head -c 400; echo
