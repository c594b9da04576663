#!/bin/bash
# Round 2, last 4-GPU check of the final tree.
set -u
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_dist.py -m gpu -q -x) > gpurun_out/pytest_dist_4gpu.log 2>&1; tail -3 gpurun_out/pytest_dist_4gpu.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 30611 bench.py --gpus 4 --steps 20 --warmup 5) > gpurun_out/bench_4gpu_final.json 2> gpurun_out/bench_4gpu_final.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_4gpu_final.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches')}, d['e2e']['value'])
c = d['extras']['C5_dist']
print(c['natural_order']); print(c['transposed_output']['ms'], c['parity'], c.get('one_gpu_ms'), c.get('child_wall_s'))
PY
