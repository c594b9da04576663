#!/bin/bash
# Multi-GPU numbers of a round:  gpurun --gpus N --timeout 400 -- bash tools/round_start_multi.sh N
# C5 (2D 32768^2) and the distributed four-step 1D transform on N GPUs, the multi-GPU tests, bench.py at N.
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
timeout 120 $RUN --master-port 29571 bench_dist.py --transports p2p,nccl --chunks 1 --steps 10 --warmup 3 --phases > gpurun_out/c5_${N}gpu.jsonl 2> gpurun_out/c5_${N}gpu.err
for lg in 28 30; do
  timeout 120 $RUN --master-port 29572 bench_dist.py --one-d $lg --transports p2p,nccl --steps 10 --warmup 3 >> gpurun_out/one_d_${N}gpu.jsonl 2>> gpurun_out/one_d_${N}gpu.err
done
timeout 120 $RUN --master-port 29573 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/*_${N}gpu.json*")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], d.get("workload", d.get("metric", ""))[:60], d.get("transport", ""), d.get("output", "")[:12],
                  "ms", round(d.get("ms", d.get("ms_per_step", 0)), 3), "gflops", round(d.get("gflops", d.get("value", 0))))
PY
