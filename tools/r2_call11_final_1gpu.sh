#!/bin/bash
# Round 2, final 1-GPU check of the tree: smoke, the whole -m gpu suite, both bench arms.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
(time python -m pytest tests -m gpu -q --durations=6 --maxfail=10) > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
(time python bench.py) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.json; echo
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>/dev/null; head -c 600 gpurun_out/bench_reference.json; echo
