#!/bin/bash
# Round 2, last call: the tree as it will be judged -- smoke, the whole -m gpu suite, the default bench line.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_final.log
(time python -m pytest tests -m gpu -q -x --durations=3) > gpurun_out/pytest_gpu_final.log 2>&1; tail -n 9 gpurun_out/pytest_gpu_final.log
(time python bench.py) > gpurun_out/bench_1gpu_final2.json 2> gpurun_out/bench_1gpu_final2.err; head -c 400 gpurun_out/bench_1gpu_final2.json; echo; tail -n 4 gpurun_out/bench_1gpu_final2.err
