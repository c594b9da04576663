#!/bin/bash
# plan-knob sweeps on the large configs (tools/quick_bench.py)
run() { echo "== $*"; env "$@" python tools/quick_bench.py $WHICH 2>&1 | cut -c1-110; }
WHICH="${WHICH:-c3 c4 2d}"
run GENFFT_CUDA_MAXLEN_F32=512
run GENFFT_CUDA_MAXLEN_F32=1024
