#!/bin/bash
# C3 with 8 double-complex points per thread at 64 registers (twice the resident warps), chained and not.
set -u
mkdir -p gpurun_out
: > gpurun_out/c3_p8.log
for env in "X=1" "GENFFT_CUDA_P_F64=8 GENFFT_CUDA_WIDE_C_F64=8" "GENFFT_CUDA_P_F64=8 GENFFT_CUDA_WIDE_C_F64=8 GENFFT_CUDA_CHAIN=0" "GENFFT_CUDA_P_F64=8 GENFFT_CUDA_WIDE_C_F64=16" "GENFFT_CUDA_CHAIN=0"; do
  echo "== $env" >> gpurun_out/c3_p8.log
  env $env timeout 120 python tools/variant_bench.py lib c3 2>&1 | grep -v "^==" | cut -c1-200 >> gpurun_out/c3_p8.log
done
cut -c1-170 gpurun_out/c3_p8.log
GENFFT_CUDA_P_F64=8 GENFFT_CUDA_WIDE_C_F64=8 python -m pytest tests/test_gpu_c2c.py -m gpu -q -k "float64 and (large or pow2 or c3)" 2>&1 | tail -3
(time python bench.py) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_1gpu.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
for k, v in d['extras'].items():
    if k != 'cpu_reference_1_thread': print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != 'plan'})
PY
