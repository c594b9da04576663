#!/bin/bash
# A/B: inter-pass twiddles from 4 table rows + products (GENFFT_TW_POW) instead of 15 table rows.
set -u
mkdir -p gpurun_out
python tools/variant_bench.py lib,lib_exp_twpow,lib,lib_exp_twpow c3 c3f c4 c5 > gpurun_out/variant_twpow.log 2>&1; cut -c1-105 gpurun_out/variant_twpow.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_twpow/libgenfft_cuda.so python -m pytest tests/test_gpu_c2c.py tests/test_gpu_real_vert_2d.py tests/test_gpu_chain.py tests/test_gpu_c5_full.py -m gpu -q -x > gpurun_out/pytest_twpow.log 2>&1; tail -2 gpurun_out/pytest_twpow.log
GENFFT_CUDA_LIB=$PWD/genfft_b200/lib_exp_twpow/libgenfft_cuda.so python - <<'PY'
import numpy as np, torch, oracle, genfft_b200 as g
ref = oracle.Ref()
for n, dt in ((1 << 20, np.float32), (1 << 22, np.float32), (1 << 20, np.float64), (1 << 24, np.float64)):
    cd = np.complex64 if dt == np.float32 else np.complex128
    rng = np.random.default_rng(1)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
    d = torch.from_numpy(x).cuda(); y = torch.empty_like(d)
    g.FFT(n, dt).forward(y, d); torch.cuda.synchronize()
    print("twpow parity", n, dt.__name__, oracle.rel_l2(y.cpu().numpy(), ref.c2c(x)), oracle.tolerance(n, dt))
PY
