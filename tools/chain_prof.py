"""Scratch: one chained and one unchained execution of a 2-pass and a 3-pass batch, for an ncu metrics pass."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g
for n, b in ((1 << 16, 2048), (1 << 21, 64)):
    p = g.FFT(n, np.float32, batch=b)
    x = torch.randn(b, n, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
    print(p.describe())
    for on in ("0", "1"):
        os.environ["GENFFT_CUDA_CHAIN"] = on
        for _ in range(2): p.forward(y, x)
        torch.cuda.synchronize()
    del x, y
