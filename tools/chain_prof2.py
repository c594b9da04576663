"""Scratch: chained 2-pass batch for a full ncu capture."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g
n, b = 1 << 16, 2048
p = g.FFT(n, np.float32, batch=b)
x = torch.randn(b, n, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
for _ in range(2): p.forward(y, x)
torch.cuda.synchronize()
