"""Scratch: run each large config once (after one warm-up) so that an ncu launch list shows per-pass times."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g
which = sys.argv[1:] or ["c3", "c4", "c5"]
if "c3" in which:
    n = 1 << 24
    p = g.FFT(n, np.float64); x = torch.randn(n, dtype=torch.complex128, device="cuda"); y = torch.empty_like(x)
    print(p.describe())
    for _ in range(2): p.forward(y, x)
    torch.cuda.synchronize(); del x, y
if "c4" in which:
    n, b = 1 << 22, 256
    p = g.RealFFT(n, np.float32, half=True, batch=b); x = torch.randn(b, n, device="cuda"); y = torch.empty((b, n // 2 + 1), dtype=torch.complex64, device="cuda")
    print(p.describe())
    for _ in range(2): p.forward(y, x)
    torch.cuda.synchronize(); del x, y
if "c5" in which:
    w = h = 32768
    p = g.FFT2D(w, h, np.float32); x = torch.randn(h, w, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
    print(p.describe())
    for _ in range(2): p.transform(y, x)
    torch.cuda.synchronize()
