"""Scratch: e2e (host-pointer) timing of C2 for several staging chunk sizes, plus raw PCIe copy rates."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g
N, B = 4096, 1 << 16
hx = torch.empty((B, N), dtype=torch.complex64, pin_memory=True); hy = torch.empty_like(hx).pin_memory()
hx.real.uniform_(-1, 1); hx.imag.uniform_(-1, 1)
d = torch.empty((B, N), dtype=torch.complex64, device="cuda"); d2 = torch.empty_like(d)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
gb = hx.numel() * 8 / 1e9
print(f"H2D alone {gb / t(lambda: d.copy_(hx, non_blocking=True)):.1f} GB/s; D2H alone {gb / t(lambda: hy.copy_(d, non_blocking=True)):.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2): hy.copy_(d2, non_blocking=True)
print(f"H2D+D2H concurrent: {gb / t(both):.1f} GB/s each direction")
for mb in (8, 16, 32, 64, 128, 256):
    os.environ["GENFFT_CUDA_HOST_CHUNK_MB"] = str(mb)
    plan = g.FFT(N, np.float32, batch=B)
    dt = t(lambda: plan.forward(hy, hx))
    print(f"chunk {mb:4d} MiB: {dt * 1e3:.2f} ms  -> {5 * N * 12 * B / dt / 1e9:.1f} GFLOP/s")
