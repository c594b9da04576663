"""Scratch: the C2 contract kernel (batched N=4096 x 2^16 fp32) timed the way bench.py times it (K back-to-back launches
between two events), over compile-time variants of the library and values of GENFFT_CUDA_TMA_TILES.
usage: python tools/c2_sweep.py <lib_dir>[,<lib_dir>...] <tiles>[,<tiles>...] [reps] [launches]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genfft_b200 as g
from genfft_b200 import _lib

libs = sys.argv[1].split(",")
tiles = [int(t) for t in sys.argv[2].split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 40
n, batch = 4096, 1 << 16
x = torch.randn(batch, n, dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
nbytes = 2 * x.numel() * 8
for name in libs:
    _lib._lib = None
    _lib.LIB_PATH = os.path.join(ROOT, "genfft_b200", name, "libgenfft_cuda.so")
    if not os.path.exists(_lib.LIB_PATH):
        print(f"== {name}: not built", flush=True)
        continue
    res = {t: [] for t in tiles}
    for r in range(reps):
        for t in tiles:
            os.environ["GENFFT_CUDA_TMA_TILES"] = str(t)
            p = g.FFT(n, np.float32, batch=batch)
            for _ in range(5):
                p.transform(y, x)
            torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(launches):
                p.transform(y, x)
            b.record(); b.synchronize()
            res[t].append(a.elapsed_time(b) / launches)
            del p
    for t in tiles:
        v = sorted(res[t])
        print(f"{name} tiles={t}: ms/launch min {v[0]:.4f} med {v[len(v)//2]:.4f} max {v[-1]:.4f}  {nbytes / v[len(v)//2] / 1e6:.0f} GB/s (med)", flush=True)
        print(f"   in order: " + " ".join(f"{x:.4f}" for x in res[t]), flush=True)
