"""Scratch: the C2 contract kernel in bench.py's regime -- an idle GPU, 5 warm-up launches, 20 timed launches between
two events -- over compile-time variants of the library, interleaved, with the GPU idling 2 s before every measurement.
usage: python tools/burst_c2.py <lib_dir>[,<lib_dir>...] [reps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genfft_b200 as g
from genfft_b200 import _lib
libs = sys.argv[1].split(",")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, batch = 4096, 1 << 16
x = torch.view_as_complex(torch.rand(batch, n, 2, device="cuda").mul_(2).sub_(1))
y = torch.empty_like(x)
res = {name: [] for name in libs}
for rep in range(reps):
    for name in libs:
        _lib._lib = None
        _lib.LIB_PATH = os.path.join(ROOT, "genfft_b200", name, "libgenfft_cuda.so")
        p = g.FFT(n, np.float32, batch=batch)
        torch.cuda.synchronize()
        time.sleep(2.0)
        for _ in range(5):
            p.transform(y, x)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            p.transform(y, x)
        b.record(); b.synchronize()
        res[name].append(a.elapsed_time(b) / 20)
        del p
for name in libs:
    v = res[name]
    print(f"burst {name}: " + " ".join(f"{t:.4f}" for t in v) + f" ms/launch; best {2 * x.numel() * 8 / min(v) / 1e6:.0f} GB/s", flush=True)
