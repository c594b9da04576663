"""Scratch: latency of a small forward + inverse pair (BASELINE config C1 and neighbours) through the C loop inside the
library, raw ctypes calls and one CUDA graph; run with GENFFT_CUDA_PDL=0/1 to A/B programmatic dependent launch."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g

def timed(fn, iters, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters * 1e3

print("GENFFT_CUDA_PDL =", os.environ.get("GENFFT_CUDA_PDL", "(default)"))
for n in (256, 1024, 4096, 16384):
    p = g.FFT(n, np.float32)
    x = torch.randn(n, dtype=torch.complex64, device="cuda"); y, z = torch.empty_like(x), torch.empty_like(x)
    lib, h, st = g.lib(), p._h, torch.cuda.current_stream().cuda_stream
    us = ctypes.c_double(0.0)
    assert lib.genfft_cuda_debug_time_c2c_pairs(h, z.data_ptr(), y.data_ptr(), x.data_ptr(), 4000, st, ctypes.byref(us)) == 0
    raw = timed(lambda: (lib.genfft_cuda_exec_c2c_dev(h, y.data_ptr(), x.data_ptr(), 0, st), lib.genfft_cuda_exec_c2c_dev(h, z.data_ptr(), y.data_ptr(), 1, st)), 1000, 100)
    graph_us = None
    try:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            p.forward(y, x); p.inverse(z, y)
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            p.forward(y, x); p.inverse(z, y)
        graph_us = timed(gr.replay, 1000, 100)
    except Exception as e:
        graph_us = repr(e)[:80]
    torch.cuda.synchronize()
    err = float((z / n - x).abs().max())
    print(f"n={n}: C loop {us.value:.2f} us/pair, raw ctypes {raw:.2f}, graph {graph_us}, round trip err {err:.1e}", flush=True)
