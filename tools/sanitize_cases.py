"""Scratch: one small launch of every kernel mode, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g
if "2x" in sys.argv[1:]:  # only RealFFT2D::forward_2x: first pass reading two images (1 and 2 row passes), and the copy path
    for dt, cd in ((np.float32, torch.complex64), (np.float64, torch.complex128)):
        rd = torch.float32 if dt == np.float32 else torch.float64
        for w, h in ((256, 64), (32768, 4), (16, 256)):
            a = torch.randn(h, w, dtype=rd, device="cuda"); b = torch.randn(h, w + 2, dtype=rd, device="cuda")
            y = torch.empty(h, w, dtype=cd, device="cuda"); p = g.RealFFT2D(w, h, dt)
            p.forward_2x(y, a, a.clone()); p.forward_2x(y, a, b, in_stride2=w + 2)
    torch.cuda.synchronize()
    print("2x cases done")
    sys.exit(0)
if "dist" in sys.argv[1:]:  # round 2: peer-store modes (chained and not), fused four-step twiddle, one-launch scatter,
    # played by one process on one device (the peers' buffers are local allocations)
    from genfft_b200.dist import CudaSlabEngine, four_step_shape
    for dt, cd in ((np.float32, torch.complex64), (np.float64, torch.complex128)):
        for world in (2, 4, 8):
            w, h = 32768, 8
            hl, wp = h // world, w // world
            eng = [CudaSlabEngine(w, h, world, dt) for _ in range(world)]
            blocks = [torch.zeros((h, wp), dtype=cd, device="cuda") for _ in range(world)]
            for r in range(world):
                x = torch.randn(hl, w, dtype=cd, device="cuda")
                eng[r].rows_to_peers(x, [b.data_ptr() for b in blocks], r, bool(r & 1))
            n = 1 << 24
            hh, ww = four_step_shape(n, world)
            e1 = [CudaSlabEngine(ww, hh, world, dt) for _ in range(world)]
            blk = [torch.zeros((hh, ww // world), dtype=cd, device="cuda") for _ in range(world)]
            mids = [torch.zeros((hh // world, ww), dtype=cd, device="cuda") for _ in range(world)]
            for r in range(world):
                e1[r].cols_blocks_to_peers(torch.randn(hh // world, ww, dtype=cd, device="cuda"), [b.data_ptr() for b in blk], r)
            torch.cuda.synchronize()
            for r in range(world):
                assert e1[r].cols_to_peers(blk[r].data_ptr(), [m.data_ptr() for m in mids], r, False, twiddle_n=n)
            torch.cuda.synchronize()
            del eng, e1, blocks, blk, mids
    print("dist cases done")
    sys.exit(0)
for dt, cd in ((np.float32, torch.complex64), (np.float64, torch.complex128)):
    rd = torch.float32 if dt == np.float32 else torch.float64
    # M_ROW / M_ROWTMA (needs >= 4 tiles per SM)
    for n, b in ((4096, 3), (1024, 2400), (4096, 600), (64, 7), (8192, 2)):
        x = torch.randn(b, n, dtype=cd, device="cuda"); y = torch.empty_like(x)
        p = g.FFT(n, dt, batch=b); p.forward(y, x); p.inverse(x, y)
    # M_FIRST / M_COLTW (2 and 3 passes), M_GEN (no_scramble, real input)
    for n in (1 << 15, 1 << 19):
        x = torch.randn(2, n, dtype=cd, device="cuda"); y = torch.empty_like(x)
        p = g.FFT(n, dt, batch=2); p.forward(y, x); p.inverse(x, y)
    x = torch.randn(1 << 15, dtype=cd, device="cuda"); g.FFT(1 << 15, dt).transform_no_scramble(x)
    r = torch.randn(4096, dtype=rd, device="cuda"); y = torch.empty(4096, dtype=cd, device="cuda"); g.FFT(4096, dt).transform_real(y, r)
    # M_ROWDIT, M_COLTWDIT, stand-alone split, c2r
    for n, b in ((4096, 5), (1 << 17, 2), (1 << 19, 1)):
        r = torch.randn(b, n, dtype=rd, device="cuda"); y = torch.empty(b, n // 2 + 1, dtype=cd, device="cuda")
        g.RealFFT(n, dt, half=True, batch=b).forward(y, r)
        g.InverseRealFFT(n, dt, batch=b).inverse(r, y)
        y2 = torch.empty(b, n, dtype=cd, device="cuda"); g.RealFFT(n, dt, half=False, batch=b).forward(y2, r)
    # M_COL (vert), 2D, RealFFT2D
    x = torch.randn(256, 37, dtype=cd, device="cuda"); y = torch.empty_like(x); g.FFTVert(256, dt).transform(y, x, 37)
    x = torch.randn(8192, 40, dtype=cd, device="cuda"); y = torch.empty_like(x); g.FFTVert(8192, dt).transform(y, x, 40)
    x = torch.randn(512, 1024, dtype=cd, device="cuda"); y = torch.empty_like(x); g.FFT2D(1024, 512, dt).transform(y, x)
    r = torch.randn(64, 256, dtype=rd, device="cuda"); y = torch.empty(64, 256, dtype=cd, device="cuda"); g.RealFFT2D(256, 64, dt).forward(y, r); g.RealFFT2D(256, 64, dt).forward_2x(y, r, r.clone())
torch.cuda.synchronize()
print("cases done")
