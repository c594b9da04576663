"""L2-resident pass chains (genfft_b200/csrc/chain_kernel.cuh) against the same passes launched one by one.

A chain runs two consecutive passes of a multi-pass transform in one launch with the intermediate kept in L2; the
arithmetic of every tile is the same code as the stand-alone pass, so the results must be BIT-identical to the
unchained execution (GENFFT_CUDA_CHAIN=0), whose parity against genFFT's CPU output the other test modules pin.
The launch counter shows that the chain really ran (one launch instead of two).
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402

CPX = {np.float32: torch.complex64, np.float64: torch.complex128}


class chain:
    def __init__(self, on, **env):
        self.env = {"GENFFT_CUDA_CHAIN": "1" if on else "0", **{k: str(v) for k, v in env.items()}}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        os.environ.update(self.env)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run_counted(fn):
    torch.cuda.synchronize()
    n0 = g.launch_count()
    fn()
    torch.cuda.synchronize()
    return g.launch_count() - n0


def rand_c(shape, dt, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.view_as_complex(torch.rand((*shape, 2), generator=gen, device="cuda", dtype=torch.float32 if dt == np.float32 else torch.float64) * 2 - 1)


# (log2 n, batch, passes, launches when chained)
C2C_CASES = [
    (np.float32, 15, 1, 2, 1), (np.float32, 15, 37, 2, 1), (np.float32, 16, 64, 2, 1), (np.float32, 18, 5, 2, 1),
    (np.float32, 21, 3, 3, 2), (np.float32, 22, 2, 3, 2), (np.float32, 24, 1, 3, 2),
    (np.float64, 14, 9, 2, 1), (np.float64, 15, 4, 2, 1), (np.float64, 16, 3, 2, 1), (np.float64, 18, 2, 2, 1),
    (np.float64, 20, 2, 3, 2), (np.float64, 21, 1, 3, 2), (np.float64, 24, 1, 3, 2),
]


@pytest.mark.parametrize("dt,lg,batch,passes,chained_launches", C2C_CASES)
@pytest.mark.parametrize("inv", [False, True])
def test_c2c_chain_is_bit_identical(dt, lg, batch, passes, chained_launches, inv):
    n = 1 << lg
    x = rand_c((batch, n), dt, lg * 100 + batch)
    plan = g.FFT(n, dt, batch=batch)
    assert plan.num_passes == passes, plan.describe()
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    with chain(False):
        l0 = run_counted(lambda: plan.transform(y0, x, inv))
    with chain(True):
        l1 = run_counted(lambda: plan.transform(y1, x, inv))
    assert l0 == passes
    assert l1 == chained_launches, plan.describe()
    assert torch.equal(y0, y1), plan.describe()
    # twice more on the same plan: the ticket/group counters are re-armed by every launch
    with chain(True):
        for _ in range(2):
            y1.zero_()
            plan.transform(y1, x, inv)
            assert torch.equal(y0, y1)


@pytest.mark.parametrize("dt,lg", [(np.float32, 16), (np.float32, 21), (np.float64, 20)])
def test_c2c_chain_vs_reference(comparand, dt, lg):
    """the chained execution itself against genFFT's CPU output (same tolerance as test_gpu_c2c)"""
    n = 1 << lg
    rng = np.random.default_rng(lg)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64 if dt == np.float32 else np.complex128)
    want = comparand.c2c(x, False)
    d = torch.from_numpy(x).cuda()
    y = torch.empty_like(d)
    with chain(True):
        g.FFT(n, dt).transform(y, d, False)
    assert oracle.rel_l2(y.cpu().numpy(), want) <= oracle.tolerance(n, dt)


@pytest.mark.parametrize("kb,lag", [(256, 1), (1024, 3), (16384, 2), (65536, 1)])
def test_chain_group_size_and_lag(kb, lag):
    """any group size / lag gives the same bits (group boundaries and the ticket order are pure scheduling)"""
    n, batch = 1 << 16, 24
    x = rand_c((batch, n), np.float32, 7)
    plan = g.FFT(n, np.float32, batch=batch)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    with chain(False):
        plan.transform(y0, x)
    with chain(True, GENFFT_CUDA_CHAIN_KB=kb, GENFFT_CUDA_CHAIN_LAG=lag):
        plan.transform(y1, x)
    assert torch.equal(y0, y1)
    n = 1 << 21
    x = rand_c((2, n), np.float32, 8)
    plan = g.FFT(n, np.float32, batch=2)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    with chain(False):
        plan.transform(y0, x)
    with chain(True, GENFFT_CUDA_CHAIN_KB=kb, GENFFT_CUDA_CHAIN_LAG=lag):
        plan.transform(y1, x)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("lg,batch,half", [(16, 5, True), (17, 3, False), (19, 2, True), (22, 3, True), (23, 1, True)])
def test_r2c_chain_is_bit_identical(lg, batch, half):
    n = 1 << lg
    gen = torch.Generator(device="cuda").manual_seed(lg)
    x = torch.rand((batch, n), generator=gen, device="cuda") * 2 - 1
    nout = n // 2 + 1 if half else n
    plan = g.RealFFT(n, np.float32, half=half, batch=batch)
    y0 = torch.zeros((batch, nout), dtype=torch.complex64, device="cuda")
    y1 = torch.zeros_like(y0)
    with chain(False):
        l0 = run_counted(lambda: plan.forward(y0, x))
    with chain(True):
        l1 = run_counted(lambda: plan.forward(y1, x))
    assert torch.equal(y0, y1), plan.describe()
    # whole transforms are the chain's groups here (the split pairs bins q and n/2 - q); above 8 MiB per transform
    # they no longer sit in L2 between the passes and the plan runs its passes one by one
    assert l1 == (l0 - 1 if n // 2 * 8 <= (8 << 20) else l0), (l0, l1, plan.describe())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("w,h", [(32768, 64), (64, 32768), (4096, 4096), (1 << 16, 4096)])
@pytest.mark.parametrize("inv", [False, True])
def test_fft2d_chain_is_bit_identical(dt, w, h, inv):
    if dt == np.float64 and w * h > (1 << 26):
        pytest.skip("keeps the case small")
    x = rand_c((h, w), dt, w + h)
    plan = g.FFT2D(w, h, dt)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    with chain(False):
        l0 = run_counted(lambda: plan.transform(y0, x, inv=inv))
    with chain(True):
        l1 = run_counted(lambda: plan.transform(y1, x, inv=inv))
    assert torch.equal(y0, y1), plan.describe()
    assert l1 < l0, (l0, l1, plan.describe())


@pytest.mark.parametrize("cols", [128, 96, 33])
def test_vert_chain_is_bit_identical(cols):
    n = 4096
    x = rand_c((n, cols), np.float32, cols)
    plan = g.FFTVert(n, np.float32)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    with chain(False):
        plan.transform(y0, x, cols)
    with chain(True):
        plan.transform(y1, x, cols)
    assert torch.equal(y0, y1)


@pytest.mark.no_emu  # needs real streams
def test_chained_plan_across_many_streams():
    """A chained plan executed on more streams than it keeps counter blocks for (32): the blocks are recycled, results
    stay bit-identical."""
    n, batch = 1 << 16, 4
    x = torch.view_as_complex(torch.rand((batch, n, 2), device="cuda") * 2 - 1)
    plan = g.FFT(n, np.float32, batch=batch)
    want = torch.empty_like(x)
    plan.forward(want, x)
    torch.cuda.synchronize()
    for k in range(40):
        s = torch.cuda.Stream()
        y = torch.empty_like(x)
        with torch.cuda.stream(s):
            plan.forward(y, x)
        s.synchronize()
        assert torch.equal(y, want)
