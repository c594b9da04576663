"""BASELINE config C5 at FULL size on one GPU against genFFT's own CPU implementation.

FFT2D<float>(32768, 32768).transform (include/genFFT/fft.h:213-241) is rows first (scramble_row_fft: every input row
through FFT<T>::transform) and then the vertical transform of every column.  The whole reference output would take a
CPU thread about a minute and 16 GiB of host memory, so the comparand is built by the same two steps restricted to what
the check needs: ALL 32768 row transforms by the compiled reference (oracle.Ref.c2c_rows, the loop of fft.h:229-241 on
the host cores), then its vertical transform (FFTVert, fft.h:145-150) on a sample of columns spread over different
column tiles.  Every output element of those columns depends on every input row, so the sample exercises every tile of
the row passes and the column passes of the sampled tiles; two linear checksums (the sum over each axis of the output
is a 1D transform of one input row / column, again by the reference) cover every output element.
"""
import numpy as np
import pytest

import oracle

pytestmark = [pytest.mark.gpu, pytest.mark.no_emu]

torch = pytest.importorskip("torch")
import genfft_b200 as g  # noqa: E402

W = H = 32768
COLS = [0, 1, 15, 16, 4097, 16384, 20011, 32767]  # both halves, tile edges, odd places


def _need_memory():
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * (1 << 30):
        pytest.skip("needs ~26 GiB of device memory")


@pytest.mark.parametrize("inv", [False, True])
def test_c5_full_size_vs_reference(checkers, inv):
    ref, port = checkers
    if ref is None:
        pytest.skip("the compiled reference (oracle/_ref) is not present")
    _need_memory()
    gen = torch.Generator(device="cuda").manual_seed(20261018)
    x = torch.view_as_complex(torch.rand((H, W, 2), generator=gen, device="cuda") * 2 - 1)
    y = torch.empty_like(x)
    plan = g.FFT2D(W, H, np.float32)
    plan.transform(y, x, inv=inv)
    torch.cuda.synchronize()

    # rows by the reference, keeping the sampled columns only
    rows_s = np.empty((H, len(COLS)), np.complex64)
    chunk = 2048
    for r0 in range(0, H, chunk):
        rows = ref.c2c_rows(x[r0:r0 + chunk].cpu().numpy(), inv)
        rows_s[r0:r0 + chunk] = rows[:, COLS]
    want = ref.vert(rows_s, inv)  # FFTVert on the (H x 8) array of sampled columns
    got = y[:, COLS].cpu().numpy()
    tol = oracle.tolerance(W * H, np.float32)  # 1e-6 * 30
    err = oracle.rel_l2(got, want)
    assert err <= tol, f"sampled columns: rel-L2 {err:.3e} > {tol:.1e}"
    for j, c in enumerate(COLS):
        assert oracle.rel_l2(got[:, j], want[:, j]) <= tol, f"column {c}"

    # checksums over every output element: sum_kx X[ky, kx] = W * DFT_H(x[:, 0])[ky], sum_ky X[ky, kx] = H * DFT_W(x[0, :])[kx]
    col0 = ref.c2c(x[:, 0].contiguous().cpu().numpy(), inv).astype(np.complex128) * W
    row0 = ref.c2c(x[0].cpu().numpy(), inv).astype(np.complex128) * H
    s_rows = torch.sum(y, dim=1, dtype=torch.complex128).cpu().numpy()
    s_cols = torch.sum(y, dim=0, dtype=torch.complex128).cpu().numpy()
    assert oracle.rel_l2(s_rows, col0) <= tol
    assert oracle.rel_l2(s_cols, row0) <= tol
    del x, y
    torch.cuda.empty_cache()
