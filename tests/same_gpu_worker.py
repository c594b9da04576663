"""Worker for tests/test_gpu_ipc_same_device.py: TWO processes that share ONE GPU run the distributed transforms'
p2p transport for real -- CUDA-IPC mapped peer buffers, the st.release.sys / ld.acquire.sys flag barrier, stores into
the other process's memory -- so the inter-process machinery of genfft_b200/dist.py is exercised on a one-GPU box
(NCCL refuses two ranks on one device, so the process group is gloo: it carries only the IPC handles and the host
barriers; no collective is on the p2p data path).  The two contexts time-slice the device, so this is slow; it checks
correctness only.  Comparand: genFFT's own CPU output (oracle.Ref) when the compiled reference is present."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (tests may use the checker)
from genfft_b200.dist import DistFFT1D, DistFFT2D, four_step_shape  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(0)  # every rank on the same device
    dist.init_process_group("gloo")
    ref = oracle.Ref() if oracle.have_ref() else None
    fails = 0
    for w, h, dt in ((64, 64, np.float32), (2048, 512, np.float32), (512, 4096, np.float32), (256, 256, np.float64)):
        cd = np.complex64 if dt == np.float32 else np.complex128
        rng = np.random.default_rng(w + h)
        full = (rng.uniform(-1, 1, (h, w)) + 1j * rng.uniform(-1, 1, (h, w))).astype(cd)
        hl, wp = h // world, w // world
        if ref is not None:
            want = {False: ref.fft2d(full, False), True: ref.fft2d(full, True)}
        else:
            f64 = full.astype(np.complex128)
            want = {False: np.fft.fft2(f64), True: np.fft.ifft2(f64) * (w * h)}
        tol = (1e-6 if dt == np.float32 else 1e-14) * np.log2(w * h)
        slab = torch.from_numpy(full[rank * hl:(rank + 1) * hl].copy()).cuda()
        for transposed in (False, True):
            plan = DistFFT2D(w, h, dt, transport="p2p", transposed_out=transposed, barrier="flags")
            for inv in (False, True):
                for rep in range(2):  # twice: the buffers and the epoch flags are reused between calls
                    got = plan.transform(slab, inv)
                    torch.cuda.synchronize()
                got = got.cpu().numpy()
                r = want[inv][:, rank * wp:(rank + 1) * wp] if transposed else want[inv][rank * hl:(rank + 1) * hl]
                err = oracle.rel_l2(got, r)
                ok = err <= tol
                fails += not ok
                print(f"[rank {rank}] 2D {w}x{h} {dt.__name__} transposed={transposed} inv={inv}: rel-L2 {err:.2e} "
                      f"{'ok' if ok else 'FAIL'}", flush=True)
            dist.barrier()
            plan.close()
    for lg, dt in ((12, np.float32), (18, np.float32)):
        n = 1 << lg
        four_step_shape(n, world)
        cd = np.complex64 if dt == np.float32 else np.complex128
        rng = np.random.default_rng(lg)
        full = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
        want = ref.c2c(full) if ref is not None else np.fft.fft(full.astype(np.complex128))
        shard = torch.from_numpy(full[rank * n // world:(rank + 1) * n // world].copy()).cuda()
        plan = DistFFT1D(n, dt, transport="p2p", barrier="flags")
        for rep in range(2):
            got = plan.transform(shard, False)
            torch.cuda.synchronize()
        err = oracle.rel_l2(got.cpu().numpy(), want[rank * n // world:(rank + 1) * n // world])
        ok = err <= 1e-6 * lg
        fails += not ok
        print(f"[rank {rank}] 1D n=2^{lg}: rel-L2 {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        dist.barrier()
        plan.close()
    t = torch.tensor([fails])
    dist.all_reduce(t)
    if rank == 0:
        print("SAMEGPU PASSED" if t.item() == 0 else f"SAMEGPU FAILED ({t.item()})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
