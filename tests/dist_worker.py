"""Worker for tests/test_gpu_dist.py (launched with torch.distributed.run, one rank per GPU):
slab-decomposed 2D FFT and four-step 1D FFT on small arrays, both transports, against genFFT's CPU output of the
full array (oracle.Ref: FFT2D::transform, fft.h:213-241 / FFT::transform, fft.h:80-85)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (tests may use the checker)
from genfft_b200.dist import DistFFT1D, DistFFT2D, four_step_shape  # noqa: E402

# the parity target is genFFT's own CPU output (the compiled reference travels to the GPU box as oracle/_ref);
# numpy (verified against it in the survey) stands in only when that build is absent
REF = oracle.Ref() if oracle.have_ref() else None


def ref_fft(x, inv):
    if REF is not None and x.shape[-1] <= (1 << 23):
        return REF.c2c(x, inv)
    x64 = x.astype(np.complex128)
    return np.fft.ifft(x64) * x.shape[-1] if inv else np.fft.fft(x64)


def ref_fft2(x, inv):
    if REF is not None:
        return REF.fft2d(x, inv)
    x64 = x.astype(np.complex128)
    return np.fft.ifft2(x64) * x.size if inv else np.fft.fft2(x64)


def one_d_cases(rank, world):
    """Distributed four-step 1D transform against numpy's fft of the whole sequence."""
    fails = 0
    for lg, dt in ((6, np.float32), (12, np.float32), (20, np.float32), (22, np.float32), (16, np.float64), (21, np.float64)):
        n = 1 << lg
        try:
            h, w = four_step_shape(n, world)
        except ValueError:
            continue
        cd = np.complex64 if dt == np.float32 else np.complex128
        rng = np.random.default_rng(lg)
        full = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cd)
        want_f, want_i = ref_fft(full, False), ref_fft(full, True)
        tol = (1e-6 if dt == np.float32 else 1e-14) * lg
        shard = torch.from_numpy(full[rank * n // world:(rank + 1) * n // world].copy()).cuda()
        for transport, bar in (("nccl", "collective"), ("p2p", "flags"), ("p2p", "collective")):
            for transposed in (False, True):
                plan = DistFFT1D(n, dt, transport=transport, transposed_out=transposed, barrier=bar)
                for inv, want in ((False, want_f), (True, want_i)):
                    for rep in range(2):  # twice: buffers are reused between calls
                        got = plan.transform(shard, inv)
                        torch.cuda.synchronize()
                    got = got.cpu().numpy()
                    if transposed:  # Z[kr][kc] = X[kr + H*kc], this rank's rows kr
                        ref = want.reshape(w, h).T[rank * h // world:(rank + 1) * h // world]
                    else:
                        ref = want[rank * n // world:(rank + 1) * n // world]
                    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                    ok = err <= tol
                    fails += not ok
                    if rank == 0 or not ok:
                        print(f"[rank {rank}] 1D n=2^{lg} {dt.__name__} {transport}/{bar} transposed={transposed} inv={inv}: "
                              f"rel-L2 {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
                dist.barrier()
                plan.close()
    return fails


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    fails = 0
    cases = [(64, 64, np.float32), (256, 128, np.float32), (2048, 512, np.float32), (512, 8192, np.float32),
             (32768, 64, np.float32), (64, 32768, np.float32), (1024, 1024, np.float64)]
    for w, h, dt in cases:
        if w % world or h % world:
            continue
        cd = np.complex64 if dt == np.float32 else np.complex128
        rng = np.random.default_rng(w + h)
        full = (rng.uniform(-1, 1, (h, w)) + 1j * rng.uniform(-1, 1, (h, w))).astype(cd)
        hl, wp = h // world, w // world
        want_f, want_i = ref_fft2(full, False), ref_fft2(full, True)
        tol = (1e-6 if dt == np.float32 else 1e-14) * np.log2(w * h)
        slab = torch.from_numpy(full[rank * hl:(rank + 1) * hl].copy()).cuda()
        # p2p: peer-memory flag barrier (default) and the collective-call barrier
        for transport, chunks, bar in (("nccl", 1, "collective"), ("p2p", 1, "flags"), ("p2p", 1, "collective"),
                                       ("p2p", 4, "flags")):
            for transposed in (False, True):
                if chunks > 1 and (w // world) // chunks < 32:
                    continue
                plan = DistFFT2D(w, h, dt, transport=transport, transposed_out=transposed, chunks=chunks, barrier=bar)
                for inv, want in ((False, want_f), (True, want_i)):
                    for rep in range(2):  # twice: buffers are reused between calls
                        got = plan.transform(slab, inv)
                        torch.cuda.synchronize()
                    got = got.cpu().numpy()
                    ref = want[:, rank * wp:(rank + 1) * wp] if transposed else want[rank * hl:(rank + 1) * hl]
                    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                    ok = err <= tol
                    fails += not ok
                    if rank == 0 or not ok:
                        print(f"[rank {rank}] {w}x{h} {dt.__name__} {transport}/{bar} chunks={chunks} transposed={transposed} inv={inv}: "
                              f"rel-L2 {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
                dist.barrier()
                plan.close()
    fails += one_d_cases(rank, world)
    t = torch.tensor([fails], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("DIST PASSED" if t.item() == 0 else f"DIST FAILED ({t.item()})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
