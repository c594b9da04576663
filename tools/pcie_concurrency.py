"""Platform floor of the host-pointer (e2e) path on a multi-GPU box: concurrent pinned H2D + D2H on 1 / 2 / 4 / 8 GPUs.

bench.py's e2e figure moves 2 GiB each way per step and GPU; round 1's SCALE run showed the per-rank step time growing
45.8 -> 80 -> 164 -> 260 ms at N = 1 / 2 / 4 / 8, i.e. the host side, not the links, limits it.  This measures what the
platform allows, with the host buffers placed two ways: torch's pinned allocator (wherever the process runs) and
genfft_cuda_host_alloc (bound to the GPU's NUMA node).  One process drives all GPUs (copy engines are asynchronous), so
the result does not depend on NCCL or process placement.
  python tools/pcie_concurrency.py [--mib 1024] > gpurun_out/pcie_concurrency.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genfft_b200.hostmem import PinnedNearGpu  # noqa: E402


def measure(devs, bufs, mode, reps=3):
    """bufs[d] = (host_in, host_out, dev_in, dev_out, stream_in, stream_out); returns GB/s per direction summed over devs"""
    def issue():
        for d in devs:
            hi, ho, di, do, si, so = bufs[d]
            if mode in ("h2d", "both"):
                with torch.cuda.stream(si):
                    di.copy_(hi, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(so):
                    ho.copy_(do, non_blocking=True)

    def sync():
        for d in devs:
            torch.cuda.synchronize(d)

    issue()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        issue()
    sync()
    dt = (time.perf_counter() - t0) / reps
    nbytes = bufs[devs[0]][0].numel() * bufs[devs[0]][0].element_size()
    return nbytes * len(devs) / dt / 1e9, dt * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    elems = args.mib * (1 << 20) // 8
    out = {"gpus": n, "mib_per_direction_per_gpu": args.mib, "cpus_allowed": len(os.sched_getaffinity(0)), "placements": {}}
    for placement in ("torch_pinned_default", "numa_local_host_alloc"):
        bufs, keep, nodes = {}, [], {}
        for d in range(n):
            torch.cuda.set_device(d)
            if placement == "torch_pinned_default":
                hi = torch.empty(elems, dtype=torch.complex64, pin_memory=True)
                ho = torch.empty(elems, dtype=torch.complex64, pin_memory=True)
            else:
                a, b = PinnedNearGpu(elems, np.complex64), PinnedNearGpu(elems, np.complex64)
                keep += [a, b]
                nodes[d] = a.numa_node
                hi, ho = a.tensor(), b.tensor()
            hi.zero_()
            ho.zero_()
            di = torch.zeros(elems, dtype=torch.complex64, device=f"cuda:{d}")
            do = torch.zeros(elems, dtype=torch.complex64, device=f"cuda:{d}")
            bufs[d] = (hi, ho, di, do, torch.cuda.Stream(d), torch.cuda.Stream(d))
        rec = {"gpu_numa_nodes": nodes, "sets": {}}
        sets = [[0]] + [list(range(k)) for k in (2, 4, 8) if k <= n]
        if n >= 8:
            sets += [[0, 4], [4, 5, 6, 7]]
        for devs in sets:
            r = {}
            for mode in ("h2d", "d2h", "both"):
                gbs, ms = measure(devs, bufs, mode)
                r[mode] = {"aggregate_gbs_per_direction": round(gbs, 1), "ms": round(ms, 2)}
            rec["sets"][",".join(map(str, devs))] = r
        out["placements"][placement] = rec
        del bufs
        for k in keep:
            k.close()
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
