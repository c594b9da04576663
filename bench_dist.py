#!/usr/bin/env python
"""Multi-GPU benchmark of BASELINE config C5 (2D complex fp32 FFT 32768 x 32768, slab-decomposed) -- run under
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_dist.py
Not the driver's contract bench (that is bench.py, workload C2); prints one JSON line per transport / mode with
the effective GFLOP/s (5*W*H*log2(W*H)/t), and the time against the NVLink all-to-all roofline of SURVEY.md 8(d):
bytes sent per GPU per transpose = (8 GiB / P) * (P-1)/P at the measured 770 GB/s per direction (900 nominal).
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from genfft_b200.dist import DistFFT1D, DistFFT2D, four_step_shape  # noqa: E402


def bench_one_d(args, rank, world):
    """One large 1D C2C transform (four-step over the slab decomposition): n = 2^args.one_d points, fp32."""
    n = 1 << args.one_d
    h, w = four_step_shape(n, world)
    gen = torch.Generator(device="cuda").manual_seed(rank)
    shard = torch.view_as_complex(torch.rand((n // world, 2), generator=gen, device="cuda") * 2 - 1)
    flop = 5.0 * n * math.log2(n)
    sent = (n * 8 / world) * (world - 1) / world
    for transport in args.transports.split(","):
        for transposed in (True, False):
            plan = DistFFT1D(n, np.float32, transport=transport, transposed_out=transposed)
            for _ in range(args.warmup):
                plan.transform(shard)
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                plan.transform(shard)
            b.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            ntr = 2 if transposed else 3
            phases = None
            if args.phases and transport == "p2p":
                plan.start_phase_timing()
                for _ in range(args.steps):
                    plan.transform(shard)
                ph = plan.phase_times_ms()
                pmax = torch.tensor(list(ph.values()), device="cuda", dtype=torch.float64)
                dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
                phases = {k: round(float(v), 4) for k, v in zip(ph, pmax.tolist())}
            if rank == 0:
                print(json.dumps({
                    "workload": f"1D C2C fp32 n=2^{args.one_d} ({h} x {w} four-step), slab-decomposed over {world} GPUs",
                    "phases_ms_max_over_ranks": phases,
                    "knobs": {k: v for k, v in os.environ.items() if k.startswith("GENFFT_CUDA_")},
                    "transport": transport, "n_gpus": world, "ms": ms, "gflops": flop / (ms * 1e-3) / 1e9,
                    "output": f"transposed Z[kr][kc] = X[kr + {h} kc] (2 global transposes)" if transposed
                              else "natural order (3 global transposes)",
                    "alltoall_bytes_sent_per_gpu_per_transpose": sent,
                    "nvlink_floor_ms_at_770GBs": ntr * sent / 770e9 * 1e3,
                    "frac_of_nvlink_roofline_770": (ntr * sent / 770e9 * 1e3) / ms if world > 1 else None,
                }), flush=True)
            plan.close()
            del plan
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunks", type=str, default="1,4")
    ap.add_argument("--fracs", type=str, default="0.7:0.3")
    ap.add_argument("--transports", type=str, default="nccl,p2p")
    ap.add_argument("--phases", action="store_true", help="p2p: also report mean ms per phase (extra untimed pass)")
    ap.add_argument("--barriers", type=str, default="flags", help="p2p transport: flags (peer memory) and/or collective")
    ap.add_argument("--outputs", type=str, default="transposed,natural", help="which output orders to time")
    ap.add_argument("--one-d", type=int, default=0, help="instead of C5: one 1D transform of 2^ONE_D points (DistFFT1D)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    if args.one_d:
        bench_one_d(args, rank, world)
        dist.destroy_process_group()
        return
    w = h = args.size
    hl = h // world
    gen = torch.Generator(device="cuda").manual_seed(rank)
    slab = torch.view_as_complex(torch.rand((hl, w, 2), generator=gen, device="cuda") * 2 - 1)
    flop = 5.0 * w * h * math.log2(w * h)
    sent = (w * h * 8 / world) * (world - 1) / world
    variants = []
    orders = [o == "transposed" for o in args.outputs.split(",")]
    for transport in args.transports.split(","):
        for transposed in orders:
            if transport == "nccl":
                variants.append((transport, transposed, 1, 1.0, 1.0, "collective"))
            else:
                for ch in [int(c) for c in args.chunks.split(",")]:
                    for fr in (args.fracs.split(",") if ch > 1 else ["1:1"]):
                        fl, frm = [float(v) for v in fr.split(":")]
                        for bar in args.barriers.split(","):
                            variants.append((transport, transposed, ch, fl, frm, bar))
    for transport, transposed, chunks, fl, frm, bar in variants:
        if True:
            plan = DistFFT2D(w, h, np.float32, transport=transport, transposed_out=transposed, chunks=chunks,
                             frac_local=fl, frac_remote=frm, barrier=bar)
            for _ in range(args.warmup):
                plan.transform(slab)
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                plan.transform(slab)
            b.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            ntr = 1 if transposed else 2
            phases = None
            if args.phases and transport == "p2p":
                plan.start_phase_timing()
                for _ in range(args.steps):
                    plan.transform(slab)
                ph = plan.phase_times_ms()
                pt = torch.tensor(list(ph.values()), device="cuda", dtype=torch.float64)
                pmax = pt.clone()
                dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
                phases = {k: [round(float(a), 4), round(float(b), 4)] for k, a, b in zip(ph, pt.tolist(), pmax.tolist())}
            if rank == 0:
                print(json.dumps({
                    "workload": f"C5: 2D C2C fp32 {w}x{h}, slab-decomposed over {world} GPUs", "transport": transport,
                    "chunks": chunks, "frac_local": fl, "frac_remote": frm, "barrier": bar,
                    "chain": os.environ.get("GENFFT_CUDA_CHAIN", "1"),
                    "knobs": {k: v for k, v in os.environ.items() if k.startswith("GENFFT_CUDA_")},
                    "output": "transposed (1 global transpose)" if transposed else "natural order (2 global transposes)",
                    "phases_ms_rank0_and_max": phases, "n_gpus": world, "ms": ms, "gflops": flop / (ms * 1e-3) / 1e9,
                    "alltoall_bytes_sent_per_gpu_per_transpose": sent,
                    "nvlink_floor_ms_at_770GBs": ntr * sent / 770e9 * 1e3,
                    "nvlink_floor_ms_at_900GBs": ntr * sent / 900e9 * 1e3,
                    "frac_of_nvlink_roofline_770": (ntr * sent / 770e9 * 1e3) / ms,
                }), flush=True)
            plan.close()
            del plan
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
