#!/usr/bin/env python
"""Benchmark of the transform hot path (contract: one JSON line on stdout from rank 0).

Workload at every N: BASELINE.json configs[1] ("C2") -- batched 1D complex fp32 C2C, N = 4096 x 2^16
forward transforms per GPU (a "step" = one pass of the hot path over that batch).  With N > 1 GPUs the
batch is sharded: every rank transforms its own 2^16 sequences, no collective on the data path
("scaling": "weak"), value = transforms of all ranks / max-over-ranks device time.

  value      effective GFLOP/s = 5 N log2 N * batch / t, inputs resident in HBM, CUDA events
  roofline   algorithmic bytes (one read + one write of the batch) / kernel time vs measured HBM peak
  e2e        same metric through the host-pointer C-ABI call (genfft_cuda_exec_c2c) on pinned host
             buffers: H2D + kernels + D2H inside the timed region
  cpu_baseline / --impl reference
             genFFT's own CPU implementation (oracle/_ref, dispatch AVX2/FMA build) on the host cores
  extras.C5_dist (N > 1 only)
             BASELINE.json configs[4]: 2D C2C fp32 32768 x 32768 slab-decomposed over the N GPUs with the all-to-all
             fused into the FFT kernels' stores (genfft_b200.dist.DistFFT2D, transport "p2p"), natural order and
             transposed output, against the NVLink all-to-all roofline, with a parity figure against genFFT's CPU
             output taken on the same multi-process path (FFT2D::transform, fft.h:213-241)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout  # replaced by claim_stdout() in main()
N_FFT = 4096
BATCH = 1 << 16
METRIC = "effective GFLOP/s (5N*log2N/t), batched 1D C2C fp32 N=4096 x 2^16 per GPU"
UNIT = "GFLOP/s"
FLOP_PER_STEP = 5.0 * N_FFT * math.log2(N_FFT) * BATCH
ALGO_BYTES_PER_STEP = 2 * N_FFT * BATCH * 8  # one read + one write of the batch


def config(n_gpus: int) -> dict:
    return {
        "workload": "C2: batched 1D C2C fp32 forward, N=4096 x 2^16 transforms per GPU (BASELINE.json configs[1])",
        "n_fft": N_FFT,
        "batch_per_gpu": BATCH,
        "global_batch": BATCH * n_gpus,
        "parallelism": f"batch-sharded x{n_gpus}, no data-path collective",
        "l2": "inputs (2 GiB per GPU) larger than L2 (126 MB); no flush needed",
    }


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled from a
    background thread (the timed region is milliseconds long, shorter than nvidia-smi's sampling period)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[int, int]] = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), int(clk), int(rs)))
                    except Exception as e:  # keep the bench alive
                        self.err = repr(e)
                        return
                    time.sleep(0.0005)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        """Median SM clock and throttle reasons of the samples taken inside the host-time window [t0, t1] (the timed
        region; the thread is started before the warm-up so that NVML's first-call latency is not inside it)."""
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": [f"unavailable: {self.err}"]}
        inside = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)]
        if not inside:  # region shorter than one NVML round trip: the sample closest to it
            mid = 0.5 * ((t0 or 0.0) + (t1 or 0.0))
            inside = [min(self.samples, key=lambda s: abs(s[0] - mid))]
        clocks = sorted(c for _, c, _ in inside)
        seen = set()
        for _, _, rs in inside:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    seen.add(name)
        return {"sm_mhz": float(clocks[len(clocks) // 2]), "sm_max_mhz": self.max_mhz, "samples": len(clocks),
                "samples_total": len(self.samples), "reasons": sorted(seen)}


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    the box sets NCCL_DEBUG), so fd 1 is pointed at stderr for the run and the line goes to the saved real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def reference_arm(args, rank: int, world: int) -> int:
    """genFFT's own CPU implementation of the path on the host cores (oracle/_ref), same metric/config."""
    if rank != 0:
        return 0
    import oracle
    if not oracle.have_ref():
        if os.path.isdir(oracle.REFERENCE_ROOT):
            oracle.build("ref")
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgenfft_ref.so was not built"}), file=OUT,
              flush=True)
        return 0
    ref = oracle.Ref()
    threads = ref.hardware_threads()
    # One step = one sweep of the reference over a bounded sample of the batch ARRAY: `per_step` distinct transforms,
    # host array in -> host array out, split over all host threads (the reference has no batch API; this is the loop
    # a CPU caller writes, and the same host-buffer contract our e2e figure is timed on).
    per_step = min(BATCH * world, 16384)
    t = ref.bench_c2c_array(N_FFT, per_step, threads, args.warmup, args.steps)
    # ms_per_step is quoted for the batch the config names (2^16 transforms per GPU); the sweep timed is a bounded
    # sample of it, scaled by the transform count (the loop is linear in it: independent transforms, DRAM-streaming)
    sample_ms = 1e3 * t
    ms = sample_ms * (BATCH * world) / per_step
    value = 5.0 * N_FFT * math.log2(N_FFT) * per_step / t / 1e9
    # the reference's own benchmark loop (fft_bench.cpp FFT_1D: one in/out buffer, fresh input per iteration) keeps
    # the 64 KiB working set in cache; reported beside the streaming figure
    t_cal = ref.bench_c2c(N_FFT, 64 * threads, threads)
    n_res = max(threads, int(64 * threads * 3.0 / max(t_cal, 1e-6)))
    t_res = ref.bench_c2c(N_FFT, n_res, threads)
    resident = 5.0 * N_FFT * math.log2(N_FFT) * n_res / t_res / 1e9
    sample = (f"{per_step} of {BATCH * world} transforms per step as a host batch array (in and out each "
              f"{per_step * N_FFT * 8 >> 20} MiB, streamed from/to DRAM), {threads} host threads, forward only")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic U(-1,1), std::mt19937_64 per thread",
        "config": dict(config(args.gpus), reference_sample=sample),
        "sample_ms_per_step": sample_ms, "sample_transforms_per_step": per_step,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                         "cache_resident_value": resident,
                         "cache_resident_sample": f"{n_res} transforms through one in/out buffer per thread "
                                                  f"(restated fft_bench.cpp FFT_1D loop)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_build": ref.describe(),
    }
    print(json.dumps(line), file=OUT, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# extras.C5_dist: BASELINE.json configs[4] on the N GPUs of the run (N > 1).  It runs in CHILD processes (one per rank,
# their own process group) so that a failure or a hang there cannot take the contract line with it.
# ---------------------------------------------------------------------------------------------------------------------
# what an all-to-all of SM-issued peer stores with NO compute reaches on this pool's 8-GPU boxes, GB/s per direction and
# GPU (tools/alltoall_store_bench.cu, profiles/r02_alltoall_store_ceiling.log; 8-byte stores in 256-byte runs; for a
# pair the 1 GiB figure of tools/peer_store_bench.cu, profiles/r02_peer_store_shape_bench_2gpu.log)
A2A_STORE_CEILING_GBS = {2: 705.0, 4: 680.0, 8: 664.0}
C5_W = C5_H = 32768
C5_COLS = [0, 1, 15, 16, 4097, 16384, 20011, 32767]  # sampled output columns: both halves, tile edges, odd places


def c5_dist_child(args) -> int:
    """One rank of the C5 measurement: DistFFT2D(32768, 32768, p2p) natural order and transposed output, per-phase
    times, and parity against genFFT's own CPU output on the same multi-process path: (a) a 2048 x 2048 transform
    against FFT2D::transform of the compiled reference, (b) at full size, sampled output columns against the
    reference's row transforms of ALL rows (every rank transforms its own slab on the host cores) followed by its
    vertical transform of the sampled columns (fft.h:229-241 then :216-217 -- the reference's own two steps)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from genfft_b200.dist import DistFFT2D

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = {"workload": f"C5: 2D C2C fp32 {C5_W}x{C5_H}, row slabs over {world} GPUs (BASELINE.json configs[4])",
           "n_gpus": world, "transport": "p2p (all-to-all fused into the FFT kernels' stores over NVLink peer memory)",
           "steps": args.c5_steps, "warmup": 2}
    w, h = C5_W, C5_H
    hl, wp = h // world, w // world
    flop = 5.0 * w * h * math.log2(w * h)
    sent = (w * h * 8 / world) * (world - 1) / world  # bytes every GPU sends per global transpose
    out["alltoall_bytes_sent_per_gpu_per_transpose"] = sent

    def max_over_ranks(v):
        t = torch.tensor(v, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(v):
        t = torch.tensor(v, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    try:
        import oracle
        ref = oracle.Ref() if oracle.have_ref() else None
    except Exception:
        ref = None
    tol = 1e-6 * math.log2(w * h)
    parity = {"tolerance_rel_l2": tol, "comparand": "genFFT CPU (oracle/_ref, unmodified reference, AVX2/FMA dispatch)"
              if ref is not None else "unavailable: oracle/_ref was not built"}

    # ---- (a) 2048 x 2048 through the same classes against the reference's FFT2D::transform ----
    if ref is not None:
        sw = sh = 2048
        rng = np.random.default_rng(7)
        full = (rng.uniform(-1, 1, (sh, sw)) + 1j * rng.uniform(-1, 1, (sh, sw))).astype(np.complex64)
        want = ref.fft2d(full)
        shl = sh // world
        small = DistFFT2D(sw, sh, np.float32, transport="p2p")
        slab_s = torch.from_numpy(full[rank * shl:(rank + 1) * shl].copy()).cuda()
        for _ in range(2):  # twice: buffers and epoch flags are reused between calls
            got = small.transform(slab_s)
            torch.cuda.synchronize()
        d = got.cpu().numpy().astype(np.complex128) - want[rank * shl:(rank + 1) * shl]
        num, den = sum_over_ranks([float(np.vdot(d, d).real), float(np.vdot(want[rank * shl:(rank + 1) * shl],
                                                                                want[rank * shl:(rank + 1) * shl]).real)])
        parity["small_2048x2048_rel_l2"] = math.sqrt(num / den)
        dist.barrier()
        small.close()
        del small, slab_s, got

    # ---- the C5 slab of this rank ----
    gen = torch.Generator(device="cuda").manual_seed(1000 + rank)
    slab = torch.view_as_complex(torch.rand((hl, w, 2), generator=gen, device="cuda") * 2 - 1)

    def measure(transposed):
        plan = DistFFT2D(w, h, np.float32, transport="p2p", transposed_out=transposed)
        for _ in range(2):
            res = plan.transform(slab)
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.c5_steps):
            res = plan.transform(slab)
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = max_over_ranks([a.elapsed_time(b) / args.c5_steps])[0]
        plan.start_phase_timing()
        for _ in range(args.c5_steps):
            res = plan.transform(slab)
        ph = plan.phase_times_ms()
        pmax = max_over_ranks(list(ph.values()))
        ntr = 1 if transposed else 2
        floor900 = ntr * sent / 900e9 * 1e3
        r = {"ms": ms, "gflops": flop / (ms * 1e-3) / 1e9, "global_transposes": ntr,
             "nvlink_floor_ms_at_900GBs": floor900, "frac_of_nvlink_900": floor900 / ms,
             "frac_of_nvlink_770_measured_peer_copy": (ntr * sent / 770e9 * 1e3) / ms,
             "frac_of_measured_alltoall_store_ceiling": (ntr * sent / (A2A_STORE_CEILING_GBS[world] * 1e9) * 1e3) / ms
             if world in A2A_STORE_CEILING_GBS else None,
             "alltoall_store_ceiling_source": "constant from profiles/r02_alltoall_store_ceiling.log (a store-only "
                                              "all-to-all kernel, no FFT); NOT measured by this run",
             "phases_ms_max_over_ranks": {k: round(v, 4) for k, v in zip(ph, pmax)}}
        return plan, res, r

    plan, res, out["natural_order"] = measure(False)

    # ---- (b) full-size parity on the natural-order result: sampled columns against the reference's two steps ----
    if ref is not None:
        threads = max(1, ref.hardware_threads() // world)
        rows_s = np.empty((hl, len(C5_COLS)), np.complex64)
        chunk = 1024
        for r0 in range(0, hl, chunk):
            rows = ref.c2c_rows(slab[r0:r0 + chunk].cpu().numpy(), False, threads)
            rows_s[r0:r0 + chunk] = rows[:, C5_COLS]
        gathered = [None] * world
        dist.all_gather_object(gathered, rows_s)
        want_cols = ref.vert(np.concatenate(gathered, axis=0))[rank * hl:(rank + 1) * hl]
        got_cols = res[:, C5_COLS].cpu().numpy()
        d = got_cols.astype(np.complex128) - want_cols
        num, den = sum_over_ranks([float(np.vdot(d, d).real), float(np.vdot(want_cols, want_cols).real)])
        parity["full_size_sampled_columns_rel_l2"] = math.sqrt(num / den)
        parity["full_size_sampled_columns"] = C5_COLS
        parity["ok"] = bool(parity["full_size_sampled_columns_rel_l2"] <= tol and
                            parity.get("small_2048x2048_rel_l2", 0.0) <= 1e-6 * 22)
    dist.barrier()
    plan.close()
    del plan, res
    torch.cuda.empty_cache()
    plan, res, out["transposed_output"] = measure(True)
    dist.barrier()
    plan.close()
    del plan, res, slab
    torch.cuda.empty_cache()
    out["parity"] = parity
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        return 0
    # the same transform on ONE GPU (rank 0 alone), so that strong scaling can be read off this line
    try:
        import genfft_b200 as g
        p1 = g.FFT2D(w, h, np.float32)
        x = torch.view_as_complex(torch.rand((h, w, 2), device="cuda") * 2 - 1)
        y = torch.empty_like(x)
        for _ in range(2):
            p1.transform(y, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            p1.transform(y, x)
        b.record()
        b.synchronize()
        ms1 = a.elapsed_time(b) / 3
        out["one_gpu_ms"] = ms1
        out["strong_scaling_speedup_natural_order"] = ms1 / out["natural_order"]["ms"]
        out["strong_scaling_efficiency_natural_order"] = ms1 / out["natural_order"]["ms"] / world
    except Exception as e:
        out["one_gpu_error"] = repr(e)
    print(json.dumps(out), file=OUT, flush=True)
    return 0


def run_c5_dist_children(torch, dist, rank: int, world: int, local_rank: int, steps: int, timeout_s: float = 420.0):
    """Every rank of the contract run launches its child; rank 0 returns the child's JSON (or an error record)."""
    import socket
    port = [0]
    if rank == 0:
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port[0] = sk.getsockname()[1]
    dist.broadcast_object_list(port, src=0)
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local_rank),
               MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port[0]))
    for k in ("TORCHELASTIC_RUN_ID", "TORCHELASTIC_USE_AGENT_STORE", "GROUP_RANK", "ROLE_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--c5-dist-child", "--c5-steps", str(steps)]
    t0 = time.perf_counter()
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        so, se = proc.communicate(timeout=timeout_s)
        rc = proc.returncode
    except subprocess.TimeoutExpired:
        proc.kill()
        so, se = proc.communicate()
        rc = -9
    if rank != 0:
        return None
    rec = None
    for ln in reversed(so.strip().splitlines()):
        if ln.startswith("{"):
            try:
                rec = json.loads(ln)
                break
            except Exception:
                pass
    if rec is None:
        rec = {"error": f"child rc={rc}", "stderr_tail": se[-1500:]}
    rec["child_wall_s"] = time.perf_counter() - t0
    return rec


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--c5-dist-child", action="store_true", help="internal: one rank of the extras.C5_dist measurement")
    ap.add_argument("--c5-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    global OUT
    OUT = claim_stdout()
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    if args.c5_dist_child:
        return c5_dist_child(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import genfft_b200 as g

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: genfft_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    plan = g.FFT(N_FFT, np.float32, batch=BATCH)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.rand((BATCH, N_FFT, 2), generator=gen, device="cuda") * 2 - 1)
    y = torch.empty_like(x)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: NVML's first calls are slow and must not eat the timed region
    for _ in range(args.warmup):
        plan.forward(y, x)
    barrier()
    launches0 = g.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_host0 = time.perf_counter()
    ev[0].record()
    for i in range(args.steps):
        plan.forward(y, x)
        ev[i + 1].record()
    barrier()
    t_host1 = time.perf_counter()
    launches = g.launch_count() - launches0
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = FLOP_PER_STEP * world / (ms_per_step * 1e-3) / 1e9

    # sanity: the timed output is a transform of the input (Parseval), so no work was skipped
    ex = float((x[:64].abs().double() ** 2).sum())
    ey = float((y[:64].abs().double() ** 2).sum())
    assert abs(ey / (N_FFT * ex) - 1) < 1e-4, "output of the timed region is not the transform of the input"

    # ---- e2e: host-pointer C-ABI call on pinned host buffers, H2D + D2H inside the timed region ----
    # Two placements of the caller's buffers are timed: torch's pinned allocator (wherever the process happens to run)
    # and genfft_cuda_host_alloc (page-locked memory bound to the GPU's NUMA node).  On a two-socket 8-GPU host the
    # second keeps every link's traffic off the socket interconnect; the better one is the headline, both are listed.
    def e2e_time(hx_t, hy_t):
        plan.forward(hy_t, hx_t)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            plan.forward(hy_t, hx_t)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            tt = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
        return sec

    placements = {}
    hx = torch.empty((BATCH, N_FFT), dtype=torch.complex64, pin_memory=True)
    hy = torch.empty((BATCH, N_FFT), dtype=torch.complex64, pin_memory=True)
    hx.copy_(x)
    placements["torch_pinned_default"] = {"ms_per_step": 1e3 * e2e_time(hx, hy)}
    err = float((hy[:8].cuda() - y[:8]).abs().max())
    assert err < 1e-3, f"host-pointer path disagrees with the device path ({err})"
    del hx, hy
    try:
        from genfft_b200.hostmem import PinnedNearGpu
        bx = PinnedNearGpu((BATCH, N_FFT), np.complex64)
        by = PinnedNearGpu((BATCH, N_FFT), np.complex64)
        hx, hy = bx.tensor(), by.tensor()
        hx.copy_(x)
        placements["numa_local_host_alloc"] = {"ms_per_step": 1e3 * e2e_time(hx, hy), "gpu_numa_node": bx.numa_node}
        err = float((hy[:8].cuda() - y[:8]).abs().max())
        assert err < 1e-3, f"host-pointer path disagrees with the device path ({err})"
        del hx, hy
        bx.close()
        by.close()
    except AssertionError:
        raise
    except Exception as e:  # the placement helper is best effort; every rank takes the same branch on one host
        placements["numa_local_host_alloc"] = {"error": repr(e)}
    ok = {k: v for k, v in placements.items() if "ms_per_step" in v}
    best = min(ok, key=lambda k: ok[k]["ms_per_step"])
    e2e_s = ok[best]["ms_per_step"] * 1e-3
    e2e_value = FLOP_PER_STEP * world / e2e_s / 1e9
    hx = hy = None

    # ---- extras.C5_dist: the slab-decomposed 2D transform on the N GPUs (every rank launches its child) ----
    c5_dist = None
    if world > 1 and not args.no_extras:
        del x, y, hx, hy, plan
        torch.cuda.empty_cache()
        barrier()
        try:
            c5_dist = run_c5_dist_children(torch, dist, rank, world, local_rank, args.c5_steps)
        except Exception as e:  # never in the way of the contract line
            c5_dist = {"error": repr(e)}
        plan = g.FFT(N_FFT, np.float32, batch=BATCH)  # for the description below

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    kernel_ms = sum(per_kernel_ms) / len(per_kernel_ms)
    achieved = ALGO_BYTES_PER_STEP / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_source = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "c2_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_source = ("constant read from profiles/c2_traffic.json (one `ncu --set full` capture of this kernel, "
                          f"{tj.get('source', 'see profiles/')}); NOT measured by this run")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic U(-1,1) generated on device, seeded per rank", "config": config(world),
        "gpu_launches": int(launches), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_FFT * BATCH * 8,
                "d2h_bytes_per_step": N_FFT * BATCH * 8, "ms_per_step": e2e_s * 1e3,
                "host_buffer_placement": best, "placements": placements,
                "api": "genfft_cuda_exec_c2c (host pointers, pinned; chunked H2D/compute/D2H overlap on two streams)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_source, "kernel": "fft_tile_kernel<float,4096,16,1,M_ROWTMA,false> (cp.async.bulk prefetch)",
                     "kernel_ms": kernel_ms, "algorithmic_bytes": ALGO_BYTES_PER_STEP, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "plan": plan.describe(),
    }

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ----
    if world == 1:
        try:
            import oracle
            ref = oracle.Ref()
            threads = ref.hardware_threads()
            # bounded sample of the same workload: sweeps over a 16384-transform host batch array (in -> out, DRAM-
            # streaming like the e2e path), ~10 s of CPU work; beside it the reference's own cache-resident bench loop
            count = 16384
            t_one_sweep = ref.bench_c2c_array(N_FFT, count, threads, 1, 2)
            reps = max(2, min(2000, int(10.0 / max(t_one_sweep, 1e-4))))
            t_cpu = ref.bench_c2c_array(N_FFT, count, threads, 1, reps)
            t_cal = ref.bench_c2c(N_FFT, 32 * threads, threads)
            n_res = max(threads, int(32 * threads * 3.0 / max(t_cal, 1e-6)))
            t_res = ref.bench_c2c(N_FFT, n_res, threads)
            t_one = ref.bench_c2c(N_FFT, 2048, 1)
            line["cpu_baseline"] = {
                "value": 5.0 * N_FFT * 12 * count / t_cpu / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{reps} sweeps over a host batch array of {count} transforms ({count / BATCH:.2f} of the batch; "
                          f"in and out each {count * N_FFT * 8 >> 20} MiB, streamed from/to DRAM; ~{reps * t_cpu:.0f} s of "
                          f"CPU work), {threads} threads, forward only; {ref.describe()}",
                "cache_resident_value": 5.0 * N_FFT * 12 * n_res / t_res / 1e9,
                "cache_resident_sample": f"{n_res} transforms through one in/out buffer per thread, fresh U(-1,1) input "
                                         f"per transform (restated fft_bench.cpp FFT_1D loop)",
                "single_core_value": 5.0 * N_FFT * 12 * 2048 / t_one / 1e9,
            }
        except Exception as e:  # the oracle is a checker; its absence must not hide the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": f"unavailable: {e}"}
        if not args.no_extras:
            line["extras"] = extras(g, np, torch)
            try:
                line["extras"]["cpu_reference_1_thread"] = cpu_reference_extras(np)
            except Exception as e:  # reported beside the GPU numbers; never in their way
                line["extras"]["cpu_reference_1_thread"] = {"error": repr(e)}
    if c5_dist is not None:
        line.setdefault("extras", {})["C5_dist"] = c5_dist
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_reference_extras(np) -> dict:
    """genFFT's own CPU implementation (oracle/_ref, one thread -- the library is single-threaded) on the secondary
    configs, timed like test/fft_bench.cpp (std::chrono around the transform calls, fresh input per call): C1 exactly,
    C3 through the factory hook that lifts the reference past its 2^23 size switch, C4 on two of its 256 transforms,
    C5 on an 8192^2 image scaled by the 5 N log2 N work ratio."""
    import oracle
    ref = oracle.Ref()
    out = {}
    t = ref.bench_c2c(1024, 20000, 1, np.float32, fwd_only=False) / 20000
    out["C1_1d_c2c_f32_n1024_fwd_inv"] = {"us_per_pair": t * 1e6, "gflops": 2 * 5 * 1024 * 10 / t / 1e9}
    t = ref.bench_c2c(1 << 24, 1, 1, np.float64)
    out["C3_1d_c2c_f64_n2^24"] = {"ms": t * 1e3, "gflops": 5 * (1 << 24) * 24 / t / 1e9}
    t = ref.bench_r2c(1 << 22, 2, 1) / 2
    out["C4_1d_r2c_f32_n2^22_x256"] = {"ms_per_transform": t * 1e3, "ms_batch_of_256_extrapolated": t * 256 * 1e3,
                                       "gflops_2.5NlogN": 2.5 * (1 << 22) * 22 / t / 1e9}
    t = ref.bench_fft2d(8192, 8192, 1)
    scale = (32768.0 ** 2 * 30) / (8192.0 ** 2 * 26)
    out["C5_2d_c2c_f32_32768^2"] = {"ms_8192^2_measured": t * 1e3, "ms_32768^2_extrapolated": t * scale * 1e3,
                                    "gflops_at_8192^2": 5 * 8192.0 ** 2 * 26 / t / 1e9}
    return out


def extras(g, np, torch) -> dict:
    """Secondary single-GPU configs of BASELINE.json (reported, not the contract metric)."""
    out = {}

    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / iters

    peak, _ = measured_peaks()
    try:
        # C1: single N=1024 fp32 forward+inverse (latency): through the Python mirror, through raw pointers, and as
        # one CUDA graph of the two launches
        p = g.FFT(1024, np.float32)
        x = torch.randn(1024, dtype=torch.complex64, device="cuda")
        y, z = torch.empty_like(x), torch.empty_like(x)
        ms = timed(lambda: (p.forward(y, x), p.inverse(z, y)), iters=200, warm=20)
        c1 = {"us_per_pair": ms * 1e3, "gflops": 2 * 5 * 1024 * 10 / (ms * 1e-3) / 1e9}
        lib, h, st = g.lib(), p._h, torch.cuda.current_stream().cuda_stream
        px, py, pz = x.data_ptr(), y.data_ptr(), z.data_ptr()
        ms = timed(lambda: (lib.genfft_cuda_exec_c2c_dev(h, py, px, 0, st), lib.genfft_cuda_exec_c2c_dev(h, pz, py, 1, st)),
                   iters=500, warm=50)
        c1["us_per_pair_c_abi_raw_pointers"] = ms * 1e3
        # the same pair issued from a C loop inside the library (no Python between the launches): what a C/C++ caller of
        # the device-pointer path pays; host wall-clock over 2000 pairs including the final synchronisation
        import ctypes
        us = ctypes.c_double(0.0)
        if lib.genfft_cuda_debug_time_c2c_pairs(h, pz, py, px, 2000, st, ctypes.byref(us)) == 0:
            c1["us_per_pair_c_loop"] = us.value
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                p.forward(y, x)
                p.inverse(z, y)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                p.forward(y, x)
                p.inverse(z, y)
            ms = timed(graph.replay, iters=500, warm=50)
            c1["us_per_pair_cuda_graph"] = ms * 1e3
            assert float((z / 1024 - x).abs().max()) < 1e-4
        except Exception as e:
            c1["cuda_graph_error"] = repr(e)
        out["C1_1d_c2c_f32_n1024_fwd_inv"] = c1
        # C3: N=2^24 fp64
        n = 1 << 24
        p = g.FFT(n, np.float64)
        x = torch.randn(n, dtype=torch.complex128, device="cuda")
        y = torch.empty_like(x)
        ms = timed(lambda: p.forward(y, x))
        out["C3_1d_c2c_f64_n2^24"] = {"ms": ms, "gflops": 5 * n * 24 / (ms * 1e-3) / 1e9,
                                      "algorithmic_gbs": 2 * n * 16 / (ms * 1e-3) / 1e9,
                                      "frac_of_measured_hbm": 2 * n * 16 / (ms * 1e-3) / 1e9 / peak, "plan": p.describe()}
        del x, y
        # C4: R2C fp32 N=2^22 x 256 (half spectrum)
        n, b = 1 << 22, 256
        p = g.RealFFT(n, np.float32, half=True, batch=b)
        x = torch.randn(b, n, device="cuda")
        y = torch.empty((b, n // 2 + 1), dtype=torch.complex64, device="cuda")
        ms = timed(lambda: p.forward(y, x), iters=5, warm=2)
        by = (n * 4 + (n // 2 + 1) * 8) * b
        out["C4_1d_r2c_f32_n2^22_x256"] = {"ms": ms, "gflops_2.5NlogN": 2.5 * n * 22 * b / (ms * 1e-3) / 1e9,
                                           "gflops_5NlogN": 5 * n * 22 * b / (ms * 1e-3) / 1e9,
                                           "algorithmic_gbs": by / (ms * 1e-3) / 1e9,
                                           "frac_of_measured_hbm": by / (ms * 1e-3) / 1e9 / peak, "plan": p.describe()}
        del x, y
        # C5 on ONE GPU (the multi-GPU slab version is bench_dist.py): 2D fp32 32768^2
        w = h = 32768
        p = g.FFT2D(w, h, np.float32)
        x = torch.randn(h, w, dtype=torch.complex64, device="cuda")
        y = torch.empty_like(x)
        ms = timed(lambda: p.transform(y, x), iters=3, warm=1)
        out["C5_2d_c2c_f32_32768^2_1gpu"] = {"ms": ms, "gflops": 5 * w * h * 30 / (ms * 1e-3) / 1e9,
                                             "algorithmic_gbs": 2 * w * h * 8 / (ms * 1e-3) / 1e9,
                                             "frac_of_measured_hbm": 2 * w * h * 8 / (ms * 1e-3) / 1e9 / peak,
                                             "plan": p.describe()}
    except Exception as e:
        out["error"] = repr(e)
    return out


if __name__ == "__main__":
    sys.exit(main())
