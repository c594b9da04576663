"""Multi-GPU layer: one process per GPU over ``torch.distributed`` (NCCL on NVLink 5 / NVSwitch).

The reference has no multi-anything; this is the part of north_star that is new:

* **Batched 1D** (BASELINE configs C2, C4) shards trivially: ``shard_batch`` gives every rank a contiguous
  range of transforms, there is no collective on the data path.
* **2D** (config C5) uses a slab decomposition, rank r owning rows ``[r*H/P, (r+1)*H/P)``:

  1. local row FFTs (length W);
  2. global transpose, after which rank r holds columns ``[r*W/P, (r+1)*W/P)`` of every row;
  3. local column FFTs (length H) on that (H x W/P) block;
  4. for natural-order output (genFFT semantics) the transpose back to row slabs.

  Two transports exist for steps 2 and 4:

  ``"p2p"`` (the product): the FFT kernel's store *is* the all-to-all -- the last pass of step 1 (3) writes
  each output element straight into the destination rank's receive buffer through CUDA-IPC mapped peer
  pointers (NVLink stores), so the transfer overlaps the butterflies tile by tile and no pack / unpack pass
  or send buffer exists.  Ranks only exchange a stream-ordered barrier.

  ``"nccl"`` (the baseline): the kernel packs per-destination blocks, ``all_to_all_single`` moves them, and a
  strided copy unpacks -- what a library-only solution does.

The orchestration is written against a small local-engine interface so that its index arithmetic and
collective wiring can be tested on CPU with the gloo backend (tests/test_dist_cpu.py injects a CPU engine;
the product engine below is CUDA-only and has no fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_batch(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) range of transforms owned by `rank` (independent units, no collective)."""
    return batch * rank // world, batch * (rank + 1) // world


def _is_pow2(n: int) -> bool:
    return n >= 1 and (n & (n - 1)) == 0


class CudaSlabEngine:
    """Local passes of the slab decomposition on the GPU (libgenfft_cuda dist_rows / dist_cols plans).

    With ``chunks > 1`` the slab is cut into row (column) chunks that alternate between two streams, each with its
    own plan (own scratch) launched on a fraction of the SMs, so that the NVLink-bound remote-store pass of chunk k
    overlaps the HBM-bound local pass of chunk k+1."""

    def __init__(self, width: int, height: int, world: int, dtype, chunks: int = 1, frac_local: float = 0.7,
                 frac_remote: float = 0.3):
        from .api import _precision
        self.w, self.h, self.p = width, height, world
        self.hl, self.wp = height // world, width // world
        self.precision = _precision(dtype)
        self.cdtype = torch.complex64 if self.precision == _lib.F32 else torch.complex128
        self.esize = 8 if self.precision == _lib.F32 else 16
        self._rows = C.c_void_p()
        self._cols = C.c_void_p()
        check(lib().genfft_cuda_plan_dist_rows(C.byref(self._rows), self.precision, width, self.hl, world))
        check(lib().genfft_cuda_plan_dist_cols(C.byref(self._cols), self.precision, height, self.wp, world))
        while chunks > 1 and (self.hl % chunks or self.wp % chunks or self.wp // chunks < 32):
            chunks //= 2
        self.chunks = max(1, chunks)
        self._chunk_plans = []
        if self.chunks > 1:
            self.streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            for _ in range(2):
                r, c = C.c_void_p(), C.c_void_p()
                check(lib().genfft_cuda_plan_dist_rows(C.byref(r), self.precision, width, self.hl // self.chunks, world))
                check(lib().genfft_cuda_plan_dist_cols(C.byref(c), self.precision, height, self.wp // self.chunks, world))
                for h in (r, c):
                    check(lib().genfft_cuda_plan_set_grid_fraction(h, frac_local, frac_remote))
                self._chunk_plans.append((r, c))

    def __del__(self):
        try:
            plain = getattr(self, "_rows_plain", None)
            for h in [self._rows, self._cols, plain] + [x for pair in self._chunk_plans for x in pair]:
                if h:
                    lib().genfft_cuda_plan_destroy(h)
        except Exception:
            pass

    def _fan_out(self, launch):
        """Runs launch(k, plan_pair) for every chunk on alternating side streams, fenced against the current stream."""
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        for s in self.streams:
            s.wait_event(start)
        for k in range(self.chunks):
            with torch.cuda.stream(self.streams[k % 2]):
                launch(k, self._chunk_plans[k % 2])
        for s in self.streams:
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.cdtype, device="cuda")

    # -- nccl transport --------------------------------------------------------------------------
    def rows_pack(self, slab, send, inv: bool):
        """send[g, r, :] = row FFT of slab[r] restricted to columns of rank g."""
        check(lib().genfft_cuda_exec_dist_rows_dev(self._rows, send.data_ptr(), None, self.hl * self.wp, 0,
                                                   slab.data_ptr(), self.w, int(inv), self._stream()))

    def cols(self, out, block, inv: bool):
        """out = FFT along axis 0 of block (H x W/P), both dense."""
        check(lib().genfft_cuda_exec_dist_cols_dev(self._cols, out.data_ptr(), None, self.wp, 0, block.data_ptr(),
                                                   self.wp, int(inv), self._stream()))

    def unpack(self, out, recv):
        """out[r, g*Wp + x] = recv[g, r, x]."""
        check(lib().genfft_cuda_copy2d_dev(self.precision, out.data_ptr(), self.w, self.wp, recv.data_ptr(), self.wp,
                                           self.hl * self.wp, self.hl, self.wp, self.p, self._stream()))

    # -- p2p transport: the store is the all-to-all -------------------------------------------------
    def rows_to_peers(self, slab, peer_ptrs, rank: int, inv: bool):
        arr = (C.c_void_p * self.p)(*peer_ptrs)
        if self.chunks == 1:
            check(lib().genfft_cuda_exec_dist_rows_dev(self._rows, None, arr, 0, rank * self.hl, slab.data_ptr(),
                                                       self.w, int(inv), self._stream()))
            return
        rk = self.hl // self.chunks
        base = slab.data_ptr()

        def launch(k, plans):
            check(lib().genfft_cuda_exec_dist_rows_dev(plans[0], None, arr, 0, rank * self.hl + k * rk,
                                                       base + k * rk * self.w * self.esize, self.w, int(inv),
                                                       self._stream()))
        self._fan_out(launch)

    def cols_to_peers(self, block_ptr: int, peer_ptrs, rank: int, inv: bool, twiddle_n: int = 0) -> bool:
        """Column transforms of this rank's (H x W/P) block, the rows scattered to the peers' row slabs.  With
        ``twiddle_n`` (distributed four-step 1D) the factor W_n^(kr*c) is fused into those stores; returns whether it
        was (False: the caller applies ``twiddle`` on the receiving slab)."""
        arr = (C.c_void_p * self.p)(*peer_ptrs)
        if self.chunks == 1:
            if twiddle_n:
                fused = C.c_int(0)
                check(lib().genfft_cuda_exec_dist_cols_tw_dev(self._cols, arr, self.w, rank * self.wp, block_ptr, self.wp,
                                                              int(inv), twiddle_n, C.byref(fused), self._stream()))
                return bool(fused.value)
            check(lib().genfft_cuda_exec_dist_cols_dev(self._cols, None, arr, self.w, rank * self.wp, block_ptr,
                                                       self.wp, int(inv), self._stream()))
            return False
        ck = self.wp // self.chunks

        def launch(k, plans):
            check(lib().genfft_cuda_exec_dist_cols_dev(plans[1], None, arr, self.w, rank * self.wp + k * ck,
                                                       block_ptr + k * ck * self.esize, self.wp, int(inv),
                                                       self._stream()))
        self._fan_out(launch)
        return False

    def cols_ptr(self, out, block_ptr: int, inv: bool):
        check(lib().genfft_cuda_exec_dist_cols_dev(self._cols, out.data_ptr(), None, self.wp, 0, block_ptr, self.wp,
                                                   int(inv), self._stream()))


    # -- distributed four-step 1D (DistFFT1D): copies, twiddle, plain row transforms, final transpose ----------
    def pack_cols(self, send, slab):
        """send[g, r, :] = slab[r, g*Wp:(g+1)*Wp] (per-destination column blocks of a row slab)."""
        check(lib().genfft_cuda_copy2d_dev(self.precision, send.data_ptr(), self.wp, self.hl * self.wp, slab.data_ptr(),
                                           self.w, self.wp, self.hl, self.wp, self.p, self._stream()))

    def cols_blocks_to_peers(self, slab, peer_ptrs, rank: int):
        """slab[r, g*Wp:(g+1)*Wp] stored straight into peer g's (H x W/P) block buffer at rows [rank*H/P, ...): the first
        global transpose of the four-step 1D transform, one launch with 16-byte accesses."""
        arr = (C.c_void_p * self.p)(*peer_ptrs)
        if slab.data_ptr() % 16 == 0 and all(q % 16 == 0 for q in peer_ptrs) and self.wp >= 2:
            check(lib().genfft_cuda_scatter_cols_dev(self.precision, arr, self.p, rank * self.hl, slab.data_ptr(), self.w,
                                                     self.hl, self.w, self._stream()))
            return
        for g in range(self.p):
            check(lib().genfft_cuda_copy2d_dev(self.precision, peer_ptrs[g] + rank * self.hl * self.wp * self.esize,
                                               self.wp, 0, slab.data_ptr() + g * self.wp * self.esize, self.w, 0,
                                               self.hl, self.wp, 1, self._stream()))

    def twiddle(self, slab, row0: int, inv: bool):
        """slab[r, c] *= W_N^((row0 + r) * c), N = H*W (conjugated for the inverse), in place."""
        check(lib().genfft_cuda_twiddle2d_dev(self.precision, slab.data_ptr(), self.w, self.hl, self.w, row0,
                                              self.w * self.h, int(inv), self._stream()))

    def rows(self, out, slab, inv: bool):
        """out[r] = FFT of slab[r] (H/P contiguous rows of length W)."""
        if getattr(self, "_rows_plain", None) is None:
            self._rows_plain = C.c_void_p()
            check(lib().genfft_cuda_plan_c2c_1d(C.byref(self._rows_plain), self.precision, self.w, self.hl, 0, 0))
        check(lib().genfft_cuda_exec_c2c_dev(self._rows_plain, out.data_ptr(), slab.data_ptr(), int(inv), self._stream()))

    def transpose(self, out, block):
        """out (W/P x H) = block (H x W/P) transposed."""
        check(lib().genfft_cuda_transpose_dev(self.precision, out.data_ptr(), self.h, block.data_ptr(), self.wp, self.h,
                                              self.wp, self._stream()))


class _PtrView:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerBuffers:
    """A device buffer per rank, allocated by libgenfft_cuda and mapped into every peer through CUDA IPC."""

    def __init__(self, nbytes: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = nbytes
        p = C.c_void_p()
        check(lib().genfft_cuda_malloc(C.byref(p), nbytes))
        self.local = p.value
        handle = C.create_string_buffer(64)
        check(lib().genfft_cuda_ipc_get_handle(self.local, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local)
            else:
                q = C.c_void_p()
                check(lib().genfft_cuda_ipc_open_handle(C.byref(q), h))
                self.ptrs.append(q.value)
                self._opened.append(q.value)

    def tensor(self, shape, cdtype):
        typestr = "<c8" if cdtype == torch.complex64 else "<c16"
        return torch.as_tensor(_PtrView(self.local, shape, typestr), device="cuda")

    def close(self):
        for q in self._opened:
            lib().genfft_cuda_ipc_close_handle(q)
        self._opened = []
        if self.local:
            torch.cuda.synchronize()
            lib().genfft_cuda_free(self.local)
            self.local = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistFFT2D:
    """Slab-decomposed genfft::FFT2D<T>(width, height) (fft.h:198-245) over a process group.

    ``transform(in_slab, inv)``: `in_slab` is this rank's (H/P x W) row slab.  Returns this rank's slab of the
    natural-order result (H/P x W), or with ``transposed_out=True`` its (H x W/P) column block (rows = ky,
    columns = this rank's kx range), which saves the second global transpose.
    """

    def __init__(self, width: int, height: int, dtype=np.float32, group=None, transport: str = "p2p",
                 transposed_out: bool = False, engine=None, chunks: int = 1, frac_local: float = 0.7,
                 frac_remote: float = 0.3, barrier: str = "flags"):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if not (_is_pow2(width) and _is_pow2(height) and _is_pow2(self.world)):
            raise ValueError("width, height and the number of ranks must be powers of two")
        if width % self.world or height % self.world or width < 2 or height < 2:
            raise ValueError("width and height must be divisible by the number of ranks")
        if transport not in ("p2p", "nccl"):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        if barrier not in ("flags", "collective"):
            raise ValueError("barrier must be 'flags' or 'collective'")
        self.w, self.h = width, height
        self.hl, self.wp = height // self.world, width // self.world
        self.transport = transport
        self.transposed_out = transposed_out
        self.engine = engine if engine is not None else CudaSlabEngine(
            width, height, self.world, dtype, chunks if transport == "p2p" else 1, frac_local, frac_remote)
        e = self.engine
        if transport == "nccl":
            self.send = e.empty(self.world, self.hl, self.wp)
            self.block = e.empty(self.h, self.wp)       # receive buffer of transpose 1 == (H x W/P) block
            self.block_out = e.empty(self.h, self.wp)
            self.recv2 = None if transposed_out else e.empty(self.world, self.hl, self.wp)
            self.out = None if transposed_out else e.empty(self.hl, self.w)
        else:
            esz = 8 if e.cdtype == torch.complex64 else 16
            self.block_buf = PeerBuffers(self.h * self.wp * esz, group)
            self.block_out = e.empty(self.h, self.wp) if transposed_out else None
            self.final_buf = None if transposed_out else PeerBuffers(self.hl * self.w * esz, group)
            self.out = None if transposed_out else self.final_buf.tensor((self.hl, self.w), e.cdtype)
            self._token = torch.zeros(1, device="cuda")
            # barrier between the passes: epoch flags in IPC-mapped peer memory (one tiny kernel, no collective call);
            # "collective" keeps the all-reduce of a token (also what an injected CPU engine gets)
            self._flags = None
            if barrier == "flags" and engine is None:
                self._flags = PeerBuffers(256, group)
                check(lib().genfft_cuda_memset_dev(self._flags.local, 0, 256))
                self._flag_ptrs = (C.c_void_p * self.world)(*self._flags.ptrs)
                self._epoch = 0
                torch.cuda.synchronize()
                dist.barrier(group=self.group)  # every rank's flags are zeroed before anyone publishes an epoch

    def _stream_barrier(self):
        # stream-ordered cross-rank barrier: a rank leaves it only after every rank's earlier kernels on
        # this stream -- whose NVLink stores target our buffers -- have completed
        if getattr(self, "_flags", None) is not None:
            self._epoch += 1
            check(lib().genfft_cuda_peer_barrier_dev(self._flag_ptrs, self.rank, self.world, self._epoch & 0xFFFFFFFF,
                                                     torch.cuda.current_stream().cuda_stream))
            return
        dist.all_reduce(self._token, group=self.group)

    def transform(self, in_slab, inv: bool = False):
        e = self.engine
        if tuple(in_slab.shape) != (self.hl, self.w):
            raise ValueError(f"expected this rank's ({self.hl} x {self.w}) row slab")
        if self.transport == "nccl":
            e.rows_pack(in_slab, self.send, inv)
            dist.all_to_all_single(self.block.view(self.world, self.hl, self.wp), self.send, group=self.group)
            e.cols(self.block_out, self.block, inv)
            if self.transposed_out:
                return self.block_out
            dist.all_to_all_single(self.recv2, self.block_out.view(self.world, self.hl, self.wp), group=self.group)
            e.unpack(self.out, self.recv2)
            return self.out
        # p2p: stores go straight into the peers' buffers
        mark = self._mark
        mark()
        # Peers must have finished reading their block buffers of the previous call before this call's row pass stores
        # into them.  In natural order the previous call ended with a barrier that every rank reached AFTER its column
        # pass (the only reader of its block buffer), so that barrier already orders it; with transposed output the
        # call ends right after the column pass and the entry barrier is needed.
        if self.transposed_out:
            self._stream_barrier()
        mark()
        e.rows_to_peers(in_slab, self.block_buf.ptrs, self.rank, inv)
        mark()
        self._stream_barrier()
        mark()
        if self.transposed_out:
            e.cols_ptr(self.block_out, self.block_buf.local, inv)
            mark()
            return self.block_out
        e.cols_to_peers(self.block_buf.local, self.final_buf.ptrs, self.rank, inv)
        mark()
        self._stream_barrier()
        mark()
        return self.out

    # optional per-phase timing of the p2p path (bench_dist.py --phases): CUDA events between the phases
    phase_names = ("barrier0", "rows+transpose1", "barrier1", "cols+transpose2", "barrier2")
    _events = None

    def start_phase_timing(self):
        self._events = []

    def _mark(self):
        if self._events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events.append(ev)

    def phase_times_ms(self):
        """mean milliseconds per phase over the transforms recorded since start_phase_timing()"""
        torch.cuda.synchronize()
        ev, self._events = self._events, None
        per = 5 if self.transposed_out else 6
        n = len(ev) // per
        out = [0.0] * (per - 1)
        for k in range(n):
            for j in range(per - 1):
                out[j] += ev[k * per + j].elapsed_time(ev[k * per + j + 1]) / n
        return dict(zip(self.phase_names, out))

    def close(self):
        if self.transport == "p2p":
            if getattr(self, "_flags", None) is not None:
                torch.cuda.synchronize()
                dist.barrier(group=self.group)  # no peer is still spinning on (or about to write) our flags
                self._flags.close()
                self._flags = None
            self.block_buf.close()
            if self.final_buf is not None:
                self.final_buf.close()


def four_step_shape(n: int, world: int) -> tuple[int, int]:
    """(H, W) with H * W = n, H <= W, both divisible by `world`: the matrix view of the distributed 1D transform."""
    if not _is_pow2(n) or not _is_pow2(world):
        raise ValueError("n and the number of ranks must be powers of two")
    lg = n.bit_length() - 1
    h = 1 << (lg // 2)
    w = n // h
    if h < 2 or h % world or w % world or n < 8:
        raise ValueError(f"n = {n} is too small to split over {world} ranks (needs H = 2^floor(log2(n)/2) >= ranks)")
    return h, w


class DistFFT1D:
    """genfft::FFT<T>(n) (fft.h:54-113) for one LARGE transform spread over a process group: the four-step algorithm
    on the slab decomposition.  The n = H*W points are an H x W row-major matrix x[r*W + c]; rank g holds rows
    [g*H/P, (g+1)*H/P), i.e. its contiguous n/P input points.  With k = kr + H*kc:

        X[kr + H*kc] = sum_c W_W^(c*kc) * W_n^(c*kr) * sum_r W_H^(r*kr) x[r*W + c]

    1. global transpose 1: every rank gets the (H x W/P) column block of its columns c;
    2. length-H column transforms on that block, the result scattered back to row slabs (global transpose 2);
    3. twiddle W_n^(kr*c) on the (H/P x W) slab;
    4. length-W row transforms; the slab now holds Z[kr][kc] = X[kr + H*kc]  (``transposed_out=True`` stops here);
    5. natural order: global transpose 3 (fused into the row transforms' stores) + a local (H x W/P) transpose,
       after which rank g holds X[g*n/P : (g+1)*n/P].

    Transports as in DistFFT2D: ``"p2p"`` stores into IPC-mapped peer buffers from the kernels (the column / row
    transforms' stores are transposes 2 and 3; transpose 1 is a strided peer copy), ``"nccl"`` packs, calls
    ``all_to_all_single`` and unpacks.  Unscaled inverse with ``inv=True``, as FFT<T>::transform<true>.
    """

    def __init__(self, n: int, dtype=np.float32, group=None, transport: str = "p2p", transposed_out: bool = False,
                 engine=None, barrier: str = "flags"):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if transport not in ("p2p", "nccl"):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        if barrier not in ("flags", "collective"):
            raise ValueError("barrier must be 'flags' or 'collective'")
        self.n = n
        self.h, self.w = four_step_shape(n, self.world)
        self.hl, self.wp = self.h // self.world, self.w // self.world
        self.transport = transport
        self.transposed_out = transposed_out
        self.engine = engine if engine is not None else CudaSlabEngine(self.w, self.h, self.world, dtype)
        e = self.engine
        self.out = e.empty(self.hl, self.w) if transposed_out else e.empty(self.wp, self.h)
        if transport == "nccl":
            self.send = e.empty(self.world, self.hl, self.wp)
            self.block = e.empty(self.h, self.wp)
            self.block_out = e.empty(self.h, self.wp)
            self.recv2 = e.empty(self.world, self.hl, self.wp)
            self.slab = e.empty(self.hl, self.w)
        else:
            esz = 8 if e.cdtype == torch.complex64 else 16
            self.block_buf = PeerBuffers(self.h * self.wp * esz, group)
            self.slab_buf = PeerBuffers(self.hl * self.w * esz, group)
            self.block = self.block_buf.tensor((self.h, self.wp), e.cdtype)
            self.slab = self.slab_buf.tensor((self.hl, self.w), e.cdtype)
            self._token = torch.zeros(1, device="cuda")
            self._flags = None
            if barrier == "flags" and engine is None:
                self._flags = PeerBuffers(256, group)
                check(lib().genfft_cuda_memset_dev(self._flags.local, 0, 256))
                self._flag_ptrs = (C.c_void_p * self.world)(*self._flags.ptrs)
                self._epoch = 0
                torch.cuda.synchronize()
                dist.barrier(group=self.group)

    _stream_barrier = DistFFT2D._stream_barrier

    def transform(self, shard, inv: bool = False):
        """`shard`: this rank's n/P contiguous input points.  Returns this rank's n/P natural-order output points, or
        with ``transposed_out`` its (H/P x W) slab Z[kr][kc] = X[kr + H*kc]."""
        e = self.engine
        if shard.numel() != self.hl * self.w:
            raise ValueError(f"expected this rank's {self.hl * self.w} contiguous points")
        x = shard.reshape(self.hl, self.w)
        row0 = self.rank * self.hl
        if self.transport == "nccl":
            e.pack_cols(self.send, x)
            dist.all_to_all_single(self.block.view(self.world, self.hl, self.wp), self.send, group=self.group)
            e.cols(self.block_out, self.block, inv)
            dist.all_to_all_single(self.recv2, self.block_out.view(self.world, self.hl, self.wp), group=self.group)
            e.unpack(self.slab, self.recv2)
            e.twiddle(self.slab, row0, inv)
            if self.transposed_out:
                e.rows(self.out, self.slab, inv)
                return self.out
            e.rows_pack(self.slab, self.send, inv)
            dist.all_to_all_single(self.block.view(self.world, self.hl, self.wp), self.send, group=self.group)
            e.transpose(self.out, self.block)
            return self.out.view(-1)
        mark = self._mark
        mark()
        self._stream_barrier()  # peers are done with the block / slab buffers of the previous call
        mark()
        e.cols_blocks_to_peers(x, self.block_buf.ptrs, self.rank)                    # transpose 1
        mark()
        self._stream_barrier()
        mark()
        # column transforms + transpose 2, the twiddle W_n^(kr*c) fused into the peer stores when the pass allows it
        fused = e.cols_to_peers(self.block_buf.local, self.slab_buf.ptrs, self.rank, inv, twiddle_n=self.n)
        mark()
        self._stream_barrier()
        mark()
        if not fused:
            e.twiddle(self.slab, row0, inv)
        mark()
        if self.transposed_out:
            e.rows(self.out, self.slab, inv)
            mark()
            return self.out
        e.rows_to_peers(self.slab, self.block_buf.ptrs, self.rank, inv)              # row transforms + transpose 3
        mark()
        self._stream_barrier()
        mark()
        e.transpose(self.out, self.block)
        mark()
        return self.out.view(-1)

    # optional per-phase timing of the p2p path (bench_dist.py --one-d N --phases)
    phase_names = ("barrier0", "transpose1", "barrier1", "cols+transpose2", "barrier2", "twiddle", "rows+transpose3",
                   "barrier3", "local_transpose")
    _events = None
    start_phase_timing = DistFFT2D.start_phase_timing
    _mark = DistFFT2D._mark

    def phase_times_ms(self):
        torch.cuda.synchronize()
        ev, self._events = self._events, None
        per = 8 if self.transposed_out else 10
        names = self.phase_names[:6] + ("rows",) if self.transposed_out else self.phase_names
        n = len(ev) // per
        out = [0.0] * (per - 1)
        for k in range(n):
            for j in range(per - 1):
                out[j] += ev[k * per + j].elapsed_time(ev[k * per + j + 1]) / n
        return dict(zip(names, out))

    def close(self):
        if self.transport == "p2p":
            if getattr(self, "_flags", None) is not None:
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
                self._flags.close()
                self._flags = None
            self.block_buf.close()
            self.slab_buf.close()
