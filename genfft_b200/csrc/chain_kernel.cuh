// L2-resident pass chains: two consecutive Stockham passes A -> B of a multi-pass transform in ONE launch, so that
// the intermediate never makes the round trip through HBM.
//
// The data is cut into groups of a few MiB whose B tiles depend only on the A tiles of the same group (whole rows /
// columns / transforms, or the column classes p mod R of a three-pass transform, see pass_chain.cu).  A persistent grid of
// CTAs takes tiles from an atomic ticket counter in the order
//     A(0) .. A(lag-1) | A(s) interleaved 1:1 with B(s-lag), s = lag .. ngroups-1 | B(ngroups-lag) .. B(ngroups-1)
// so the chip always works on an HBM-bound pass (A reads its input from HBM) and on an L2-bound pass (B reads what
// A wrote a few microseconds earlier, still resident in the 126 MB L2) at the same time.  A per-group counter of
// finished A tiles gates the B tiles of that group; a B tile's producers always hold lower tickets, i.e. they are
// running or done, so the wait cannot deadlock whatever order the hardware dispatches CTAs in.  When B is the pass
// that stores to the peer GPUs (distributed 2D), the same interleaving overlaps the NVLink-bound pass with the
// HBM-bound one tile by tile.
//
// Measured with plain copies (tools/chain_bench.cu, B200): two launches X->Y, Y->Y 1.27 ms per 2 GiB; one chained
// launch with 4-8 MiB groups 0.88 ms (L2-hit reads, one write-back), falling back to HBM speed above ~32 MiB groups.
#pragma once
#include "tile_kernel.cuh"

namespace genfft_cuda {

struct ChainParams {
  PassParams a, b;  // addressing of group 0
  // group g = (ghi, glo) = (g / gdiv, g % gdiv); element offsets added to the passes' in/out (and to out_peer[])
  uint32_t gdiv;
  long long a_in_hi, a_in_lo, a_out_hi, a_out_lo;
  long long b_in_hi, b_in_lo, b_out_hi, b_out_lo;
  uint32_t a_p_lo, b_p_lo;  // twiddle-column offset per glo
  uint32_t ngroups, lag;
  uint32_t ta, tb;  // tiles per group
  uint32_t* ctr;    // [0]: ticket, [1 + g]: finished A tiles of group g; zeroed before the launch
};

#ifndef GENFFT_EMU  // emu_device.h restates these for the CPU tests
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

template <typename T, int THREADS, int P = 16>
constexpr int chain_min_blocks() {
  return min_blocks<T, THREADS, P>();
}

// KA, KB: TileKernel instantiations with the same thread count.  The grid is persistent (one CTA per resident slot);
// every CTA loops over tickets and fetches the next one while it works on the current tile, so the atomic's round
// trip is off the critical path.
template <typename T, class KA, class KB>
__global__ void __launch_bounds__(KA::THREADS, chain_min_blocks<T, KA::THREADS, KA::TN == 0 ? 16 : KA::P_PER_THREAD>())
fft_chain_kernel(const __grid_constant__ ChainParams cp) {
  static_assert(KA::THREADS == KB::THREADS, "chained passes must have the same CTA size");
  GENFFT_DYN_SMEM(smem_raw);
  __shared__ uint32_t s_ticket[2];
  const uint32_t ta = cp.ta, tb = cp.tb, ng = cp.ngroups, lag = cp.lag;
  const uint32_t per = ta + tb;
  const uint32_t total = ng * per;
  uint32_t next_ticket = 0;  // thread 0 only
  if (threadIdx.x == 0) s_ticket[0] = atomicAdd(cp.ctr, 1u);
  __syncthreads();
  for (uint32_t it = 0;; it ^= 1u) {
    const uint32_t t = s_ticket[it];
    if (t >= total) break;
    if (threadIdx.x == 0) next_ticket = atomicAdd(cp.ctr, 1u);
    bool is_b;
    uint32_t g, i;
    if (t < lag * ta) {
      is_b = false;
      g = t / ta;
      i = t - g * ta;
    } else {
      const uint32_t t1 = t - lag * ta;
      const uint32_t mid = (ng - lag) * per;
      if (t1 < mid) {
        const uint32_t s = t1 / per, r = t1 - s * per;
        const uint32_t m = min(ta, tb);
        if (r < 2u * m) {
          is_b = r & 1u;
          i = r >> 1;
        } else {
          is_b = tb > ta;
          i = r - m;
        }
        g = is_b ? s : s + lag;
      } else {
        const uint32_t t2 = t1 - mid;
        is_b = true;
        g = ng - lag + t2 / tb;
        i = t2 % tb;
      }
    }
    const uint32_t ghi = g / cp.gdiv, glo = g - ghi * cp.gdiv;
    if (!is_b) {
      KA::template tile_once<CO_STREAM, CO_DEFAULT>(cp.a, reinterpret_cast<cpx<T>*>(smem_raw), i,
                                                    (long long)ghi * cp.a_in_hi + (long long)glo * cp.a_in_lo,
                                                    (long long)ghi * cp.a_out_hi + (long long)glo * cp.a_out_lo,
                                                    glo * cp.a_p_lo);
      if (threadIdx.x == 0) s_ticket[it ^ 1u] = next_ticket;
      __syncthreads();  // every thread's stores are issued; the exchange buffer is free for the next tile
      // release at device scope, cumulative over the CTA's stores ordered before it by the barrier
      if (threadIdx.x == 0) red_release_gpu_add(cp.ctr + 1 + g, 1u);
    } else {
      if (threadIdx.x == 0) {
        uint32_t spins = 0;
        while (ld_relaxed_gpu(cp.ctr + 1 + g) < ta) {
          __nanosleep(32);
          if (++spins > (1u << 26)) __trap();  // seconds: a scheduling bug must fail loudly, never hang the GPU
        }
        (void)ld_acquire_gpu(cp.ctr + 1 + g);  // synchronizes with the producers' release increments
      }
      __syncthreads();
      KB::template tile_once<CO_L2ONLY, CO_STREAM>(cp.b, reinterpret_cast<cpx<T>*>(smem_raw), i,
                                                   (long long)ghi * cp.b_in_hi + (long long)glo * cp.b_in_lo,
                                                   (long long)ghi * cp.b_out_hi + (long long)glo * cp.b_out_lo,
                                                   glo * cp.b_p_lo);
      if (threadIdx.x == 0) s_ticket[it ^ 1u] = next_ticket;
      __syncthreads();
    }
  }
}

}  // namespace genfft_cuda
