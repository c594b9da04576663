// The one butterfly kernel of the engine: a CTA transforms a tile of C independent length-L
// sequences ("columns") entirely on chip -- registers for the radix-P butterflies, padded shared
// memory for the exchanges between radix stages (Stockham auto-sort, so no bit-reversal pass
// exists anywhere: the reference's scramble, include/genFFT/FFTUtil.h:37-83, which is 20-80 % of
// its run time, has no counterpart here).
//
// Everything the hot path does is this kernel with different addressing:
//   * batched contiguous 1D (FFT<T>::transform, fft.h:80-85): columns = transforms, lanes run along
//     the sequence (map B) so global loads/stores are fully coalesced;
//   * column / strided passes (FFTVert, fft.h:145-150; the column pass of FFT2D, fft.h:216-217;
//     every pass of a large multi-pass 1D transform): columns are adjacent in memory, lanes run
//     across columns (map A) so each row of the tile is one contiguous segment;
//   * multi-pass large N: the Stockham inter-pass twiddle W_M^(p*i) is fused into the load, the
//     output permutation into the store;
//   * distributed 2D: the store can scatter the tile over up to 8 peer GPUs' buffers (NVLink
//     stores), fusing the all-to-all transpose into the producing pass.
#pragma once
#include <cstdint>
#include <type_traits>
#include "launch.h"
#include "radix.cuh"

namespace genfft_cuda {

constexpr int kMaxPeers = 8;

struct PassParams {
  const void* in;
  void* out;
  void* out_peer[kMaxPeers];  // used when out_split_log2 >= 0 && use_peers
  // tile enumeration: tile -> (t0, t1, t2), t2 fastest; t2 enumerates blocks of C columns
  uint32_t n1, n2;
  uint32_t ntiles;
  long long in_t0, in_t1;    // element offsets per tile index (t2 handled through the column index)
  long long out_t0, out_t1;
  long long in_stride_i;     // between consecutive sequence elements
  long long in_stride_c;     // between adjacent columns
  long long out_stride_k;    // between consecutive output bins (low part when split)
  long long out_stride_c;
  long long out_stride_khi;  // stride of (k >> out_split_log2) when out_split_log2 >= 0
  int out_split_log2;        // -1: plain affine output addressing
  int use_peers;             // 1: (k >> out_split_log2) selects out_peer[], out_stride_khi ignored
  int ncols;                 // valid columns along t2*C + c (tail tiles are masked)
  int map_load, map_store;   // 0 = lanes across columns (A), 1 = lanes along the sequence (B)
  int mode;                  // host side only: which compiled addressing mode to launch (enum Mode)
  int pdl;                   // host side only: launch with programmatic stream serialization (see fft_tile_kernel)
  float grid_frac;           // host side only: fraction of the resident-CTA capacity to launch (0 = all)
  uint32_t tiles_per_cta;    // 0: grid-stride loop over tiles (persistent grid); K > 0: CTA b owns tiles [b*K, b*K+K)
  // tile -> (t0, t1, t2) without integer division: q = umulhi(x, mul) >> shr (mul == 0: divisor is 1), see fast_div()
  uint32_t n1_mul, n1_shr, n2_mul, n2_shr;
  int inverse;               // conjugate on load and on store
  int in_real;               // 1: input is real scalars (imag = 0), FFT<T>::transform_real, fft.h:90-94
                             // 2: real part from `in`, imaginary part from `in2`, transform_interleave, fft.h:100-105
  const void* in2;
  // bit-reversed input (transform_no_scramble contract, fft.h:69-73 / 132-136):
  //   offset = brev(t1*g_t1 + col*g_c + idx*g_i, brev_bits) * brev_stride + t0*in_t0 + col*in_stride_c
  int brev_bits;
  long long g_t1, g_c, g_i, brev_stride;
  // inter-pass Stockham twiddle W_M^(p*idx), M = Ns*L, p = t1*p_t1 + col*p_c, idx = u + i*TN, factored as
  //   W_M^(p*u)            one per thread, two-level table:  tw_hi[e >> tw_shift] * tw_lo[e & mask], e = p*u
  //   W_M^(p*i*TN)         = W_{P*Ns}^(p*i), table tw_b[i*tw_b_stride + p] (lanes read consecutive p: coalesced)
  const void* tw_hi;
  const void* tw_lo;
  int tw_shift;
  const void* tw_b;
  long long tw_b_stride;  // = Ns
  // M_PEER* only, distributed four-step 1D transform: the twiddle W_n^(kr*c) between its column and row transforms
  // fused into the column pass's peer store (kr = output row p + k*Ns, c = tw2_col0 + column), two-level table
  // tw2_hi[e >> tw2_shift] * tw2_lo[e & mask] of W_n; null = no twiddle
  const void* tw2_hi;
  const void* tw2_lo;
  int tw2_shift;
  uint32_t tw2_col0;
  // fused real-FFT split (M_ROWDIT): the tile's L-point complex transforms are the packed halves of 2L-point
  // real signals; dit_tw[k] = W_{2L}^k, k < L; dit_half: write L+1 bins only, else all 2L
  const void* dit_tw;
  int dit_half;
  // M_COLTWDIT: the L-point pass is the last of an M = Ns*L point packed transform (n = 2M real points):
  //   W_n^(p + k*Ns) = dit_a[p] * dit_tw[k],  dit_a[p] = W_n^p (p < Ns),  dit_tw[k] = W_{2L}^k (k < L)
  const void* dit_a;
  int p_t1, p_c;
  uint32_t p_mask;  // p &= p_mask
  // on-chip stage twiddles, one block per radix stage s >= 1 laid out [q][p]:
  //   tw_L[stage_tw_offset(s) + q*NS + p] = W_{NS*R}^(p*q)   (lanes read consecutive p: coalesced)
  const void* tw_L;
  // The half-spectrum inverse's pre-process (aux_kernels.cuh c2r_pre_kernel) fused into the first pass's load,
  // in_real == 3 (measured: batched n = 4096 0.879 -> 0.743 ms, profiles/r02_ab_fused_c2r.log).  Element s of the packed spectrum is built from the bins X[s] and X[M - s] of
  // the n/2+1 input bins: s = col*c2r_sc + idx*in_stride_i relative to the transform's first bin, which is at
  // t.in_off (+ col*in_stride_c when the columns are whole transforms, c2r_sc == 0)
  uint32_t c2r_m;
  int c2r_sc;
  const void* c2r_hi;  // two-level table of W_n^e, n = 2M
  const void* c2r_lo;
  int c2r_shift;
};

__host__ __device__ constexpr int stage_radix(int L, int P, int s) {
  int rem = L;
  for (int k = 0; k < s; k++) rem /= P;
  return rem >= P ? P : rem;
}
__host__ __device__ constexpr int num_stages(int L, int P) {
  int n = 0, rem = L;
  while (rem > 1) { rem = rem >= P ? rem / P : 1; n++; }
  return n;
}
__host__ __device__ constexpr int stage_ns(int L, int P, int s) {
  int ns = 1;
  for (int k = 0; k < s; k++) ns *= stage_radix(L, P, k);
  return ns;
}

// offset of stage s's block in the stage-twiddle table (stage 0 needs none)
__host__ __device__ constexpr int stage_tw_offset(int L, int P, int s) {
  int off = 0;
  for (int k = 1; k < s; k++) off += stage_ns(L, P, k) * stage_radix(L, P, k);
  return off;
}
__host__ __device__ constexpr int stage_tw_size(int L, int P) { return stage_tw_offset(L, P, num_stages(L, P)); }

// Shared-memory padding: one element per 2^PADSH.  A 128-byte wavefront covers 16 float2 or 8 double2 elements; the
// radix-P scatter of the first stage has element stride P, which one pad element per 16 makes conflict-free for
// P = 16 (stride 17) in both precisions, while radix-8 on double2 needs one per 8 (stride 9).
template <typename T, int P>
constexpr int pad_shift() { return (sizeof(T) == 8 && P == 8) ? 3 : 4; }
template <int PADSH>
__host__ __device__ constexpr int pad_idx_t(int i) { return i + (i >> PADSH); }
template <int PADSH>
__host__ __device__ constexpr int tile_pitch_t(int L) { return pad_idx_t<PADSH>(L) | 1; }

// Addressing modes.  M_GEN handles everything at run time (bit-reversed / real input, split and peer
// stores, any mapping); the others are the hot paths with the addressing resolved at compile time so
// that global and shared accesses use immediate offsets from one base register and the direction
// (INV: conjugate on load and store) folds into the butterflies' operand negations.
//   M_ROW   : unit stride in and out, lanes along the sequence        (batched contiguous 1D, 2D rows)
//   M_COL   : strided in and out, lanes across columns                (FFTVert, 2D columns, last pass)
//   M_COLTW : M_COL + inter-pass Stockham twiddle fused into the load (later passes of a large N)
//   M_FIRST : strided in (lanes across columns), unit-stride out      (first pass of a large N)
//   M_ROWTMA: M_ROW with the next tile prefetched by cp.async.bulk (TMA) into a second shared buffer while
//             the current one is transformed; an mbarrier signals arrival        (batched contiguous 1D)
//   M_ROWDIT: M_ROW + the real-FFT split fused after the last butterfly stage    (RealFFT<T>, n <= 2*Lmax)
//   M_COLTWDIT: last pass of a large real transform: M_COLTW on a PAIR of column groups {p} and {Ns-p} so that the
//             real-FFT split, which couples bins q and M-q, is fused after the last butterfly stage
//   M_PEER2/4/8: M_COLTW whose store is the all-to-all of the distributed transforms: the last pass of a length-N
//             transform split over NP ranks sends bin k = u + i*TN to rank k / (L/NP) = i / (16/NP) -- a compile-time
//             function of the register index -- so a thread's 16 stores are NP groups of 16/NP stores at immediate
//             multiples of one stride from NP peer bases (CUDA-IPC mapped buffers, NVLink stores)
enum Mode { M_GEN = 0, M_ROW = 1, M_COL = 2, M_COLTW = 3, M_FIRST = 4, M_ROWTMA = 5, M_ROWDIT = 6, M_COLTWDIT = 7,
            M_PEER2 = 8, M_PEER4 = 9, M_PEER8 = 10 };
__host__ __device__ constexpr int peer_mode_ranks(int mode) { return mode == M_PEER2 ? 2 : mode == M_PEER4 ? 4 : mode == M_PEER8 ? 8 : 0; }

// Division of a tile index (< 2^31) by a launch-invariant divisor d without the ~25-instruction software division:
// host side  shr = ceil(log2 d) - 1, mul = ceil(2^(32 + shr) / d)  (d >= 2; mul = 0 encodes d == 1),
// device side q = umulhi(x, mul) >> shr.  Exact for x < 2^31 (the usual round-up magic number; checked exhaustively
// over the divisors the plans use in tests/test_abi.py through genfft_cuda_debug_fast_div).
struct FastDiv {
  uint32_t mul, shr;
};
inline FastDiv make_fast_div(uint32_t d) {
  FastDiv f{0u, 0u};
  if (d <= 1) return f;
  uint32_t lg = 0;
  while ((1ull << lg) < d) lg++;  // ceil(log2 d) >= 1
  f.shr = lg - 1;
  f.mul = (uint32_t)((((unsigned long long)1 << (32 + f.shr)) + d - 1) / d);
  return f;
}
__host__ __device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t mul, uint32_t shr) {
#ifdef __CUDA_ARCH__
  return mul ? (__umulhi(x, mul) >> shr) : x;
#else
  return mul ? (uint32_t)(((unsigned long long)x * mul) >> 32) >> shr : x;
#endif
}

#ifdef GENFFT_EMU
#include "emu_device.h"  // tests/emu: host restatement of the PTX helpers below (CPU tests only)
#else
// ---- mbarrier / bulk-copy PTX (sm_90+; SASS: SYNCS / UBLKCP) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif  // GENFFT_EMU

// Cache operators of the data loads/stores.  The stand-alone kernels use the defaults; the L2-resident pass chains
// (chain_kernel.cuh) stream their HBM-side traffic (evict-first) and read the intermediate, which another SM wrote
// during the same launch, past the non-coherent L1.
enum CacheOp { CO_DEFAULT = 0, CO_STREAM = 1, CO_L2ONLY = 2 };
template <int OP, typename V>
__device__ __forceinline__ V ld_data(const V* p) {
  if constexpr (OP == CO_STREAM) return __ldcs(p);
  else if constexpr (OP == CO_L2ONLY) return __ldcg(p);
  else return *p;
}
template <int OP, typename V>
__device__ __forceinline__ void st_data(V* p, const V& v) {
  if constexpr (OP == CO_STREAM) __stcs(p, v);
  else *p = v;
}

#ifndef GENFFT_EMU
// Strided global access  base[stride * k]  with the address formed by ONE instruction (IMAD.WIDE.U32 with an
// immediate): `stride` is a 32-bit element stride, `k` a compile-time element count after unrolling.  Written in PTX
// because the compiler otherwise strength-reduces the sixteen addresses of a tile column into chains of 64-bit
// adds (2-3 instructions each); the access itself is PTX too so that it stays a global (not generic) access.
template <typename V>
__device__ __forceinline__ unsigned long long strided_addr(const V* base, uint32_t stride, uint32_t k) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(stride), "r"(k * (uint32_t)sizeof(V)), "l"(base));
  return r;
}
template <int OP>
__device__ __forceinline__ float2 ld_strided(const float2* base, uint32_t stride, uint32_t k) {
  float2 v;
  const unsigned long long a = strided_addr(base, stride, k);
  if constexpr (OP == CO_STREAM) asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a));
  else if constexpr (OP == CO_L2ONLY) asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a));
  else asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a));
  return v;
}
template <int OP>
__device__ __forceinline__ double2 ld_strided(const double2* base, uint32_t stride, uint32_t k) {
  double2 v;
  const unsigned long long a = strided_addr(base, stride, k);
  if constexpr (OP == CO_STREAM) asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a));
  else if constexpr (OP == CO_L2ONLY) asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a));
  else asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a));
  return v;
}
template <int OP>
__device__ __forceinline__ void st_strided(float2* base, uint32_t stride, uint32_t k, const float2& v) {
  const unsigned long long a = strided_addr(base, stride, k);
  if constexpr (OP == CO_STREAM) asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(v.x), "f"(v.y) : "memory");
  else asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(v.x), "f"(v.y) : "memory");
}
template <int OP>
__device__ __forceinline__ void st_strided(double2* base, uint32_t stride, uint32_t k, const double2& v) {
  const unsigned long long a = strided_addr(base, stride, k);
  if constexpr (OP == CO_STREAM) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(a), "d"(v.x), "d"(v.y) : "memory");
  else asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(a), "d"(v.x), "d"(v.y) : "memory");
}
// read-only table entry (twiddles): non-coherent path
__device__ __forceinline__ float2 ldg_strided(const float2* base, uint32_t stride, uint32_t k) {
  float2 v;
  asm("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(strided_addr(base, stride, k)));
  return v;
}
__device__ __forceinline__ double2 ldg_strided(const double2* base, uint32_t stride, uint32_t k) {
  double2 v;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(strided_addr(base, stride, k)));
  return v;
}
#endif  // GENFFT_EMU

template <typename T, int L, int P, int C, int MODE, bool INV>
struct TileKernel {
  static constexpr int TN = L / P;
  static constexpr int P_PER_THREAD = P;
  static constexpr int THREADS = TN * C;
  static constexpr int NST = num_stages(L, P);
  static constexpr int PADSH = pad_shift<T, P>();
  static constexpr int PADN = 1 << PADSH;
  static constexpr int PITCH = tile_pitch_t<PADSH>(L);
  static __host__ __device__ constexpr int pad_idx(int i) { return pad_idx_t<PADSH>(i); }
  static constexpr bool TMA = MODE == M_ROWTMA;
  static constexpr bool DIT = MODE == M_ROWDIT;
  static constexpr bool PAIR = MODE == M_COLTWDIT;
  static constexpr int NPEER = peer_mode_ranks(MODE);
  static constexpr bool PEER = NPEER > 0;
  static constexpr int HALF_C = C / 2;
  static constexpr bool ROWLIKE = MODE == M_ROW || TMA || DIT;
  static constexpr size_t XBUF_BYTES = (NST > 1 || DIT || PAIR) ? sizeof(cpx<T>) * (size_t)PITCH * C : 0;
  // TMA mode: [exchange buffer][input buffer C*L][mbarrier]
  static constexpr size_t INBUF_OFFSET = (XBUF_BYTES + 127) & ~(size_t)127;
  static constexpr size_t SMEM_BYTES = TMA ? INBUF_OFFSET + sizeof(cpx<T>) * (size_t)L * C + 16 : XBUF_BYTES;
  static constexpr bool GEN = MODE == M_GEN;
  static constexpr bool UNIT_IN = ROWLIKE;
  static constexpr bool UNIT_OUT = ROWLIKE || MODE == M_FIRST;
  using V = typename vec2<T>::type;

  static __device__ __forceinline__ uint32_t brev(uint32_t v, int bits) { return __brev(v) >> (32 - bits); }

  struct Tile {
    long long in_off, out_off;
    uint32_t p_base;
    uint32_t col0;
    long long g_base;
    uint32_t grp_col;  // chains: first column of the tile's group within the pass (column passes; 0 otherwise)
  };

  static __device__ __forceinline__ Tile decode(const PassParams& prm, uint32_t tile) {
    Tile t;
    t.grp_col = 0;
    if constexpr (ROWLIKE) {  // columns are whole transforms; no outer tile indices
      t.col0 = tile * C;
      t.in_off = t.out_off = 0;
      t.p_base = 0;
      t.g_base = 0;
    } else {
    const uint32_t t01 = fast_div(tile, prm.n2_mul, prm.n2_shr);
    const uint32_t t2 = tile - t01 * prm.n2;
    const uint32_t t0 = fast_div(t01, prm.n1_mul, prm.n1_shr);
    const uint32_t t1 = t01 - t0 * prm.n1;
    t.col0 = t2 * C;
    t.in_off = (long long)t0 * prm.in_t0 + (long long)t1 * prm.in_t1;
    t.out_off = (long long)t0 * prm.out_t0 + (long long)t1 * prm.out_t1;
    t.p_base = t1 * (uint32_t)prm.p_t1;
    t.g_base = (long long)t1 * prm.g_t1;
    if (GEN && prm.brev_bits) t.in_off = (long long)t0 * prm.in_t0;
    }
    return t;
  }

  // column (sequence index within the tile enumeration) handled by lane-column c of this tile
  static __device__ __forceinline__ uint32_t column_of(const PassParams& prm, const Tile& t, int c, bool& valid) {
    if constexpr (PAIR) {
      // pair tile j: columns {j*H+1 .. (j+1)*H} and their mirrors {Ns-(j+1)*H .. Ns-j*H-1}; the extra last tile is
      // column 0 alone (its bins pair with themselves).  Ns/2 would appear twice: its mirror-side copy is masked.
      const uint32_t Ns = (uint32_t)prm.ncols;
      const uint32_t j = t.col0 / C;
      if (j == Ns / C) {
        valid = (c == 0);
        return 0;
      }
      const uint32_t p = c < HALF_C ? j * HALF_C + 1 + c : Ns - (j + 1) * HALF_C + (c - HALF_C);
      valid = !(c >= HALF_C && p == Ns / 2);
      return p;
    } else {
      const uint32_t col = t.col0 + c;
      valid = col < (uint32_t)prm.ncols;
      return col;
    }
  }

  static __device__ __forceinline__ bool is_inverse(const PassParams& prm) { return GEN ? prm.inverse != 0 : INV; }

  static __device__ __forceinline__ void apply_pass_twiddle(const PassParams& prm, const Tile& t, uint32_t col, int u,
                                                            cpx<T> (&x)[P]) {
    const uint32_t p = (t.p_base + col * (uint32_t)prm.p_c) & prm.p_mask;
    const V* hi = reinterpret_cast<const V*>(prm.tw_hi);
    const V* lo = reinterpret_cast<const V*>(prm.tw_lo);
    const uint32_t e = p * (uint32_t)u;
    V ah = __ldg(hi + (e >> prm.tw_shift));
    V al = __ldg(lo + (e & ((1u << prm.tw_shift) - 1u)));
    const cpx<T> a = cmul(cpx<T>(ah.x, ah.y), cpx<T>(al.x, al.y));  // W_M^(p*u)
    const V* tb = reinterpret_cast<const V*>(prm.tw_b) + p;
    const uint32_t ts32 = (uint32_t)prm.tw_b_stride;
    if constexpr (P == 16) {
      // b_i = W_{16 Ns}^(p*i): rows 1, 4, 8, 12 of the table are read, the rest are products b_{4h} * b_j with
      // b_2 = b_1^2, b_3 = b_2 b_1 -- 4 loads and 33 multiplications instead of 15 loads and 32 multiplications, at
      // most three roundings deep (measured, profiles/r02_ab_twiddle_powers.log: C3 272.7 -> 258 us, 2^21 x 256
      // 3.79 -> 3.65 ms, C5 -1..-5 %; rel-L2 vs genFFT 2.4e-7 in float, 9e-16 in double at 2^24)
      const V v1 = ldg_strided(tb, ts32, 1u), v4 = ldg_strided(tb, ts32, 4u), v8 = ldg_strided(tb, ts32, 8u),
              v12 = ldg_strided(tb, ts32, 12u);
      const cpx<T> b1(v1.x, v1.y);
      const cpx<T> b2 = cmul(b1, b1), b3 = cmul(b2, b1);
      const cpx<T> a4[4] = {a, cmul(a, cpx<T>(v4.x, v4.y)), cmul(a, cpx<T>(v8.x, v8.y)), cmul(a, cpx<T>(v12.x, v12.y))};
#pragma unroll
      for (int h = 0; h < 4; h++) {
        x[4 * h + 0] = cmul(x[4 * h + 0], a4[h]);
        x[4 * h + 1] = cmul(x[4 * h + 1], cmul(a4[h], b1));
        x[4 * h + 2] = cmul(x[4 * h + 2], cmul(a4[h], b2));
        x[4 * h + 3] = cmul(x[4 * h + 3], cmul(a4[h], b3));
      }
      return;
    }
    V b[P];
#pragma unroll
    for (int i = 1; i < P; i++) b[i] = ldg_strided(tb, ts32, (uint32_t)i);
    x[0] = cmul(x[0], a);
#pragma unroll
    for (int i = 1; i < P; i++) x[i] = cmul(x[i], cmul(a, cpx<T>(b[i].x, b[i].y)));
  }

  template <int LDOP = CO_DEFAULT>
  static __device__ __forceinline__ void load(const PassParams& prm, const Tile& t, int c, int u, cpx<T> (&x)[P]) {
    bool valid;
    const uint32_t col = column_of(prm, t, c, valid);
    if constexpr (!GEN) {
      // Strides are below 2^32 elements here (launch_pass sends anything larger to M_GEN): every strided address is
      // base + stride32 * constant.  (The branch around the loads is kept on purpose: without it ptxas hoists the
      // later stages' twiddle loads to the top of the kernel and spills them.)
      const long long base = t.in_off + (long long)col * prm.in_stride_c;
      if (valid) {
        if constexpr (UNIT_IN) {
          const V* src = reinterpret_cast<const V*>(prm.in) + base + u;
#pragma unroll
          for (int i = 0; i < P; i++) {
            V v = ld_data<LDOP>(src + i * TN);
            x[i] = cpx<T>(v.x, INV ? -v.y : v.y);
          }
        } else {
          const V* src = reinterpret_cast<const V*>(prm.in) + base + (long long)u * prm.in_stride_i;
          const uint32_t s32 = (uint32_t)prm.in_stride_i;
#pragma unroll
          for (int i = 0; i < P; i++) {
            V v = ld_strided<LDOP>(src, s32, (uint32_t)(i * TN));
            x[i] = cpx<T>(v.x, INV ? -v.y : v.y);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < P; i++) x[i] = cpx<T>(T(0), T(0));
      }
      if constexpr (MODE == M_COLTW || PAIR || PEER) apply_pass_twiddle(prm, t, col, u, x);
    } else {
    const long long base = t.in_off + (long long)col * prm.in_stride_c;
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int idx = u + i * TN;
      long long off;
      if (prm.brev_bits) {
        uint32_t g = (uint32_t)(t.g_base + (long long)col * prm.g_c + (long long)idx * prm.g_i);
        off = base + (long long)brev(g, prm.brev_bits) * prm.brev_stride;
      } else {
        off = base + (long long)idx * prm.in_stride_i;
      }
      if (valid) {
        if (prm.in_real == 2) {
          x[i] = cpx<T>(reinterpret_cast<const T*>(prm.in)[off], reinterpret_cast<const T*>(prm.in2)[off]);
        } else if (prm.in_real == 3) {
          // Z'[s] = (X[s] + conj X[M-s]) + i (X[s] - conj X[M-s]) conj(W_n^s)   (c2r_pre_kernel, aux_kernels.cuh)
          const uint32_t sidx = col * (uint32_t)prm.c2r_sc + (uint32_t)idx * (uint32_t)prm.in_stride_i;
          const long long tbase = t.in_off + (prm.c2r_sc ? 0 : (long long)col * prm.in_stride_c);
          const V xk = reinterpret_cast<const V*>(prm.in)[tbase + sidx];
          const V xm = reinterpret_cast<const V*>(prm.in)[tbase + (prm.c2r_m - sidx)];
          const T ax = xk.x + xm.x, ay = xk.y - xm.y;
          const T dx = xk.x - xm.x, dy = xk.y + xm.y;
          const V wh = __ldg(reinterpret_cast<const V*>(prm.c2r_hi) + (sidx >> prm.c2r_shift));
          const V wl = __ldg(reinterpret_cast<const V*>(prm.c2r_lo) + (sidx & ((1u << prm.c2r_shift) - 1u)));
          const cpx<T> w = cmul(cpx<T>(wh.x, wh.y), cpx<T>(wl.x, wl.y));  // W_n^s
          const T tx = dx * w.x + dy * w.y, ty = dy * w.x - dx * w.y;    // d * conj(w)
          x[i] = cpx<T>(ax - ty, ay + tx);
        } else if (prm.in_real) {
          x[i] = cpx<T>(reinterpret_cast<const T*>(prm.in)[off], T(0));
        } else {
          V v = ld_data<LDOP>(reinterpret_cast<const V*>(prm.in) + off);
          x[i] = cpx<T>(v.x, v.y);
        }
      } else {
        x[i] = cpx<T>(T(0), T(0));
      }
    }
    if (prm.inverse) {
#pragma unroll
      for (int i = 0; i < P; i++) x[i].y = -x[i].y;
    }
    if (prm.tw_hi) apply_pass_twiddle(prm, t, col, u, x);
    }
  }

  template <int STOP = CO_DEFAULT>
  static __device__ __forceinline__ void store(const PassParams& prm, const Tile& t, int c, int u, cpx<T> (&x)[P]) {
    const uint32_t col = t.col0 + c;
    if (col >= (uint32_t)prm.ncols) return;
    const long long base = t.out_off + (long long)col * prm.out_stride_c;
    if constexpr (PEER) {
      static_assert(!PEER || (P == 16 && NST >= 1), "peer modes are built for 16 points per thread");
      constexpr int G = P / (NPEER ? NPEER : 1);  // consecutive register slots that go to the same rank
      const uint32_t s32 = (uint32_t)prm.out_stride_k;
      const long long off = base + (long long)u * prm.out_stride_k;
      if (prm.tw2_hi) {  // uniform: the four-step 1D transform's W_n^(kr*c), kr = p + k*Ns, c = global column
        // k = u + i*TN: W^(kr*c) = a * s^i with a = W^((p + u*Ns)*c) and s = W^(TN*Ns*c), two table lookups per thread;
        // s^i = (s^4)^(i/4) * s^(i%4) from a few squarings (at most four multiplications deep, so the rounding stays
        // at the level of the two-level table itself) instead of sixteen dependent lookups in L2
        const V* hi = reinterpret_cast<const V*>(prm.tw2_hi);
        const V* lo = reinterpret_cast<const V*>(prm.tw2_lo);
        const uint32_t ns = (uint32_t)prm.tw_b_stride;
        const uint32_t cg = prm.tw2_col0 + t.grp_col + col;
        const uint32_t mask = (1u << prm.tw2_shift) - 1u;
        auto root = [&](uint32_t e) {  // e < n <= 2^31
          const V wh = __ldg(hi + (e >> prm.tw2_shift));
          const V wl = __ldg(lo + (e & mask));
          return cmul(cpx<T>(wh.x, wh.y), cpx<T>(wl.x, wl.y));
        };
        const cpx<T> a = root(((t.p_base & prm.p_mask) + (uint32_t)u * ns) * cg);
        const cpx<T> s1 = root((uint32_t)TN * ns * cg);
        const cpx<T> s2 = cmul(s1, s1), s3 = cmul(s2, s1), s4 = cmul(s2, s2);
        static_assert(P == 16, "the fused twiddle is written for 16 points per thread");
        cpx<T> ah = a;  // a * s^(4*hi)
#pragma unroll
        for (int h4 = 0; h4 < 4; h4++) {
          x[4 * h4 + 0] = cmul(x[4 * h4 + 0], ah);
          x[4 * h4 + 1] = cmul(x[4 * h4 + 1], cmul(ah, s1));
          x[4 * h4 + 2] = cmul(x[4 * h4 + 2], cmul(ah, s2));
          x[4 * h4 + 3] = cmul(x[4 * h4 + 3], cmul(ah, s3));
          if (h4 < 3) ah = cmul(ah, s4);
        }
      }
#pragma unroll
      for (int g = 0; g < NPEER; g++) {
        V* dst = reinterpret_cast<V*>(prm.out_peer[g]) + off;
#pragma unroll
        for (int j = 0; j < G; j++) {
          V v;
          v.x = x[g * G + j].x;
          v.y = INV ? -x[g * G + j].y : x[g * G + j].y;
          st_strided<CO_DEFAULT>(dst, s32, (uint32_t)(j * TN), v);
        }
      }
    } else if constexpr (!GEN) {
      if constexpr (UNIT_OUT) {
        V* dst = reinterpret_cast<V*>(prm.out) + base + u;
#pragma unroll
        for (int i = 0; i < P; i++) {
          V v;
          v.x = x[i].x;
          v.y = INV ? -x[i].y : x[i].y;
          st_data<STOP>(dst + i * TN, v);
        }
      } else {
        V* dst = reinterpret_cast<V*>(prm.out) + base + (long long)u * prm.out_stride_k;
        const uint32_t s32 = (uint32_t)prm.out_stride_k;
#pragma unroll
        for (int i = 0; i < P; i++) {
          V v;
          v.x = x[i].x;
          v.y = INV ? -x[i].y : x[i].y;
          st_strided<STOP>(dst, s32, (uint32_t)(i * TN), v);
        }
      }
    } else {
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int k = u + i * TN;
      V v;
      v.x = x[i].x;
      v.y = prm.inverse ? -x[i].y : x[i].y;
      if (prm.out_split_log2 < 0) {
        st_data<STOP>(reinterpret_cast<V*>(prm.out) + base + (long long)k * prm.out_stride_k, v);
      } else {
        const int khi = k >> prm.out_split_log2;
        const int klo = k & ((1 << prm.out_split_log2) - 1);
        if (prm.use_peers) {
          reinterpret_cast<V*>(prm.out_peer[khi])[base + (long long)klo * prm.out_stride_k] = v;
        } else {
          st_data<STOP>(reinterpret_cast<V*>(prm.out) + base + (long long)klo * prm.out_stride_k + (long long)khi * prm.out_stride_khi, v);
        }
      }
    }
    }
  }

  // radix stage S on registers; scatters to shared memory unless it is the last stage
  template <int S>
  static __device__ __forceinline__ void stage(const PassParams& prm, cpx<T> (&x)[P], cpx<T>* sm, int u) {
    constexpr int R = stage_radix(L, P, S);
    constexpr int NS = stage_ns(L, P, S);
    constexpr int M = P / R;  // butterflies per thread
    constexpr bool LAST = (S == NST - 1);
    const V* twL = reinterpret_cast<const V*>(prm.tw_L);
#pragma unroll
    for (int m = 0; m < M; m++) {
      const int j = u + m * TN;
      const int p = j & (NS - 1);
      cpx<T> y[R];
#pragma unroll
      for (int q = 0; q < R; q++) y[q] = x[m + q * M];
      if (NS > 1) {
        const V* tw = twL + (stage_tw_offset(L, P, S) + p);
#pragma unroll
        for (int q = 1; q < R; q++) {
          V w = __ldg(tw + q * NS);
          y[q] = cmul(y[q], cpx<T>(w.x, w.y));
        }
      }
      RegDFT<R>::run(y);
      if (LAST) {
        // bin k -> position p + k*NS = u + (m + k*M)*TN  -> register slot m + k*M
#pragma unroll
        for (int s = 0; s < R; s++) x[m + RegDFT<R>::out_bin(s) * M] = y[s];
      } else {
        const int base = (j - p) * R + p;
        if constexpr (NS >= PADN || NS == 1) {
          // pad_idx(base + b*NS) == pad_idx(base) + b*(NS + NS/PADN): immediate offsets from one address
          V* dst = reinterpret_cast<V*>(sm) + pad_idx(base);
#pragma unroll
          for (int s = 0; s < R; s++) {
            const int b = RegDFT<R>::out_bin(s);
            V v;
            v.x = y[s].x;
            v.y = y[s].y;
            dst[NS == 1 ? b : b * (NS + NS / PADN)] = v;
          }
        } else {
#pragma unroll
          for (int s = 0; s < R; s++) {
            const int pos = base + RegDFT<R>::out_bin(s) * NS;
            V v;
            v.x = y[s].x;
            v.y = y[s].y;
            reinterpret_cast<V*>(sm)[pad_idx(pos)] = v;
          }
        }
      }
    }
  }

  static __device__ __forceinline__ void gather(cpx<T> (&x)[P], const cpx<T>* sm, int u) {
    if constexpr (TN >= PADN) {
      const V* src = reinterpret_cast<const V*>(sm) + pad_idx(u);
#pragma unroll
      for (int i = 0; i < P; i++) {
        V v = src[i * (TN + TN / PADN)];
        x[i] = cpx<T>(v.x, v.y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; i++) {
        V v = reinterpret_cast<const V*>(sm)[pad_idx(u + i * TN)];
        x[i] = cpx<T>(v.x, v.y);
      }
    }
  }

  template <int S>
  static __device__ __forceinline__ void run_stages(const PassParams& prm, cpx<T> (&x)[P], cpx<T>* smem,
                                                     int& c, int& u, int c_st, int u_st) {
    if constexpr (S < NST) {
      stage<S>(prm, x, smem + (size_t)c * PITCH, u);
      if constexpr (S < NST - 1) {
        __syncthreads();
        if constexpr (S == NST - 2) {  // switch to the store mapping before the last stage
          c = c_st;
          u = u_st;
        }
        gather(x, smem + (size_t)c * PITCH, u);
        if constexpr (S < NST - 2) __syncthreads();  // buffer is rewritten by the next scatter
      }
      run_stages<S + 1>(prm, x, smem, c, u, c_st, u_st);
    }
  }

  // The split of one bin: X[q] = E - t O, E = (z + conj zp)/2, O = (z - conj zp)/2, t = i w, with the partner bin
  // zp = Z[M - q] and w = W_n^q (adjust_DIT_impl, include/genFFT/generic/fft_dit_impl_generic.inl:47-54).
  static __device__ __forceinline__ V split_bin(const cpx<T>& z, const V& zp, const cpx<T>& w) {
    // written with pair operations (FFMA2 / FMUL2 on the interleaved value in float, radix.cuh): S = 2E, D = 2O,
    // Q = (Im, Re) of w D, X = (S + (Q.x, -Q.y)) / 2 -- 8 instructions instead of the 16 of the scalar form
    const cpx<T> zq(zp.x, zp.y);
    const cpx<T> S = pair_fma(zq, cpx<T>(T(1), T(-1)), z);
    const cpx<T> D = pair_fma(zq, cpx<T>(T(-1), T(1)), z);
    const cpx<T> Q(w.x * D.y + w.y * D.x, w.x * D.x - w.y * D.y);
    const cpx<T> X = pair_mul(pair_fma(Q, cpx<T>(T(1), T(-1)), S), cpx<T>(T(0.5), T(0.5)));
    V f;
    f.x = X.x;
    f.y = X.y;
    return f;
  }

  // ---- fused real-FFT split (adjust_DIT_impl, include/genFFT/generic/fft_dit_impl_generic.inl:27-61) ----
  // Every thread holds bins k = u + i*TN of Z (the L-point transform of the packed real signal).  The tile is
  // parked in shared memory so that each thread can fetch the partner bins Z[L-k]; then for all k in [0, L):
  //   X[k] = E - t*O,  E = (Z[k] + conj Z[L-k])/2,  O = (Z[k] - conj Z[L-k])/2,  t = i*W_{2L}^k = (sin, cos)
  // (the reference's formula for i < L/2; for the upper half it is algebraically the reference's
  // conj(E_j + t_j O_j)), plus X[L] = Re Z0 - Im Z0, and the conjugate mirror when !half.
  static __device__ __forceinline__ void dit_store(const PassParams& prm, const Tile& t, int c, int u, cpx<T> (&x)[P],
                                                   cpx<T>* smem) {
    V* sm = reinterpret_cast<V*>(smem + (size_t)c * PITCH);
    if (NST > 1) __syncthreads();  // the last gather of the exchange buffer is complete
#pragma unroll
    for (int i = 0; i < P; i++) {
      V v;
      v.x = x[i].x;
      v.y = x[i].y;
      sm[pad_idx(u + i * TN)] = v;
    }
    __syncthreads();
    const uint32_t col = t.col0 + c;
    if (col >= (uint32_t)prm.ncols) return;
    V* dst = reinterpret_cast<V*>(prm.out) + (long long)col * prm.out_stride_c;
    const V* tw = reinterpret_cast<const V*>(prm.dit_tw);
    auto bins = [&](auto full) {  // two copies of the loop: the half-spectrum one carries no mirror-store code
#pragma unroll
      for (int i = 0; i < P; i++) {
        const int k = u + i * TN;
        const int kp = (L - k) & (L - 1);
        const V zp = sm[pad_idx(kp)];
        const V w = __ldg(tw + k);
        V f = split_bin(x[i], zp, cpx<T>(w.x, w.y));
        if (k == 0) {
          f.y = T(0);  // exactly real, as in the reference (F[0] = zeroval)
          V nyq;
          nyq.x = x[i].x - x[i].y;
          nyq.y = T(0);
          dst[L] = nyq;
        }
        dst[k] = f;
        if constexpr (decltype(full)::value) {
          if (k != 0) {
            V g;
            g.x = f.x;
            g.y = -f.y;
            dst[2 * L - k] = g;
          }
        }
      }
    };
    if (prm.dit_half) bins(std::false_type());
    else bins(std::true_type());
  }

  // ---- fused real-FFT split for the last pass of a multi-pass transform (see M_COLTWDIT) ----
  // Thread (c, u) holds bins q = p + k*Ns, k = u + i*TN, of column p; the partner bin M - q lives in column Ns - p
  // (lane-column C-1-c of the same tile) at k' = L-1-k; column 0 and column Ns/2 pair with themselves.
  static __device__ __forceinline__ void pair_dit_store(const PassParams& prm, const Tile& t, int c, int u,
                                                        cpx<T> (&x)[P], cpx<T>* smem) {
    if (NST > 1) __syncthreads();
    {
      V* sm = reinterpret_cast<V*>(smem + (size_t)c * PITCH);
#pragma unroll
      for (int i = 0; i < P; i++) {
        V v;
        v.x = x[i].x;
        v.y = x[i].y;
        sm[pad_idx(u + i * TN)] = v;
      }
    }
    __syncthreads();
    bool valid;
    const uint32_t p = column_of(prm, t, c, valid);
    if (!valid) return;
    const uint32_t Ns = (uint32_t)prm.ncols;
    const bool self = (p == 0) || (p == Ns / 2);
    const int cp = self ? c : C - 1 - c;
    const V* smp = reinterpret_cast<const V*>(smem + (size_t)cp * PITCH);
    V* dst = reinterpret_cast<V*>(prm.out) + t.out_off;
    const V av = __ldg(reinterpret_cast<const V*>(prm.dit_a) + p);
    const cpx<T> a(av.x, av.y);
    const V* tw = reinterpret_cast<const V*>(prm.dit_tw);
    const long long M = (long long)Ns * L;
    auto split = [](const cpx<T>& z, const V& zp, const cpx<T>& w) { return split_bin(z, zp, w); };
    if (p != 0) {
      // Every column but the self-paired column 0 (one lane of one tile per transform): q = p + k Ns is never 0, the
      // partner index is L-1-k, and all addresses of the thread's 16 bins are a base plus a multiple of a fixed
      // step -- the output with a 32-bit stride (Ns < 2^32), the twiddles and the partner slots with immediates.
      const V* twu = tw + u;
      const V* smu = smp + pad_idx(L - 1 - u);
      V* dq = dst + p + (long long)u * Ns;
      V* dm = dst + (2 * M - p - (long long)u * Ns);
      auto bins = [&](auto full) {  // two copies of the loop: the half-spectrum one carries no mirror-store code
#pragma unroll
        for (int i = 0; i < P; i++) {
          V zp;
          if constexpr (TN >= PADN) {
            zp = *(smu - i * (TN + TN / PADN));  // pad_idx(L-1-u - i*TN) = pad_idx(L-1-u) - i*(TN + TN/PADN)
          } else {
            zp = smp[pad_idx(L - 1 - u - i * TN)];
          }
          const V bv = __ldg(twu + i * TN);
          const cpx<T> w = cmul(a, cpx<T>(bv.x, bv.y));  // W_n^q
          const V f = split(x[i], zp, w);
          st_strided<CO_DEFAULT>(dq, Ns, (uint32_t)(i * TN), f);
          if constexpr (decltype(full)::value) {
            V g;
            g.x = f.x;
            g.y = -f.y;
            *(dm - (long long)Ns * (i * TN)) = g;
          }
        }
      };
      if (prm.dit_half) bins(std::false_type());
      else bins(std::true_type());
      return;
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
      const int k = u + i * TN;
      const int kp = (L - k) & (L - 1);
      const V zp = smp[pad_idx(kp)];
      const V bv = __ldg(tw + k);
      const cpx<T> w = cmul(a, cpx<T>(bv.x, bv.y));  // W_n^q
      V f = split(x[i], zp, w);
      const long long q = (long long)k * Ns;
      if (q == 0) {
        f.y = T(0);
        V nyq;
        nyq.x = x[i].x - x[i].y;
        nyq.y = T(0);
        dst[M] = nyq;
      }
      dst[q] = f;
      if (!prm.dit_half && q != 0) {
        V g;
        g.x = f.x;
        g.y = -f.y;
        dst[2 * M - q] = g;
      }
    }
  }

  // One tile of a pass chain (chain_kernel.cuh): the CTA transforms tile `tile` of the group whose element offsets
  // (in_g, out_g) and twiddle-column offset p_g are added to the pass's group-0 addressing.
  template <int LDOP, int STOP>
  static __device__ __forceinline__ void tile_once(const PassParams& prm, cpx<T>* smem, uint32_t tile, long long in_g,
                                                   long long out_g, uint32_t p_g) {
    static_assert(!TMA, "chains run the plain tile modes");
    const int tid = threadIdx.x;
    const bool ld_b = GEN ? prm.map_load != 0 : ROWLIKE;
    const bool st_b = GEN ? prm.map_store != 0 : (ROWLIKE || MODE == M_FIRST);
    const int c_ld = ld_b ? tid / TN : tid % C;
    const int u_ld = ld_b ? tid % TN : tid / C;
    const int c_st = st_b ? tid / TN : tid % C;
    const int u_st = st_b ? tid % TN : tid / C;
    cpx<T> x[P];
    Tile t = decode(prm, tile);
    t.in_off += in_g;
    t.out_off += out_g;
    t.p_base += p_g;
    if constexpr (PEER) t.grp_col = (uint32_t)out_g;  // column passes: the group offset is a column count
    load<LDOP>(prm, t, c_ld, u_ld, x);
    int c = c_ld, u = u_ld;
    run_stages<0>(prm, x, smem, c, u, c_st, u_st);
    if constexpr (DIT) {
      dit_store(prm, t, c, u, x, smem);
    } else if constexpr (PAIR) {
      pair_dit_store(prm, t, c, u, x, smem);
    } else {
      store<STOP>(prm, t, c, u, x);
    }
  }

  static __device__ __forceinline__ void body(const PassParams& prm, cpx<T>* smem) {
    const int tid = threadIdx.x;
    const bool ld_b = GEN ? prm.map_load != 0 : ROWLIKE;
    const bool st_b = GEN ? prm.map_store != 0 : (ROWLIKE || MODE == M_FIRST);
    const int c_ld = ld_b ? tid / TN : tid % C;
    const int u_ld = ld_b ? tid % TN : tid / C;
    const int c_st = st_b ? tid / TN : tid % C;
    const int u_st = st_b ? tid % TN : tid / C;
    cpx<T> x[P];
    if constexpr (TMA) {
      // Persistent CTA: tile k+1 is fetched by one cp.async.bulk per sequence into `inbuf` while tile k is being
      // transformed.  `inbuf` is free again once every thread has copied its points to registers, which the first
      // __syncthreads of the stage pipeline guarantees.
      cpx<T>* inbuf = reinterpret_cast<cpx<T>*>(reinterpret_cast<unsigned char*>(smem) + INBUF_OFFSET);
      uint64_t* bar = reinterpret_cast<uint64_t*>(inbuf + (size_t)L * C);
      const V* gin = reinterpret_cast<const V*>(prm.in);
      auto issue = [&](uint32_t tile) {
        const uint32_t col0 = tile * C;
        const uint32_t nvalid = min((uint32_t)C, (uint32_t)prm.ncols - col0);
        mbar_arrive_expect_tx(bar, nvalid * (uint32_t)(L * sizeof(cpx<T>)));
        for (uint32_t cc = 0; cc < nvalid; cc++)
          bulk_load(inbuf + (size_t)cc * L, gin + (long long)(col0 + cc) * prm.in_stride_c, (uint32_t)(L * sizeof(cpx<T>)), bar);
      };
      if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
      }
      __syncthreads();
      const uint32_t K = prm.tiles_per_cta;
      const uint32_t step = K ? 1u : gridDim.x;
      const uint32_t tile_end = K ? min(prm.ntiles, (blockIdx.x + 1u) * K) : prm.ntiles;
      uint32_t tile = K ? blockIdx.x * K : blockIdx.x;
      if (tid == 0 && tile < tile_end) issue(tile);
      uint32_t parity = 0;
      for (; tile < tile_end; tile += step) {
        Tile t = decode(prm, tile);
        mbar_wait(bar, parity);
        parity ^= 1;
        {
          const V* src = reinterpret_cast<const V*>(inbuf + (size_t)c_ld * L) + u_ld;
#pragma unroll
          for (int i = 0; i < P; i++) {
            V v = src[i * TN];
            x[i] = cpx<T>(v.x, INV ? -v.y : v.y);
          }
        }
        int c = c_ld, u = u_ld;
        // stage 0 + first barrier, then the prefetch of the next tile, then the remaining stages
        stage<0>(prm, x, smem + (size_t)c * PITCH, u);
        __syncthreads();
        if (tid == 0 && tile + step < tile_end) issue(tile + step);
        if constexpr (NST > 1) {
          if constexpr (NST == 2) {
            c = c_st;
            u = u_st;
          }
          gather(x, smem + (size_t)c * PITCH, u);
          if constexpr (NST > 2) __syncthreads();
          run_stages<1>(prm, x, smem, c, u, c_st, u_st);
        }
        store(prm, t, c, u, x);
        if (NST > 1) __syncthreads();
      }
    } else {
      const uint32_t K = prm.tiles_per_cta;
      const uint32_t step = K ? 1u : gridDim.x;
      const uint32_t tile_end = K ? min(prm.ntiles, (blockIdx.x + 1u) * K) : prm.ntiles;
      for (uint32_t tile = K ? blockIdx.x * K : blockIdx.x; tile < tile_end; tile += step) {
        Tile t = decode(prm, tile);
        load(prm, t, c_ld, u_ld, x);
        int c = c_ld, u = u_ld;
        run_stages<0>(prm, x, smem, c, u, c_st, u_st);
        if constexpr (DIT) {
          dit_store(prm, t, c, u, x, smem);
          __syncthreads();
        } else if constexpr (PAIR) {
          pair_dit_store(prm, t, c, u, x, smem);
          __syncthreads();
        } else {
          store(prm, t, c, u, x);
          if (NST > 1) __syncthreads();  // next tile's first scatter must not overtake this tile's gathers
        }
      }
    }
  }
};

// resident-thread target per SM: 1024 (64 registers) in float, 512 (128 registers) in double, where a
// thread's 16 complex points alone are 64 registers
template <typename T, int THREADS, int P = 16>
constexpr int min_blocks() {
#ifndef GENFFT_F64_TARGET_THREADS
#define GENFFT_F64_TARGET_THREADS 512
#endif
#ifndef GENFFT_F32_TARGET_THREADS
#define GENFFT_F32_TARGET_THREADS 1024
#endif
  // (8 double-complex points per thread at 64 registers, i.e. twice the resident warps, were measured in round 2:
  // C3 311 us against 258 us for 16 points at 128 registers -- profiles/r02_c3_8_points_per_thread_64_registers.log)
  constexpr int target = sizeof(T) == 4 ? GENFFT_F32_TARGET_THREADS : GENFFT_F64_TARGET_THREADS;
  return THREADS >= target ? 1 : target / THREADS;
}

template <typename T, int L, int P, int C, int MODE, bool INV>
__global__ void __launch_bounds__(TileKernel<T, L, P, C, MODE, INV>::THREADS,
                                  min_blocks<T, TileKernel<T, L, P, C, MODE, INV>::THREADS, P>())
fft_tile_kernel(const __grid_constant__ PassParams prm) {
  GENFFT_DYN_SMEM(smem_raw);
#ifndef GENFFT_EMU
  // Programmatic dependent launch: when this grid was launched with programmatic stream serialization, it may become
  // resident while the previous kernel of the stream is still running -- launch_dependents lets the NEXT kernel do the
  // same as soon as every CTA of this grid has started, wait blocks until the previous grid has completed and its
  // memory is visible.  Nothing above the wait touches memory a kernel writes, so back-to-back small transforms (a
  // forward + inverse pair of N = 1024 is two single-CTA launches) no longer pay a full launch latency each.  Both
  // instructions are no-ops in a normally launched grid.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  TileKernel<T, L, P, C, MODE, INV>::body(prm, reinterpret_cast<cpx<T>*>(smem_raw));
}

}  // namespace genfft_cuda
