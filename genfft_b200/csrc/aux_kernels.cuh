// Small non-butterfly kernels: the real-FFT split ("DIT") as a stand-alone pass and a strided copy.
#pragma once
#include <cstdint>
#include "radix.cuh"

namespace genfft_cuda {

// Real-FFT post-process.  Restates adjust_DIT_impl
// (include/genFFT/generic/fft_dit_impl_generic.inl:27-61; x86 versions
// include/genFFT/x86/fft_dit_impl_x86.inl:30-177): Z is the N/2-point complex FFT of the real signal
// packed as (even, odd); F is its N-point spectrum.  Bins i and N/2-i are produced together, so the
// pass is safe in place (F == Z) -- the property RealFFT<T>::forward relies on (FFTReal.h:211).
// With half != 0 exactly N/2+1 bins are written and nothing beyond (test/fft_test_impl.h:102-105).
struct DitParams {
  const void* in;
  void* out;
  long long in_dist, out_dist;  // complex elements between transforms
  int n;                        // real transform size N
  int half;
  int batch;
  int in_is_real_scalar;        // N == 1: input is one real scalar
  const void* tw_hi;            // two-level table of W_N^e = (cos, -sin)(2*pi*e/N)
  const void* tw_lo;
  int tw_shift;
};

template <typename T>
__global__ void __launch_bounds__(256) dit_kernel(const __grid_constant__ DitParams p) {
  using V = typename vec2<T>::type;
  const int N = p.n;
  const int quarter = N >> 2;
  for (long long b = blockIdx.y; b < p.batch; b += gridDim.y) {
  const V* Z = reinterpret_cast<const V*>(p.in) + b * p.in_dist;
  V* F = reinterpret_cast<V*>(p.out) + b * p.out_dist;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= quarter; i += gridDim.x * blockDim.x) {
    if (i == 0) {
      if (p.in_is_real_scalar) {  // RealFFT n == 1 (FFTReal.h:206-207)
        V r;
        r.x = reinterpret_cast<const T*>(p.in)[b * p.in_dist];
        r.y = T(0);
        F[0] = r;
        continue;
      }
      V z0 = Z[0];
      V a, c;
      a.x = z0.x + z0.y;
      a.y = T(0);
      c.x = z0.x - z0.y;
      c.y = T(0);
      F[0] = a;
      if (N >= 2) F[N >> 1] = c;
    } else if (i == quarter) {
      V z = Z[i];
      V f;
      f.x = z.x;
      f.y = -z.y;
      F[i] = f;
      if (!p.half) F[N - i] = z;
    } else {
      const int j = (N >> 1) - i;
      V zi = Z[i], zj = Z[j];
      // E = (Z[i] + conj Z[j]) / 2, O = (Z[i] - conj Z[j]) / 2
      const T er = (zi.x + zj.x) * T(0.5), ei = (zi.y - zj.y) * T(0.5);
      const T orr = (zi.x - zj.x) * T(0.5), oi = (zi.y + zj.y) * T(0.5);
      const uint32_t e = (uint32_t)i;
      V wh = __ldg(reinterpret_cast<const V*>(p.tw_hi) + (e >> p.tw_shift));
      V wl = __ldg(reinterpret_cast<const V*>(p.tw_lo) + (e & ((1u << p.tw_shift) - 1u)));
      cpx<T> w = cmul(cpx<T>(wh.x, wh.y), cpx<T>(wl.x, wl.y));  // (cos, -sin)
      const T tr = -w.y, ti = w.x;                              // t = sin + i cos
      const T pr = tr * orr - ti * oi, pi = tr * oi + ti * orr; // t * O
      V fi, fj;
      fi.x = er - pr;
      fi.y = ei - pi;
      fj.x = er + pr;
      fj.y = -(ei + pi);
      F[i] = fi;
      F[j] = fj;
      if (!p.half) {
        V ci, cj;
        ci.x = fi.x;
        ci.y = -fi.y;
        cj.x = fj.x;
        cj.y = -fj.y;
        F[N - i] = ci;
        F[N - j] = cj;
      }
    }
  }
  }
}

// separate_2x_real_FFT (include/genFFT/FFTReal.h:35-66): Fz is the N-point transform of x + i*y with x, y real;
// recovers Fx and Fy.  Bins i and N-i are handled by one thread, so out1 or out2 may alias in.
struct SeparateParams {
  const void* in;
  void* out1;
  void* out2;
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256) separate_kernel(const __grid_constant__ SeparateParams p) {
  using V = typename vec2<T>::type;
  const V* Fz = reinterpret_cast<const V*>(p.in);
  V* Fx = reinterpret_cast<V*>(p.out1);
  V* Fy = reinterpret_cast<V*>(p.out2);
  const int N = p.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N / 2; i += gridDim.x * blockDim.x) {
    if (i == 0) {
      V z = Fz[0], a, b;
      a.x = z.x;
      a.y = T(0);
      b.x = z.y;
      b.y = T(0);
      Fx[0] = a;
      Fy[0] = b;
      continue;
    }
    const int k = N - i;
    const V zi = Fz[i], zk = Fz[k];
    V x, y, xc, yc;
    x.x = (zi.x + zk.x) * T(0.5);
    x.y = (zi.y - zk.y) * T(0.5);
    y.x = (zk.y + zi.y) * T(0.5);
    y.y = (zk.x - zi.x) * T(0.5);
    xc.x = x.x;
    xc.y = -x.y;
    yc.x = y.x;
    yc.y = -y.y;
    Fx[i] = x;
    Fy[i] = y;
    Fx[k] = xc;
    Fy[k] = yc;
  }
}

// Hermitian completion of a real image's spectrum (RealFFT2D<T>::forward, include/genFFT/FFTReal.h:90-103):
// out[(H-i)%H][j] = conj(out[i][W-j]) for W/2 < j < W.
struct MirrorParams {
  void* data;
  long long stride;
  int w, h;
};

template <typename T>
__global__ void __launch_bounds__(256) mirror2d_kernel(const __grid_constant__ MirrorParams p) {
  using V = typename vec2<T>::type;
  V* d = reinterpret_cast<V*>(p.data);
  const int first = p.w / 2 + 1;
  for (int i = blockIdx.y; i < p.h; i += gridDim.y) {
    const int k = i ? p.h - i : 0;
    for (int j = first + blockIdx.x * blockDim.x + threadIdx.x; j < p.w; j += gridDim.x * blockDim.x) {
      V v = d[(long long)i * p.stride + (p.w - j)];
      v.y = -v.y;
      d[(long long)k * p.stride + j] = v;
    }
  }
}

// Half-spectrum inverse of a real transform (not in the reference: README.txt:51-52 "Maybe").  Pre-process of the
// N/2+1 bins X into the N/2-point packed spectrum Z' = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) conj(W_N^k),
// M = N/2, whose unscaled inverse complex transform is N * (x[2m] + i x[2m+1]).
struct C2rParams {
  const void* in;
  void* out;
  long long in_dist, out_dist;  // complex elements
  int n;
  int batch;
  const void* tw_hi;
  const void* tw_lo;
  int tw_shift;
};

template <typename T>
__global__ void __launch_bounds__(256) c2r_pre_kernel(const __grid_constant__ C2rParams p) {
  using V = typename vec2<T>::type;
  const int M = p.n >> 1;
  for (long long b = blockIdx.y; b < p.batch; b += gridDim.y) {
    const V* X = reinterpret_cast<const V*>(p.in) + b * p.in_dist;
    V* Z = reinterpret_cast<V*>(p.out) + b * p.out_dist;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) {
      const V xk = X[k], xm = X[M - k];
      const T ax = xk.x + xm.x, ay = xk.y - xm.y;
      const T dx = xk.x - xm.x, dy = xk.y + xm.y;
      const uint32_t e = (uint32_t)k;
      V wh = __ldg(reinterpret_cast<const V*>(p.tw_hi) + (e >> p.tw_shift));
      V wl = __ldg(reinterpret_cast<const V*>(p.tw_lo) + (e & ((1u << p.tw_shift) - 1u)));
      const cpx<T> w = cmul(cpx<T>(wh.x, wh.y), cpx<T>(wl.x, wl.y));  // W_N^k
      const T tx = dx * w.x + dy * w.y, ty = dy * w.x - dx * w.y;    // d * conj(w)
      V z;
      z.x = ax - ty;
      z.y = ay + tx;
      Z[k] = z;
    }
  }
}

// out[b*out_dist + r*out_stride + c] = in[b*in_dist + r*in_stride + c]  (complex elements)
struct CopyParams {
  const void* in;
  void* out;
  long long in_stride, out_stride, in_dist, out_dist;
  long long rows, cols;
  int in_real;  // 1: widen real scalars to complex; 2: real part from `in`, imaginary part from `in2` (forward_2x)
  const void* in2;
  long long in2_stride;
};

template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(const __grid_constant__ CopyParams p) {
  using V = typename vec2<T>::type;
  const long long b = blockIdx.z;
  for (long long r = blockIdx.y; r < p.rows; r += gridDim.y) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < p.cols;
         c += (long long)gridDim.x * blockDim.x) {
      V v;
      if (p.in_real) {
        v.x = reinterpret_cast<const T*>(p.in)[b * p.in_dist + r * p.in_stride + c];
        v.y = p.in_real == 2 ? reinterpret_cast<const T*>(p.in2)[r * p.in2_stride + c] : T(0);
      } else {
        v = reinterpret_cast<const V*>(p.in)[b * p.in_dist + r * p.in_stride + c];
      }
      reinterpret_cast<V*>(p.out)[b * p.out_dist + r * p.out_stride + c] = v;
    }
  }
}

// First global transpose of the distributed four-step 1D transform as ONE launch: the (rows x width) slab of this rank
// goes, column block by column block, into the peers' (H x width/P) blocks at rows [row0, row0 + rows):
//   peer[c / wp][(row0 + r) * wp + c % wp] = in[r * in_stride + c],   wp = width / nparts (a power of two).
// 16-byte accesses (two float2 or one double2 per thread), both sides coalesced; NVLink stores for the remote blocks.
struct ScatterParams {
  const void* in;
  void* peer[8];
  long long in_stride;
  long long rows, width, row0;
  int wp_log2;
};

template <typename T>
__global__ void __launch_bounds__(256) scatter_cols_kernel(const __grid_constant__ ScatterParams p) {
  constexpr int EPV = sizeof(T) == 4 ? 2 : 1;  // complex elements per 16-byte vector
  const long long vec_per_row = p.width / EPV;
  const long long total = p.rows * vec_per_row;
  const long long wp = 1LL << p.wp_log2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / vec_per_row;
    const long long c = (idx - r * vec_per_row) * EPV;
    const int g = (int)(c >> p.wp_log2);
    const long long x = c & (wp - 1);
    const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(p.in) + (size_t)(r * p.in_stride + c) * (2 * sizeof(T)));
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(p.peer[g]) + (size_t)((p.row0 + r) * wp + x) * (2 * sizeof(T))) = v;
  }
}

// Inter-step twiddle of the distributed four-step 1D transform (N = H*W points seen as an H x W row-major matrix whose
// row slabs live on different GPUs): after the length-H column transforms, element (kr, c) is multiplied by
// W_N^(kr*c) before the length-W row transforms.  In place on a (rows x cols) slab whose first row is global row row0.
struct Twiddle2dParams {
  void* data;
  long long stride;  // complex elements between rows
  long long rows, cols, row0;
  const void* tw_hi;  // two-level table of W_N^e = (cos, -sin)(2*pi*e/N)
  const void* tw_lo;
  int tw_shift;
  int inverse;
};

template <typename T>
__global__ void __launch_bounds__(256) twiddle2d_kernel(const __grid_constant__ Twiddle2dParams p) {
  using V = typename vec2<T>::type;
  V* d = reinterpret_cast<V*>(p.data);
  const unsigned long long lo_mask = (1ull << p.tw_shift) - 1ull;
  for (long long r = blockIdx.y; r < p.rows; r += gridDim.y) {
    const unsigned long long kr = (unsigned long long)(p.row0 + r);
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < p.cols;
         c += (long long)gridDim.x * blockDim.x) {
      const unsigned long long e = kr * (unsigned long long)c;  // < H*W = N
      const V wh = __ldg(reinterpret_cast<const V*>(p.tw_hi) + (e >> p.tw_shift));
      const V wl = __ldg(reinterpret_cast<const V*>(p.tw_lo) + (e & lo_mask));
      cpx<T> w = cmul(cpx<T>(wh.x, wh.y), cpx<T>(wl.x, wl.y));
      if (p.inverse) w.y = -w.y;
      const V v = d[r * p.stride + c];
      const cpx<T> z = cmul(cpx<T>(v.x, v.y), w);
      V o;
      o.x = z.x;
      o.y = z.y;
      d[r * p.stride + c] = o;
    }
  }
}

// out[c*out_stride + r] = in[r*in_stride + c]: 32 x 32 tiles through padded shared memory, both sides coalesced.
// Last step of the distributed four-step 1D transform when the result is wanted in natural order.
struct TransposeParams {
  const void* in;
  void* out;
  long long in_stride, out_stride;
  long long rows, cols;  // of `in`
};

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const __grid_constant__ TransposeParams p) {
  using V = typename vec2<T>::type;
  __shared__ V tile[32][33];
  const V* in = reinterpret_cast<const V*>(p.in);
  V* out = reinterpret_cast<V*>(p.out);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  const long long tiles_c = (p.cols + 31) / 32, tiles_r = (p.rows + 31) / 32;
  for (long long t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const long long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const long long r = r0 + ty + j, c = c0 + tx;
      if (r < p.rows && c < p.cols) tile[ty + j][tx] = in[r * p.in_stride + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const long long c = c0 + ty + j, r = r0 + tx;
      if (r < p.rows && c < p.cols) out[c * p.out_stride + r] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

// Cross-GPU barrier over peer memory (distributed 2D, p2p transport): every rank owns an array of `world` epoch
// flags that all peers have mapped through CUDA IPC.  Thread r of the single CTA publishes this rank's arrival in
// peer r's array, then waits until peer r's arrival shows up in the local array.  Stream order puts the kernel after
// the pass whose NVLink stores it fences (a kernel boundary completes them system-wide) and before the pass that
// reads what the peers stored, so a rank leaves the barrier only when every peer's earlier kernels are done --
// in a few microseconds, without a collective library call on the critical path.
struct PeerBarrierParams {
  uint32_t* peer_flags[8];  // [kMaxPeers]: peer r's flag array (the own array at index `rank`)
  int rank, world;
  uint32_t epoch;  // strictly increasing per barrier
};

#ifndef GENFFT_EMU
static __global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerBarrierParams p) {
  const int r = threadIdx.x;
  if (r >= p.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[r] + p.rank), "r"(p.epoch) : "memory");
  const uint32_t* mine = p.peer_flags[p.rank] + r;
  uint32_t v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - p.epoch) >= 0) break;
    __nanosleep(100);
    if (++spins > (1u << 27)) __trap();  // ~15 s: a peer that never arrives must fail loudly, never hang the GPU
  }
}
#else
// emulator: ranks run one after another in one process, so the barrier only publishes the arrival
inline void peer_barrier_kernel(const PeerBarrierParams p) {
  const int r = threadIdx.x;
  if (r < p.world) p.peer_flags[r][p.rank] = p.epoch;
}
#endif

}  // namespace genfft_cuda
