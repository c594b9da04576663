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

// out[b*out_dist + r*out_stride + c] = in[b*in_dist + r*in_stride + c]  (complex elements)
struct CopyParams {
  const void* in;
  void* out;
  long long in_stride, out_stride, in_dist, out_dist;
  long long rows, cols;
  int in_real;  // widen real scalars to complex
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
        v.y = T(0);
      } else {
        v = reinterpret_cast<const V*>(p.in)[b * p.in_dist + r * p.in_stride + c];
      }
      reinterpret_cast<V*>(p.out)[b * p.out_dist + r * p.out_stride + c] = v;
    }
  }
}

}  // namespace genfft_cuda
