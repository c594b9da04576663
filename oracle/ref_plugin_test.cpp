// TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN classes (include/genFFT/fft.h, FFTReal.h -- compiled from /root/reference where they lie,
// native mode) instantiated with the CUDA factories of include/genfft_cuda/backend.h through the reference's
// factory template parameters (fft.h:56,117,173; FFTReal.h:188), checked against the same classes with the
// reference's native AVX2/FMA back-end in the same process: this is the literal drop-in at the reference's
// plug-in boundary.  Built by `make -C oracle plugin` into oracle/_ref/ref_plugin_test (GPU box runs the
// prebuilt binary; it needs libgenfft_cuda.so and a B200).
#include <genFFT/fft.h>
#include <genfft_cuda/backend.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

#include "test_util.h"  // the reference's DummyData / FFT_Eps (test/test_util.h)

static int g_fail = 0;

template <class T>
static double rel_l2(const std::vector<std::complex<T>>& a, const std::vector<std::complex<T>>& b, size_t n) {
  double num = 0, den = 0;
  for (size_t i = 0; i < n; i++) {
    num += std::norm(std::complex<double>(a[i]) - std::complex<double>(b[i]));
    den += std::norm(std::complex<double>(b[i]));
  }
  return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

template <class T> double tol(double n) { return (sizeof(T) == 4 ? 1e-6 : 1e-14) * std::max(1.0, std::log2(n)); }

#define EXPECT(cond, ...)                                                      \
  do {                                                                         \
    if (!(cond)) { std::printf("FAIL: "); std::printf(__VA_ARGS__); std::printf("\n"); g_fail++; } \
  } while (0)

template <class T>
void run() {
  using CudaFFT = genfft::FFT<T, genfft::impl_cuda::GetImpl>;
  using CudaVert = genfft::FFTVert<T, genfft::impl_cuda::GetVertImpl>;
  using CudaDIT = genfft::DIT<T, genfft::impl_cuda::GetDITImpl>;
  using CudaReal = genfft::RealFFT<T, genfft::impl_cuda::GetImpl, genfft::impl_cuda::GetDITImpl>;
  for (int n = 1; n <= (1 << 18); n += n) {
    std::vector<std::complex<T>> in(n), a(n), b(n);
    DummyData(in, false);
    for (int inv = 0; inv < 2; inv++) {
      CudaFFT cu(n);
      genfft::FFT<T> cpu(n);
      if (inv) { cu.template transform<true>(a.data(), in.data()); cpu.template transform<true>(b.data(), in.data()); }
      else     { cu.template transform<false>(a.data(), in.data()); cpu.template transform<false>(b.data(), in.data()); }
      double e = rel_l2(a, b, n);
      EXPECT(e <= tol<T>(n), "FFT<%s> n=%d inv=%d rel-L2 %.3g", sizeof(T) == 4 ? "float" : "double", n, inv, e);
    }
    for (int half = 0; half < 2; half++) {
      std::vector<T> rin(n);
      DummyData(rin);
      std::vector<std::complex<T>> ra(n, std::complex<T>(43, 21)), rb(n, std::complex<T>(43, 21));
      CudaReal cu(n);
      genfft::RealFFT<T> cpu(n);
      cu.forward(ra.data(), rin.data(), half != 0);
      cpu.forward(rb.data(), rin.data(), half != 0);
      size_t lim = n == 1 ? 1 : (half ? n / 2 + 1 : n);
      double e = rel_l2(ra, rb, lim);
      EXPECT(e <= tol<T>(n), "RealFFT n=%d half=%d rel-L2 %.3g", n, half, e);
      for (size_t i = lim; i < (size_t)n; i++) EXPECT(ra[i] == std::complex<T>(43, 21), "RealFFT wrote past n/2+1 (n=%d)", n);
    }
    if (n >= 2 && n <= (1 << 16)) {
      std::vector<std::complex<T>> z(n), da(n), db(n);
      DummyData(z, false);
      CudaDIT cu(n);
      genfft::DIT<T> cpu(n);
      cu.apply(da.data(), z.data(), false);
      cpu.apply(db.data(), z.data(), false);
      double e = rel_l2(da, db, n);
      EXPECT(e <= tol<T>(n), "DIT n=%d rel-L2 %.3g", n, e);
    }
  }
  const int vert[][2] = {{2, 3}, {4, 7}, {8, 33}, {16, 47}, {32, 63}, {128, 767}, {256, 999}, {1024, 31}};
  for (auto& v : vert) {
    const int n = v[0], cols = v[1];
    std::vector<std::complex<T>> in((size_t)n * cols), a((size_t)n * cols), b((size_t)n * cols);
    DummyData(in, false);
    CudaVert cu(n);
    genfft::FFTVert<T> cpu(n);
    cu.template transform<false>(a.data(), cols, in.data(), cols, cols);
    cpu.template transform<false>(b.data(), cols, in.data(), cols, cols);
    double e = rel_l2(a, b, a.size());
    EXPECT(e <= tol<T>(n), "FFTVert n=%d cols=%d rel-L2 %.3g", n, cols, e);
  }
}

int main() {
  run<float>();
  run<double>();
  std::printf("%s (%d failures)\n", g_fail ? "FAILED" : "PASSED", g_fail);
  return g_fail ? 1 : 0;
}
