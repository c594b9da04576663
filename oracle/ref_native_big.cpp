// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// The reference's size switch stops at 2^23 (include/genFFT/x86/fft_double_impl_x86.inl:356-389,
// fft_float_impl_x86.inl:464-497) but BASELINE.json config C3 is N = 2^24 in double.
// The reference's public classes take the implementation factory as a template
// parameter (include/genFFT/fft.h:56), so its own kernels FFTDouble<N> / FFTFloat<N>
// (native back-end, include/genFFT/x86/fft_x86_native.h:33-39) can be instantiated at
// 2^24 / 2^25 through that hook without touching the reference.  This TU is
// compiled in *native* mode (-mavx2 -mfma), separately from ref_driver.cpp which is
// in dispatch mode; the two modes may not meet in one TU (FFTBackend.h:30-33).
#include <genFFT/fft.h>

#include <chrono>
#include <complex>
#include <random>
#include <vector>

namespace {

template <int LOG2>
std::shared_ptr<genfft::impl::FFTBase<double>> BigDouble(int, double) {
  return genfft::impl::FFTLevel<(1 << LOG2), double,
                                genfft::impl_native::FFTDouble<(1 << LOG2)>>::GetInstance();
}
template <int LOG2>
std::shared_ptr<genfft::impl::FFTBase<float>> BigFloat(int, float) {
  return genfft::impl::FFTLevel<(1 << LOG2), float,
                                genfft::impl_native::FFTFloat<(1 << LOG2)>>::GetInstance();
}

template <class FFT, class T>
int run(T *out, const T *in, int n, int inv) {
  FFT fft(n);
  if (inv)
    fft.template transform<true>((std::complex<T> *)out, (const std::complex<T> *)in);
  else
    fft.template transform<false>((std::complex<T> *)out, (const std::complex<T> *)in);
  return 0;
}

}  // namespace

extern "C" {

// supported: n == 2^24 (double, float) and 2^25 (float)
int genfft_ref_big_c2c_f64(double *out, const double *in, long n, int inv) {
  if (out == in) return 1;
  if (n == (1L << 24)) return run<genfft::FFT<double, BigDouble<24>>, double>(out, in, (int)n, inv);
  return 1;
}
int genfft_ref_big_c2c_f32(float *out, const float *in, long n, int inv) {
  if (out == in) return 1;
  if (n == (1L << 24)) return run<genfft::FFT<float, BigFloat<24>>, float>(out, in, (int)n, inv);
  return 1;
}

// fft_bench.cpp-style timing (std::chrono around the transform, plan built before) of one
// forward transform of n = 2^24 in double on one core; returns seconds for `count` transforms.
double genfft_ref_bench_big_c2c_f64(long n, long count) {
  if (n != (1L << 24)) return -1;
  std::mt19937_64 rng;
  std::uniform_real_distribution<double> dist(-1, 1);
  genfft::FFT<double, BigDouble<24>> fft((int)n);
  std::vector<std::complex<double>> in(n), out(n);
  double acc = 0;
  for (long k = 0; k < count; k++) {
    for (auto &c : in) c = {dist(rng), dist(rng)};
    auto t0 = std::chrono::steady_clock::now();
    fft.transform<false>(out.data(), in.data());
    acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return acc;
}

}  // extern "C"
