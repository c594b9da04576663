// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI driver around the UNMODIFIED reference (mzient/genFFT) compiled in place
// from /root/reference by oracle/Makefile into oracle/_ref/libgenfft_ref.so.
// It is used by tests/ to pin the C restatement (oracle/genfft_oracle.c), to
// generate tests/golden/ fixtures, and by bench.py as the `cpu_baseline` /
// `--impl reference` arm.  Nothing in genfft_b200/ may load this library.
//
// The reference is driven through its own public classes in *dispatch* mode
// (include/genFFT/fft_dispatch.h:27-36, src/fft_x86_dispatch.cpp:41-139), i.e.
// the "best CPU dispatch build" north_star names.  The generic scalar back-end
// (src/fft_generic.cpp:8-10) is exposed beside it so that the plain-C oracle can
// be pinned bit-for-bit against reference code that uses no SIMD/FMA.
#include <genFFT/fft_dispatch.h>
#include <genFFT/x86/x86_features.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

// the reference's own test helpers: DummyData (test/test_util.h:36-57),
// reference_impl::FFT_pow2 / FFT_pow2_vert / DFT (test/fft_ref_impl.h:37-135)
#include "test_util.h"

namespace genfft {
namespace impl_generic {
// defined by DISPATCH_ALL() in src/fft_generic.cpp:8-10 (src/dispatch_helper.h:25-30)
std::shared_ptr<impl::FFTBase<float>> GetDispatchImpl(int n, float);
std::shared_ptr<impl::FFTBase<double>> GetDispatchImpl(int n, double);
std::shared_ptr<impl::FFTVertBase<float>> GetVertDispatchImpl(int n, float);
std::shared_ptr<impl::FFTVertBase<double>> GetVertDispatchImpl(int n, double);
std::shared_ptr<impl::FFTDITBase<float>> GetDITDispatchImpl(int n, float);
std::shared_ptr<impl::FFTDITBase<double>> GetDITDispatchImpl(int n, double);
}  // namespace impl_generic
}  // namespace genfft

namespace {

// Release builds define NDEBUG, so the reference's size asserts vanish and an
// unsupported size falls off a non-void function (x86/fft_float_impl_x86.inl:494).
// Validate here instead.
bool valid_pow2(long n, int max_log2 = 23) {
  return n >= 1 && (n & (n - 1)) == 0 && n <= (1L << max_log2);
}

template <class T>
using cpx = std::complex<T>;

template <class T, bool generic>
struct Factories;
template <class T>
struct Factories<T, false> {
  using FFT = genfft::FFT<T>;
  using Vert = genfft::FFTVert<T>;
  using DIT = genfft::DIT<T>;
  using Real = genfft::RealFFT<T>;
};
template <class T>
struct Factories<T, true> {
  using FFT = genfft::FFT<T, genfft::impl_generic::GetDispatchImpl>;
  using Vert = genfft::FFTVert<T, genfft::impl_generic::GetVertDispatchImpl>;
  using DIT = genfft::DIT<T, genfft::impl_generic::GetDITDispatchImpl>;
  using Real = genfft::RealFFT<T, genfft::impl_generic::GetDispatchImpl,
                               genfft::impl_generic::GetDITDispatchImpl>;
};

template <class T, bool generic>
int c2c(T *out, const T *in, int n, int inverse) {
  if (!valid_pow2(n) || out == in) return 1;
  typename Factories<T, generic>::FFT fft(n);
  if (inverse)
    fft.template transform<true>((cpx<T> *)out, (const cpx<T> *)in);
  else
    fft.template transform<false>((cpx<T> *)out, (const cpx<T> *)in);
  return 0;
}

template <class T, bool generic>
int c2c_no_scramble(T *inout, int n, int inverse) {
  if (!valid_pow2(n)) return 1;
  typename Factories<T, generic>::FFT fft(n);
  if (inverse)
    fft.template transform_no_scramble<true>((cpx<T> *)inout);
  else
    fft.template transform_no_scramble<false>((cpx<T> *)inout);
  return 0;
}

template <class T, bool generic>
int r2c(T *out, const T *in, int n, int half) {
  if (!valid_pow2(n)) return 1;
  typename Factories<T, generic>::Real fft(n);
  fft.forward((cpx<T> *)out, in, half != 0);
  return 0;
}

template <class T, bool generic>
int dit(T *out, const T *in, int n, int half) {
  if (!valid_pow2(n)) return 1;
  typename Factories<T, generic>::DIT d(n);
  d.apply(out, in, half != 0);
  return 0;
}

template <class T, bool generic>
int vert(T *out, long out_stride, const T *in, long in_stride, int n, int cols, int inverse) {
  if (!valid_pow2(n) || out == in || cols < 0) return 1;
  typename Factories<T, generic>::Vert fft(n);
  if (inverse)
    fft.template transform<true>((cpx<T> *)out, out_stride, (const cpx<T> *)in, in_stride, cols);
  else
    fft.template transform<false>((cpx<T> *)out, out_stride, (const cpx<T> *)in, in_stride, cols);
  return 0;
}

template <class T>
int fft2d(T *out, long out_stride, const T *in, long in_stride, int width, int height, int inverse) {
  if (!valid_pow2(width) || !valid_pow2(height) || out == in) return 1;
  genfft::FFT2D<T> fft(width, height);
  if (inverse)
    fft.template transform<true>((cpx<T> *)out, out_stride, (const cpx<T> *)in, in_stride);
  else
    fft.template transform<false>((cpx<T> *)out, out_stride, (const cpx<T> *)in, in_stride);
  return 0;
}

template <class T>
int real_fft2d(T *out, int out_stride, const T *in, int in_stride, int width, int height) {
  if (!valid_pow2(width) || !valid_pow2(height)) return 1;
  genfft::RealFFT2D<T> fft(width, height);
  fft.forward((cpx<T> *)out, out_stride, in, in_stride);
  return 0;
}

// RealFFT2D<T>::forward_2x (FFTReal.h:114-118).  Its row recursion offsets the lower half of in2 with in_stride1
// (FFTReal.h:178), so the reference is only meaningful for in_stride1 == in_stride2; the driver passes one stride.
template <class T>
int real_fft2d_2x(T *out, int out_stride, const T *in1, const T *in2, int in_stride, int width, int height) {
  if (!valid_pow2(width) || !valid_pow2(height)) return 1;
  genfft::RealFFT2D<T> fft(width, height);
  fft.forward_2x((cpx<T> *)out, out_stride, in1, in_stride, in2, in_stride);
  return 0;
}

double now_s() {
  using clk = std::chrono::steady_clock;
  return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

// Restatement of the timing loops of test/fft_bench.cpp (google-benchmark is
// not installed): fresh U(-1,1) input from a std::mt19937_64 per iteration,
// std::chrono around the transform calls only (fft_bench.cpp:61-74, :98-110).
// `threads` workers each own a plan-sharing FFT object and private buffers and
// loop over their share of `count` independent transforms; returns the wall
// time of the slowest worker's timed section summed over its transforms.
template <class T, class Body>
double timed_parallel(long count, int threads, Body body) {
  threads = std::max(1, threads);
  std::vector<double> elapsed(threads, 0.0);
  std::vector<std::thread> pool;
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t] {
      long lo = count * t / threads, hi = count * (t + 1) / threads;
      ready++;
      while (!go.load()) std::this_thread::yield();
      elapsed[t] = body(t, lo, hi);
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  go = true;
  for (auto &th : pool) th.join();
  return *std::max_element(elapsed.begin(), elapsed.end());
}

}  // namespace

// `count` independent length-n transforms of a host array (rows `n` apart), each thread looping
// FFT<T>::transform (fft.h:80-85) over its share with one plan held alive -- the comparand of the full-size
// 2D parity tests (the row pass of FFT2D::scramble_row_fft, fft.h:229-241, is exactly this loop).
template <class T>
static int c2c_rows(T *out, const T *in, int n, long count, int threads, int inverse) {
  if (!valid_pow2(n) || out == in || count < 0) return 1;
  genfft::FFT<T> keep_plan(n);
  timed_parallel<T>(count, threads, [&](int, long lo, long hi) {
    genfft::FFT<T> fft(n);
    for (long k = lo; k < hi; k++) {
      if (inverse)
        fft.template transform<true>((cpx<T> *)out + (size_t)k * n, (const cpx<T> *)in + (size_t)k * n);
      else
        fft.template transform<false>((cpx<T> *)out + (size_t)k * n, (const cpx<T> *)in + (size_t)k * n);
    }
    return 0.0;
  });
  return 0;
}
extern "C" {

const char *genfft_ref_describe() {
  static char buf[256];
  auto f = genfft::GetCPUFeatures();
  snprintf(buf, sizeof buf,
           "genFFT reference, dispatch backend (SSE=%d SSE2=%d SSE3=%d SSE41=%d AVX=%d FMA=%d AVX2=%d)",
           (int)f.SSE, (int)f.SSE2, (int)f.SSE3, (int)f.SSE41, (int)f.AVX, (int)f.FMA, (int)f.AVX2);
  return buf;
}

// ---- dispatch (best ISA) back-end -------------------------------------------------
int genfft_ref_c2c_f32(float *out, const float *in, int n, int inv) { return c2c<float, false>(out, in, n, inv); }
int genfft_ref_c2c_f64(double *out, const double *in, int n, int inv) { return c2c<double, false>(out, in, n, inv); }
int genfft_ref_c2c_no_scramble_f32(float *io, int n, int inv) { return c2c_no_scramble<float, false>(io, n, inv); }
int genfft_ref_c2c_no_scramble_f64(double *io, int n, int inv) { return c2c_no_scramble<double, false>(io, n, inv); }
int genfft_ref_r2c_f32(float *out, const float *in, int n, int half) { return r2c<float, false>(out, in, n, half); }
int genfft_ref_r2c_f64(double *out, const double *in, int n, int half) { return r2c<double, false>(out, in, n, half); }
int genfft_ref_dit_f32(float *out, const float *in, int n, int half) { return dit<float, false>(out, in, n, half); }
int genfft_ref_dit_f64(double *out, const double *in, int n, int half) { return dit<double, false>(out, in, n, half); }
int genfft_ref_vert_f32(float *out, long os, const float *in, long is, int n, int cols, int inv) {
  return vert<float, false>(out, os, in, is, n, cols, inv);
}
int genfft_ref_vert_f64(double *out, long os, const double *in, long is, int n, int cols, int inv) {
  return vert<double, false>(out, os, in, is, n, cols, inv);
}
int genfft_ref_fft2d_f32(float *out, long os, const float *in, long is, int w, int h, int inv) {
  return fft2d<float>(out, os, in, is, w, h, inv);
}
int genfft_ref_fft2d_f64(double *out, long os, const double *in, long is, int w, int h, int inv) {
  return fft2d<double>(out, os, in, is, w, h, inv);
}
int genfft_ref_real_fft2d_f32(float *out, long os, const float *in, long is, int w, int h) {
  return real_fft2d<float>(out, (int)os, in, (int)is, w, h);
}
int genfft_ref_real_fft2d_f64(double *out, long os, const double *in, long is, int w, int h) {
  return real_fft2d<double>(out, (int)os, in, (int)is, w, h);
}
int genfft_ref_real_fft2d_2x_f32(float *out, long os, const float *in1, const float *in2, long is, int w, int h) {
  return real_fft2d_2x<float>(out, (int)os, in1, in2, (int)is, w, h);
}
int genfft_ref_real_fft2d_2x_f64(double *out, long os, const double *in1, const double *in2, long is, int w, int h) {
  return real_fft2d_2x<double>(out, (int)os, in1, in2, (int)is, w, h);
}
// FFT::transform_real / transform_interleave + separate_2x_real_FFT (fft.h:90-105, FFTReal.h:35-66)
int genfft_ref_transform_real_f32(float *out, const float *in, int n) {
  if (!valid_pow2(n)) return 1;
  genfft::FFT<float> fft(n);
  fft.transform_real((cpx<float> *)out, in);
  return 0;
}
int genfft_ref_transform_real_f64(double *out, const double *in, int n) {
  if (!valid_pow2(n)) return 1;
  genfft::FFT<double> fft(n);
  fft.transform_real((cpx<double> *)out, in);
  return 0;
}
int genfft_ref_two_real_f32(float *out1, float *out2, const float *in1, const float *in2, int n) {
  if (!valid_pow2(n)) return 1;
  genfft::FFT<float> fft(n);
  std::vector<cpx<float>> tmp(n);
  fft.transform_interleave(tmp.data(), in1, in2);
  genfft::separate_2x_real_FFT((cpx<float> *)out1, (cpx<float> *)out2, tmp.data(), n);
  return 0;
}
int genfft_ref_two_real_f64(double *out1, double *out2, const double *in1, const double *in2, int n) {
  if (!valid_pow2(n)) return 1;
  genfft::FFT<double> fft(n);
  std::vector<cpx<double>> tmp(n);
  fft.transform_interleave(tmp.data(), in1, in2);
  genfft::separate_2x_real_FFT((cpx<double> *)out1, (cpx<double> *)out2, tmp.data(), n);
  return 0;
}

// ---- generic scalar back-end (the arithmetic spec; no SIMD, no FMA) ------------------
int genfft_ref_generic_c2c_f32(float *out, const float *in, int n, int inv) { return c2c<float, true>(out, in, n, inv); }
int genfft_ref_generic_c2c_f64(double *out, const double *in, int n, int inv) { return c2c<double, true>(out, in, n, inv); }
int genfft_ref_generic_r2c_f32(float *out, const float *in, int n, int half) { return r2c<float, true>(out, in, n, half); }
int genfft_ref_generic_r2c_f64(double *out, const double *in, int n, int half) { return r2c<double, true>(out, in, n, half); }
int genfft_ref_generic_dit_f32(float *out, const float *in, int n, int half) { return dit<float, true>(out, in, n, half); }
int genfft_ref_generic_dit_f64(double *out, const double *in, int n, int half) { return dit<double, true>(out, in, n, half); }
int genfft_ref_generic_vert_f32(float *out, long os, const float *in, long is, int n, int cols, int inv) {
  return vert<float, true>(out, os, in, is, n, cols, inv);
}
int genfft_ref_generic_vert_f64(double *out, long os, const double *in, long is, int n, int cols, int inv) {
  return vert<double, true>(out, os, in, is, n, cols, inv);
}

// ---- the reference's in-test comparands (test/fft_ref_impl.h) ---------------------------
int genfft_ref_testref_fft_pow2_f32(float *out, const float *in, int n, int inv) {
  reference_impl::FFT_pow2((cpx<float> *)out, (const cpx<float> *)in, n, inv != 0);
  return 0;
}
int genfft_ref_testref_fft_pow2_f64(double *out, const double *in, int n, int inv) {
  reference_impl::FFT_pow2((cpx<double> *)out, (const cpx<double> *)in, n, inv != 0);
  return 0;
}
int genfft_ref_testref_dft_f64(double *out, const double *in, int n, int inv) {
  reference_impl::DFT((cpx<double> *)out, (const cpx<double> *)in, n, inv != 0);
  return 0;
}

// ---- the reference's test data generator (test/test_util.h:36-57) --------------------
void genfft_ref_dummy_complex_f32(float *dst, long n, int real) {
  std::vector<cpx<float>> v(n);
  DummyData(v, real != 0);
  memcpy(dst, v.data(), n * sizeof(cpx<float>));
}
void genfft_ref_dummy_complex_f64(double *dst, long n, int real) {
  std::vector<cpx<double>> v(n);
  DummyData(v, real != 0);
  memcpy(dst, v.data(), n * sizeof(cpx<double>));
}
void genfft_ref_dummy_real_f32(float *dst, long n) {
  std::vector<float> v(n);
  DummyData(v);
  memcpy(dst, v.data(), n * sizeof(float));
}
void genfft_ref_dummy_real_f64(double *dst, long n) {
  std::vector<double> v(n);
  DummyData(v);
  memcpy(dst, v.data(), n * sizeof(double));
}

// ---- CPU baseline timing (restated test/fft_bench.cpp loops) ---------------------------
// Each returns the seconds spent inside the transform calls by the slowest thread
// for `count` transforms spread over `threads` workers; <0 on bad arguments.

// FFT_1D (fft_bench.cpp:41-76): forward only when fwd_only!=0, else forward+inverse pair.
double genfft_ref_bench_c2c_f32(int n, long count, int threads, int fwd_only) {
  if (!valid_pow2(n)) return -1;
  genfft::FFT<float> warm(n);  // build the plan (twiddles) outside the timed region, like fft_bench.cpp:51
  return timed_parallel<float>(count, threads, [&](int t, long lo, long hi) {
    std::mt19937_64 rng(5489u + t);
    std::uniform_real_distribution<float> dist(-1, 1);
    genfft::FFT<float> fft(n);
    std::vector<cpx<float>> in(n), out(n), iout(n);
    double acc = 0;
    for (long k = lo; k < hi; k++) {
      for (auto &c : in) c = {dist(rng), dist(rng)};
      double t0 = now_s();
      fft.transform<false>(out.data(), in.data());
      if (!fwd_only) fft.transform<true>(iout.data(), out.data());
      acc += now_s() - t0;
    }
    return acc;
  });
}

// The same transforms on a BATCH ARRAY held in host memory: `count` distinct length-n sequences in, `count` out, each
// thread looping FFT<float>::transform over its contiguous share (the reference has no batch API, fft.h:80-85, so
// this is what a caller of the CPU library does with the bench workload).  Unlike the fft_bench.cpp loop above,
// whose single in/out buffer stays in L1/L2, the data streams from and to DRAM -- the same host-buffer-to-host-
// buffer contract the CUDA library's host-pointer entry point is timed on.  Returns the wall time of the slowest
// thread for one sweep over the array: mean over `reps` sweeps after `warm` untimed ones.
double genfft_ref_bench_c2c_array_f32(int n, long count, int threads, int warm, int reps) {
  if (!valid_pow2(n) || count < 1) return -1;
  genfft::FFT<float> keep_plan(n);  // holds the per-N singleton alive across the sweeps
  std::vector<cpx<float>> in((size_t)n * count), out((size_t)n * count);
  {  // fill in parallel, untimed (one generator per thread)
    std::vector<std::thread> pool;
    const int nt = std::max(1, threads);
    for (int t = 0; t < nt; t++)
      pool.emplace_back([&, t] {
        std::mt19937_64 rng(5489u + t);
        std::uniform_real_distribution<float> dist(-1, 1);
        const size_t lo = in.size() * t / nt, hi = in.size() * (t + 1) / nt;
        for (size_t k = lo; k < hi; k++) in[k] = {dist(rng), dist(rng)};
        for (size_t k = lo; k < hi; k++) out[k] = {0.f, 0.f};  // touch the pages
      });
    for (auto &th : pool) th.join();
  }
  reps = std::max(1, reps);
  double sum = 0;
  for (int r = -std::max(0, warm); r < reps; r++) {
    double t = timed_parallel<float>(count, threads, [&](int, long lo, long hi) {
      genfft::FFT<float> fft(n);
      double t0 = now_s();
      for (long k = lo; k < hi; k++) fft.transform<false>(out.data() + (size_t)k * n, in.data() + (size_t)k * n);
      return now_s() - t0;
    });
    if (r >= 0) sum += t;
  }
  return sum / reps;
}

double genfft_ref_bench_c2c_f64(int n, long count, int threads, int fwd_only) {
  if (!valid_pow2(n)) return -1;
  genfft::FFT<double> warm(n);
  return timed_parallel<double>(count, threads, [&](int t, long lo, long hi) {
    std::mt19937_64 rng(5489u + t);
    std::uniform_real_distribution<double> dist(-1, 1);
    genfft::FFT<double> fft(n);
    std::vector<cpx<double>> in(n), out(n), iout(n);
    double acc = 0;
    for (long k = lo; k < hi; k++) {
      for (auto &c : in) c = {dist(rng), dist(rng)};
      double t0 = now_s();
      fft.transform<false>(out.data(), in.data());
      if (!fwd_only) fft.transform<true>(iout.data(), out.data());
      acc += now_s() - t0;
    }
    return acc;
  });
}

// RealFFT_1D (fft_bench.cpp:80-111): RealFFT<float>::forward(out,in,half=true)
double genfft_ref_bench_r2c_f32(int n, long count, int threads) {
  if (!valid_pow2(n)) return -1;
  genfft::RealFFT<float> warm(n);
  return timed_parallel<float>(count, threads, [&](int t, long lo, long hi) {
    std::mt19937_64 rng(5489u + t);
    std::uniform_real_distribution<float> dist(-1, 1);
    genfft::RealFFT<float> fft(n);
    std::vector<float> in(n);
    std::vector<cpx<float>> out(n);
    double acc = 0;
    for (long k = lo; k < hi; k++) {
      for (auto &r : in) r = dist(rng);
      double t0 = now_s();
      fft.forward(out.data(), in.data(), true);
      acc += now_s() - t0;
    }
    return acc;
  });
}

// 2D (no 2D benchmark exists in the reference; same methodology applied to FFT2D, fft.h:213-218).
// The reference is single-threaded: one transform on one core.
double genfft_ref_bench_fft2d_f32(int w, int h, long count) {
  if (!valid_pow2(w) || !valid_pow2(h)) return -1;
  std::mt19937_64 rng;
  std::uniform_real_distribution<float> dist(-1, 1);
  genfft::FFT2D<float> fft(w, h);
  std::vector<cpx<float>> in((size_t)w * h), out((size_t)w * h);
  double acc = 0;
  for (long k = 0; k < count; k++) {
    for (auto &c : in) c = {dist(rng), dist(rng)};
    double t0 = now_s();
    fft.transform<false>(out.data(), w, in.data(), w);
    acc += now_s() - t0;
  }
  return acc;
}

int genfft_ref_c2c_rows_f32(float *out, const float *in, int n, long count, int threads, int inv) {
  return c2c_rows<float>(out, in, n, count, threads, inv);
}
int genfft_ref_c2c_rows_f64(double *out, const double *in, int n, long count, int threads, int inv) {
  return c2c_rows<double>(out, in, n, count, threads, inv);
}

int genfft_ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
