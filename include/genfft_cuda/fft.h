// genfft_cuda/fft.h -- header-only C++ mirror of genFFT's public classes on top of libgenfft_cuda.
//
// Drop-in for `#include <genFFT/fft.h>` on the transform path: the same class names in namespace
// genfft, the same member signatures and argument meaning as the reference
//   genfft::FFT<T>      include/genFFT/fft.h:54-113      genfft::FFTVert<T>  fft.h:115-171
//   genfft::DIT<T>      fft.h:173-196                    genfft::FFT2D<T>    fft.h:198-245
//   genfft::RealFFT<T>  include/genFFT/FFTReal.h:186-221 separate_2x_real_FFT FFTReal.h:35-66
// but every transform runs on the GPU through the C ABI of include/genfft_cuda.h.  Pointers passed to
// the reference-named members are HOST pointers (what a CPU caller of genFFT has); the *_dev members and
// the batched constructors are additions for callers that keep data on the device.
//
// Do not include this header together with the reference's <genFFT/fft.h> (same names).  To plug the
// CUDA implementation into the reference's own headers instead, use genfft_cuda/backend.h.
//
// Error behaviour: the reference asserts (and is undefined under NDEBUG,
// include/genFFT/x86/fft_float_impl_x86.inl:494-495); here every failure throws genfft::cuda_error.
#ifndef GENFFT_CUDA_FFT_H
#define GENFFT_CUDA_FFT_H

#include <complex>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "../genfft_cuda.h"

namespace genfft {

using stride_t = std::ptrdiff_t;  // include/genFFT/FFTTypes.h:35
using index_t = int;              // include/genFFT/FFTTypes.h:36

struct cuda_error : std::runtime_error {
  int code;
  cuda_error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace detail {

template <class T> struct precision_of;
template <> struct precision_of<float> { static constexpr int value = GENFFT_CUDA_F32; };
template <> struct precision_of<double> { static constexpr int value = GENFFT_CUDA_F64; };

inline void check(int rc) {
  if (rc != GENFFT_CUDA_OK) {
    const char* msg = genfft_cuda_last_error_string();
    throw cuda_error(rc, std::string("libgenfft_cuda: ") + (msg ? msg : "unknown error"));
  }
}

struct plan_deleter {
  void operator()(genfft_cuda_plan_s* p) const { genfft_cuda_plan_destroy(p); }
};
// plans are immutable and shared by copies of the user object, like the reference's per-N singletons
// held through std::shared_ptr (include/genFFT/FFTLevel.h:76-92)
using plan_ptr = std::shared_ptr<genfft_cuda_plan_s>;

inline plan_ptr own(genfft_cuda_plan_t p) { return plan_ptr(p, plan_deleter()); }

}  // namespace detail

///@brief A 1D FFT for densely packed data (mirror of genfft::FFT<T>, fft.h:54-113)
template <class T>
struct FFT {
  FFT() = default;
  explicit FFT(int n) : FFT(n, 1) {}
  /// addition: `batch` transforms, in_dist/out_dist complex elements apart (0 = n)
  FFT(long long n, long long batch, long long in_dist = 0, long long out_dist = 0) {
    genfft_cuda_plan_t p = nullptr;
    detail::check(genfft_cuda_plan_c2c_1d(&p, detail::precision_of<T>::value, n, batch, in_dist, out_dist));
    impl = detail::own(p);
    this->n = (int)n;
  }

  ///@brief Computes transform in-place, without data reordering (input in bit-reversed order), fft.h:69-73
  template <bool inv>
  void transform_no_scramble(std::complex<T>* inout) {
    detail::check(genfft_cuda_exec_c2c_no_scramble(impl.get(), inout, inv));
  }

  ///@brief Computes transform (fft.h:80-85); out must not be equal to in; inverse is unscaled
  template <bool inv>
  void transform(std::complex<T>* out, const std::complex<T>* in) {
    detail::check(genfft_cuda_exec_c2c(impl.get(), out, in, inv));
  }

  /// README.txt:28-31 spelling (the reference documents forward/inverse but only ships transform<inv>)
  void forward(std::complex<T>* out, const std::complex<T>* in) { transform<false>(out, in); }
  void inverse(std::complex<T>* out, const std::complex<T>* in) { transform<true>(out, in); }
  void forward(T* out, const T* in) { transform<false>((std::complex<T>*)out, (const std::complex<T>*)in); }
  void inverse(T* out, const T* in) { transform<true>((std::complex<T>*)out, (const std::complex<T>*)in); }

  ///@brief Computes forward transform of real data (fft.h:90-94)
  void transform_real(std::complex<T>* out, const T* in) {
    detail::check(genfft_cuda_exec_c2c_real_in(impl.get(), out, in));
  }

  ///@brief Computes forward transform of interleaved real data (fft.h:100-105); separate the two spectra with
  ///       separate_2x_real_FFT.
  void transform_interleave(std::complex<T>* out, const T* in1, const T* in2) {
    detail::check(genfft_cuda_exec_c2c_interleave(impl.get(), out, in1, in2));
  }

  /// additions: device pointers, asynchronous on `stream` (a cudaStream_t)
  template <bool inv>
  void transform_dev(void* d_out, const void* d_in, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_c2c_dev(impl.get(), d_out, d_in, inv, stream));
  }
  template <bool inv>
  void transform_no_scramble_dev(void* d_inout, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_c2c_no_scramble_dev(impl.get(), d_inout, inv, stream));
  }

  int size() const noexcept { return n; }
  explicit operator bool() const noexcept { return (bool)impl; }
  genfft_cuda_plan_t plan() const noexcept { return impl.get(); }

 private:
  int n = 0;
  detail::plan_ptr impl;
};

///@brief Column-wise 1D FFT for multiple columns (mirror of genfft::FFTVert<T>, fft.h:115-171)
template <class T>
struct FFTVert {
  FFTVert() = default;
  explicit FFTVert(int n) {
    genfft_cuda_plan_t p = nullptr;
    detail::check(genfft_cuda_plan_vert(&p, detail::precision_of<T>::value, n));
    impl = detail::own(p);
    this->n = n;
  }

  ///@param stride row stride, in complex numbers, of the data array (fft.h:132-136)
  template <bool inv>
  void transform_no_scramble(std::complex<T>* data, stride_t stride, index_t cols) {
    detail::check(genfft_cuda_exec_vert_no_scramble(impl.get(), data, stride, cols, inv));
  }

  ///@brief fft.h:145-150; out must not be equal to in; strides in complex numbers
  template <bool inv>
  void transform(std::complex<T>* out, stride_t out_stride, const std::complex<T>* in, stride_t in_stride, index_t cols) {
    detail::check(genfft_cuda_exec_vert(impl.get(), out, out_stride, in, in_stride, cols, inv));
  }

  ///@brief dense overload, fft.h:158-163
  template <bool inv>
  void transform(std::complex<T>* out, const std::complex<T>* in, index_t cols) {
    transform<inv>(out, cols, in, cols, cols);
  }

  template <bool inv>
  void transform_dev(void* d_out, stride_t out_stride, const void* d_in, stride_t in_stride, index_t cols,
                     void* stream = nullptr) {
    detail::check(genfft_cuda_exec_vert_dev(impl.get(), d_out, out_stride, d_in, in_stride, cols, inv, stream));
  }

  int size() const noexcept { return n; }
  explicit operator bool() const noexcept { return (bool)impl; }

 private:
  int n = 0;
  detail::plan_ptr impl;
};

///@brief Real-FFT split / post-process (mirror of genfft::DIT<T>, fft.h:173-196)
template <class T>
class DIT {
 public:
  DIT() = default;
  explicit DIT(int n) : n(n) {
    genfft_cuda_plan_t p = nullptr;
    detail::check(genfft_cuda_plan_dit(&p, detail::precision_of<T>::value, n));
    impl = detail::own(p);
  }

  void apply(T* out, const T* in, bool half) { detail::check(genfft_cuda_exec_dit(impl.get(), out, in, half)); }
  void apply(std::complex<T>* out, const std::complex<T>* in, bool half) { apply((T*)out, (const T*)in, half); }
  void apply_dev(void* d_out, const void* d_in, bool half, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_dit_dev(impl.get(), d_out, d_in, half, stream));
  }

  int size() const noexcept { return n; }
  explicit operator bool() const noexcept { return (bool)impl; }

 private:
  int n = 0;
  detail::plan_ptr impl;
};

///@brief 2D FFT (mirror of genfft::FFT2D<T>, fft.h:198-245); note the (width, height) order
template <class T>
class FFT2D {
 public:
  FFT2D() = default;
  FFT2D(int width, int height) : w(width), h(height) {
    genfft_cuda_plan_t p = nullptr;
    detail::check(genfft_cuda_plan_c2c_2d(&p, detail::precision_of<T>::value, width, height));
    impl = detail::own(p);
  }

  ///@brief fft.h:213-218; out must not be equal to in; strides in complex elements
  template <bool inv>
  void transform(std::complex<T>* out, stride_t out_stride, const std::complex<T>* in, stride_t in_stride) {
    detail::check(genfft_cuda_exec_c2c_2d(impl.get(), out, out_stride, in, in_stride, inv));
  }
  template <bool inv>
  void transform_dev(void* d_out, stride_t out_stride, const void* d_in, stride_t in_stride, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_c2c_2d_dev(impl.get(), d_out, out_stride, d_in, in_stride, inv, stream));
  }

  int cols() const noexcept { return impl ? w : 0; }
  int rows() const noexcept { return impl ? h : 0; }
  explicit operator bool() const noexcept { return (bool)impl; }

 private:
  int w = 0, h = 0;
  detail::plan_ptr impl;
};

///@brief Recovers two transforms of real data from one interleaved transform (FFTReal.h:35-66).
/// Host pointers; runs on the device through the C ABI like everything else (no CPU arithmetic in this header).
/// Input and outputs may alias as in the reference.
template <class T>
void separate_2x_real_FFT(std::complex<T>* out1, std::complex<T>* out2, const std::complex<T>* in, int N) {
  detail::check(genfft_cuda_separate_2x_real(detail::precision_of<T>::value, out1, out2, in, N));
}
///@brief the same on device pointers
template <class T>
void separate_2x_real_FFT_dev(void* d_out1, void* d_out2, const void* d_in, int N, void* stream = nullptr) {
  detail::check(genfft_cuda_separate_2x_real_dev(detail::precision_of<T>::value, d_out1, d_out2, d_in, N, stream));
}

///@brief 2D FFT of a real image (mirror of genfft::RealFFT2D<T>, FFTReal.h:71-184); forward only, as in the reference
template <class T>
class RealFFT2D {
 public:
  RealFFT2D() = default;
  RealFFT2D(int width, int height) : w(width), h(height) {
    genfft_cuda_plan_t p = nullptr;
    detail::check(genfft_cuda_plan_r2c_2d(&p, detail::precision_of<T>::value, width, height));
    impl = detail::own(p);
  }

  ///@param out_stride stride, in complex elements, of the output array
  ///@param in_stride stride, in scalar elements, of the input array (FFTReal.h:83)
  void forward(std::complex<T>* out, int out_stride, const T* in, int in_stride) {
    detail::check(genfft_cuda_exec_r2c_2d(impl.get(), out, out_stride, in, in_stride));
  }
  void forward_dev(void* d_out, int out_stride, const void* d_in, int in_stride, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_r2c_2d_dev(impl.get(), d_out, out_stride, d_in, in_stride, stream));
  }

  ///@brief Forward transform of two real images at once: the spectrum of in1 + i*in2 (FFTReal.h:106-118)
  ///@param in_stride1 stride, in scalar elements, of the 1st input array
  ///@param in_stride2 stride, in scalar elements, of the 2nd input array
  void forward_2x(std::complex<T>* out, int out_stride, const T* in1, int in_stride1, const T* in2, int in_stride2) {
    detail::check(genfft_cuda_exec_r2c_2d_2x(impl.get(), out, out_stride, in1, in_stride1, in2, in_stride2));
  }
  void forward_2x_dev(void* d_out, int out_stride, const void* d_in1, int in_stride1, const void* d_in2,
                      int in_stride2, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_r2c_2d_2x_dev(impl.get(), d_out, out_stride, d_in1, in_stride1, d_in2, in_stride2,
                                                 stream));
  }

  int cols() const { return impl ? w : 0; }
  int rows() const { return impl ? h : 0; }

 private:
  int w = 0, h = 0;
  detail::plan_ptr impl;
};

///@brief 1D FFT of real input (mirror of genfft::RealFFT<T>, FFTReal.h:186-221)
template <class T>
struct RealFFT {
  RealFFT() = default;
  RealFFT(int n) : n(n) {
    // `half` is a run-time argument of forward() in the reference; one plan per value is created lazily
  }

  /// @brief Computes forward transform of real-valued signal (FFTReal.h:204-213)
  /// @param half if true, only N/2+1 values are stored; if false, the upper half of the spectrum is reconstituted
  void forward(std::complex<T>* out, const T* in, bool half) {
    detail::check(genfft_cuda_exec_r2c(plan(half), out, in));
  }
  void forward_dev(void* d_out, const void* d_in, bool half, void* stream = nullptr) {
    detail::check(genfft_cuda_exec_r2c_dev(plan(half), d_out, d_in, stream));
  }

  /// addition (the reference has no inverse real FFT, README.txt:51-52): n/2+1 bins -> n real points, unscaled
  void inverse(T* out, const std::complex<T>* in) {
    if (!inv_impl) {
      genfft_cuda_plan_t raw = nullptr;
      detail::check(genfft_cuda_plan_c2r_1d(&raw, detail::precision_of<T>::value, n, 1, 0, 0));
      inv_impl = detail::own(raw);
    }
    detail::check(genfft_cuda_exec_c2r(inv_impl.get(), out, in));
  }

  int size() const noexcept { return n; }

 private:
  detail::plan_ptr inv_impl;
  genfft_cuda_plan_t plan(bool half) {
    detail::plan_ptr& p = impl[half ? 1 : 0];
    if (!p) {
      genfft_cuda_plan_t raw = nullptr;
      detail::check(genfft_cuda_plan_r2c_1d(&raw, detail::precision_of<T>::value, n, 1, half, 0, 0));
      p = detail::own(raw);
    }
    return p.get();
  }
  detail::plan_ptr impl[2];
  int n = 0;
};

}  // namespace genfft

#endif  // GENFFT_CUDA_FFT_H
