// genfft_cuda/backend.h -- plugs the CUDA implementation into the REFERENCE's own headers.
//
// The reference selects its back-end at its "L2" boundary: factory functions with the signature
//   std::shared_ptr<impl::FFTBase<T>>(int n, T)          (FFTImplFactory<T>,     include/genFFT/fft.h:41-52)
//   std::shared_ptr<impl::FFTVertBase<T>>(int n, T)      (FFTVertImplFactory<T>)
//   std::shared_ptr<impl::FFTDITBase<T>>(int n, T)       (FFTDITImplFactory<T>)
// which are (a) template parameters of the public classes (fft.h:56,117,173; FFTReal.h:188) and (b) what
// the namespace alias genfft::backend::{GetImpl,GetVertImpl,GetDITImpl} resolves to
// (include/genFFT/FFTBackend.h:40-70; dispatch mode: src/fft_x86_dispatch.cpp:64-139).
// This header provides those three factories in namespace genfft::impl_cuda, so that
//
//     #include <genFFT/fft.h>              // the reference, unmodified
//     #include <genfft_cuda/backend.h>
//     genfft::FFT<float, genfft::impl_cuda::GetImpl> fft(4096);     // reference class, CUDA kernels
//     genfft::RealFFT<float, genfft::impl_cuda::GetImpl, genfft::impl_cuda::GetDITImpl> rfft(1 << 20);
//
// works today, and the three-line FFTBackend.h arm shown in INTEGRATION.md makes it the default.
//
// Contract of the returned objects (what the reference's classes rely on):
//   FFTBase<T>::forward/inverse(T* data)                   in place, HOST pointer, input BIT-REVERSED
//                                                          (the class ran scramble() first, fft.h:83-84)
//   FFTVertBase<T>::forward/inverse(T* data, int stride, int columns)   stride in SCALARS (fft.h:135,149)
//   FFTDITBase<T>::apply(T* out, const T* in, bool half) const noexcept (FFTDIT.h:43-49)
// Include AFTER the reference's fft.h (it needs impl::FFTBase etc.).
#ifndef GENFFT_CUDA_BACKEND_H
#define GENFFT_CUDA_BACKEND_H

#ifndef GEN_FFT_LEVEL_H
#error "include the reference's <genFFT/fft.h> (or <genFFT/FFTLevel.h> and <genFFT/FFTDIT.h>) before genfft_cuda/backend.h"
#endif

#include <cstdio>
#include <cstdlib>
#include <memory>

#include "../genfft_cuda.h"

namespace genfft {
namespace impl_cuda {

namespace detail {
template <class T> struct prec;
template <> struct prec<float> { static constexpr int value = GENFFT_CUDA_F32; };
template <> struct prec<double> { static constexpr int value = GENFFT_CUDA_F64; };

// The reference's interfaces return void and are noexcept in places; like the reference's assert, a
// failure here is fatal.
inline void must(int rc, const char* what) {
  if (rc != GENFFT_CUDA_OK) {
    std::fprintf(stderr, "genfft_cuda backend: %s failed: %s\n", what, genfft_cuda_last_error_string());
    std::abort();
  }
}
}  // namespace detail

template <class T>
struct CudaFFT : impl::FFTBase<T> {
  explicit CudaFFT(int n) { detail::must(genfft_cuda_plan_c2c_1d(&plan, detail::prec<T>::value, n, 1, 0, 0), "plan_c2c_1d"); }
  ~CudaFFT() override { genfft_cuda_plan_destroy(plan); }
  void forward(T* data) override { detail::must(genfft_cuda_exec_c2c_no_scramble(plan, data, 0), "exec_c2c_no_scramble"); }
  void inverse(T* data) override { detail::must(genfft_cuda_exec_c2c_no_scramble(plan, data, 1), "exec_c2c_no_scramble"); }
  genfft_cuda_plan_t plan = nullptr;
};

template <class T>
struct CudaFFTVert : impl::FFTVertBase<T> {
  explicit CudaFFTVert(int n) { detail::must(genfft_cuda_plan_vert(&plan, detail::prec<T>::value, n), "plan_vert"); }
  ~CudaFFTVert() override { genfft_cuda_plan_destroy(plan); }
  void forward(T* data, int stride, int columns) override {
    detail::must(genfft_cuda_exec_vert_no_scramble(plan, data, stride / 2, columns, 0), "exec_vert_no_scramble");
  }
  void inverse(T* data, int stride, int columns) override {
    detail::must(genfft_cuda_exec_vert_no_scramble(plan, data, stride / 2, columns, 1), "exec_vert_no_scramble");
  }
  genfft_cuda_plan_t plan = nullptr;
};

template <class T>
struct CudaDIT : impl::FFTDITBase<T> {
  explicit CudaDIT(int n) { detail::must(genfft_cuda_plan_dit(&plan, detail::prec<T>::value, n), "plan_dit"); }
  ~CudaDIT() override { genfft_cuda_plan_destroy(plan); }
  void apply(T* out, const T* in, bool half) const noexcept override {
    detail::must(genfft_cuda_exec_dit(plan, out, in, half), "exec_dit");
  }
  genfft_cuda_plan_t plan = nullptr;
};

// The six symbols that replace impl_x86_dispatch::Get*Impl (include/genFFT/x86/fft_x86_dispatch.h:33-40).
inline std::shared_ptr<impl::FFTBase<float>> GetImpl(int n, float) { return std::make_shared<CudaFFT<float>>(n); }
inline std::shared_ptr<impl::FFTBase<double>> GetImpl(int n, double) { return std::make_shared<CudaFFT<double>>(n); }
inline std::shared_ptr<impl::FFTVertBase<float>> GetVertImpl(int n, float) { return std::make_shared<CudaFFTVert<float>>(n); }
inline std::shared_ptr<impl::FFTVertBase<double>> GetVertImpl(int n, double) { return std::make_shared<CudaFFTVert<double>>(n); }
inline std::shared_ptr<impl::FFTDITBase<float>> GetDITImpl(int n, float) { return std::make_shared<CudaDIT<float>>(n); }
inline std::shared_ptr<impl::FFTDITBase<double>> GetDITImpl(int n, double) { return std::make_shared<CudaDIT<double>>(n); }

}  // namespace impl_cuda
}  // namespace genfft

#endif  // GENFFT_CUDA_BACKEND_H
