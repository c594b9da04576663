// Explicit instantiations of the tile kernel.  This file is compiled several times with
// -DGENFFT_KSET=<n> so that the instantiations build in parallel:
//   0/1/2 : float  small / mid / large        3/4/5 : double small / mid / large
// Two shapes exist per length where they differ:
//   "narrow" (C small, ~256 threads)   for contiguous batched transforms (lanes along the sequence)
//   "wide"   (C = 16 float / 8 double) for column passes (lanes across 128 B of adjacent columns)
#include "registry.h"

namespace genfft_cuda {

// N: narrow only, W: wide only, B: used both ways
#define ADD_N(T, L, P, C) v.push_back(make_entry<T, L, P, C, true, false>())
#define ADD_W(T, L, P, C) v.push_back(make_entry<T, L, P, C, false, true>())
#define ADD_B(T, L, P, C) v.push_back(make_entry<T, L, P, C, true, true>())

#if GENFFT_KSET == 0
void register_kernels_f32_small(std::vector<KernelEntry>& v) {
  ADD_B(float, 2, 2, 128);
  ADD_B(float, 4, 4, 128);
  ADD_B(float, 8, 8, 128);
  ADD_B(float, 16, 16, 128);
  ADD_B(float, 32, 16, 64);
  ADD_B(float, 64, 16, 32);
  ADD_B(float, 128, 16, 16);
  ADD_W(float, 128, 16, 32);
  ADD_B(float, 256, 16, 16);
  ADD_W(float, 256, 16, 32);
}
#elif GENFFT_KSET == 1
void register_kernels_f32_mid(std::vector<KernelEntry>& v) {
  ADD_N(float, 512, 16, 8);
  ADD_W(float, 512, 16, 16);
  ADD_N(float, 1024, 16, 4);
  ADD_W(float, 1024, 16, 8);
  ADD_W(float, 1024, 16, 16);
  ADD_N(float, 2048, 16, 2);
  ADD_W(float, 2048, 16, 4);
  ADD_W(float, 2048, 16, 8);
}
#elif GENFFT_KSET == 2
void register_kernels_f32_large(std::vector<KernelEntry>& v) {
  ADD_N(float, 4096, 16, 1);
  ADD_W(float, 4096, 16, 4);
  ADD_N(float, 8192, 16, 1);
  ADD_W(float, 8192, 16, 2);
  ADD_B(float, 16384, 16, 1);
}
#elif GENFFT_KSET == 3
void register_kernels_f64_small(std::vector<KernelEntry>& v) {
  ADD_B(double, 2, 2, 128);
  ADD_B(double, 4, 4, 128);
  ADD_B(double, 8, 8, 128);
  ADD_B(double, 16, 16, 128);
  ADD_B(double, 32, 16, 64);
  ADD_B(double, 64, 16, 32);
  ADD_B(double, 128, 16, 16);
  ADD_B(double, 256, 16, 8);
  ADD_W(double, 256, 16, 16);
}
#elif GENFFT_KSET == 4
void register_kernels_f64_mid(std::vector<KernelEntry>& v) {
  ADD_B(double, 512, 16, 8);
  ADD_N(double, 1024, 16, 4);
  ADD_W(double, 1024, 16, 8);
  ADD_N(double, 2048, 16, 2);
  ADD_W(double, 2048, 16, 4);
}
#elif GENFFT_KSET == 5
void register_kernels_f64_large(std::vector<KernelEntry>& v) {
  ADD_N(double, 4096, 16, 1);
  ADD_W(double, 4096, 16, 2);
  ADD_B(double, 8192, 16, 1);
}
#else
#error "GENFFT_KSET must be 0..5"
#endif

}  // namespace genfft_cuda
