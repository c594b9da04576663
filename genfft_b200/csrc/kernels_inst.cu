// Explicit instantiations of the tile kernel.  This file is compiled several times with
// -DGENFFT_KSET=<n> so that the instantiations build in parallel:
//   0/1/2 : float  small / mid / large        3/4/5 : double small / mid / large
// Two shapes exist per length where they differ:
//   "narrow" (C small, ~256 threads)   for contiguous batched transforms (lanes along the sequence)
//   "wide"   (C = 16 float / 8 double) for column passes (lanes across 128 B of adjacent columns)
#include "registry.h"

namespace genfft_cuda {

#define ADD(T, L, P, C) v.push_back(make_entry<T, L, P, C>())

#if GENFFT_KSET == 0
void register_kernels_f32_small(std::vector<KernelEntry>& v) {
  ADD(float, 2, 2, 128);
  ADD(float, 4, 4, 128);
  ADD(float, 8, 8, 128);
  ADD(float, 16, 16, 128);
  ADD(float, 32, 16, 64);
  ADD(float, 64, 16, 32);
  ADD(float, 128, 16, 16);
  ADD(float, 256, 16, 16);
}
#elif GENFFT_KSET == 1
void register_kernels_f32_mid(std::vector<KernelEntry>& v) {
  ADD(float, 512, 16, 8);
  ADD(float, 512, 16, 16);
  ADD(float, 1024, 16, 4);
  ADD(float, 1024, 16, 16);
  ADD(float, 2048, 16, 2);
  ADD(float, 2048, 16, 8);
}
#elif GENFFT_KSET == 2
void register_kernels_f32_large(std::vector<KernelEntry>& v) {
  ADD(float, 4096, 16, 1);
  ADD(float, 4096, 16, 4);
  ADD(float, 8192, 16, 1);
  ADD(float, 8192, 16, 2);
  ADD(float, 16384, 16, 1);
}
#elif GENFFT_KSET == 3
void register_kernels_f64_small(std::vector<KernelEntry>& v) {
  ADD(double, 2, 2, 128);
  ADD(double, 4, 4, 128);
  ADD(double, 8, 8, 128);
  ADD(double, 16, 16, 128);
  ADD(double, 32, 16, 64);
  ADD(double, 64, 16, 32);
  ADD(double, 128, 16, 16);
  ADD(double, 256, 16, 8);
  ADD(double, 256, 16, 16);
}
#elif GENFFT_KSET == 4
void register_kernels_f64_mid(std::vector<KernelEntry>& v) {
  ADD(double, 512, 16, 8);
  ADD(double, 1024, 16, 4);
  ADD(double, 1024, 16, 8);
  ADD(double, 2048, 16, 2);
  ADD(double, 2048, 16, 4);
}
#elif GENFFT_KSET == 5
void register_kernels_f64_large(std::vector<KernelEntry>& v) {
  ADD(double, 4096, 16, 1);
  ADD(double, 4096, 16, 2);
  ADD(double, 8192, 16, 1);
}
#else
#error "GENFFT_KSET must be 0..5"
#endif

}  // namespace genfft_cuda
