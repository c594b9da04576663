// Registry of compiled tile-kernel instantiations (host side).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "tile_kernel.cuh"

namespace genfft_cuda {

struct KernelEntry {
  int L, P, C;
  int threads;
  size_t smem;
  const void* func;  // __global__ function pointer (for attributes / occupancy)
  void (*launch)(const PassParams& prm, int grid, cudaStream_t stream);
  int max_ctas_per_sm;  // filled lazily per device
};

template <typename T, int L, int P, int C>
void launch_tile(const PassParams& prm, int grid, cudaStream_t stream) {
  using K = TileKernel<T, L, P, C>;
  fft_tile_kernel<T, L, P, C><<<grid, K::THREADS, K::SMEM_BYTES, stream>>>(prm);
}

template <typename T, int L, int P, int C>
KernelEntry make_entry() {
  using K = TileKernel<T, L, P, C>;
  KernelEntry e;
  e.L = L;
  e.P = P;
  e.C = C;
  e.threads = K::THREADS;
  e.smem = K::SMEM_BYTES;
  e.func = reinterpret_cast<const void*>(&fft_tile_kernel<T, L, P, C>);
  e.launch = &launch_tile<T, L, P, C>;
  e.max_ctas_per_sm = 0;
  return e;
}

// defined in kernels_*.cu
void register_kernels_f32_small(std::vector<KernelEntry>& v);
void register_kernels_f32_mid(std::vector<KernelEntry>& v);
void register_kernels_f32_large(std::vector<KernelEntry>& v);
void register_kernels_f64_small(std::vector<KernelEntry>& v);
void register_kernels_f64_mid(std::vector<KernelEntry>& v);
void register_kernels_f64_large(std::vector<KernelEntry>& v);

}  // namespace genfft_cuda
