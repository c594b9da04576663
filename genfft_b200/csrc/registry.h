// Registry of compiled tile-kernel instantiations (host side).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "tile_kernel.cuh"

namespace genfft_cuda {

constexpr int kNumModes = 8;

struct KernelEntry {
  int L, P, C;
  int threads;
  size_t smem;
  size_t smem_mode[kNumModes];  // dynamic shared memory of each compiled mode (TMA mode adds an input buffer)
  // [mode][inverse]; null when that variant is not compiled for this shape (M_GEN always is, with
  // the direction taken at run time, stored in both slots)
  const void* func[kNumModes][2];
  void (*launch[kNumModes][2])(const PassParams& prm, int grid, cudaStream_t stream);
};

template <typename T, int L, int P, int C, int MODE, bool INV>
void launch_tile(const PassParams& prm, int grid, cudaStream_t stream) {
  using K = TileKernel<T, L, P, C, MODE, INV>;
  fft_tile_kernel<T, L, P, C, MODE, INV><<<grid, K::THREADS, K::SMEM_BYTES, stream>>>(prm);
}

template <typename T, int L, int P, int C, int MODE>
void add_mode(KernelEntry& e) {
  e.smem_mode[MODE] = TileKernel<T, L, P, C, MODE, false>::SMEM_BYTES;
  if (MODE == M_ROWDIT || MODE == M_COLTWDIT) {  // forward only
    e.func[MODE][0] = reinterpret_cast<const void*>(&fft_tile_kernel<T, L, P, C, MODE, false>);
    e.launch[MODE][0] = &launch_tile<T, L, P, C, MODE, false>;
    return;
  }
  e.func[MODE][0] = reinterpret_cast<const void*>(&fft_tile_kernel<T, L, P, C, MODE, false>);
  e.launch[MODE][0] = &launch_tile<T, L, P, C, MODE, false>;
  if (MODE == M_GEN) {
    e.func[MODE][1] = e.func[MODE][0];
    e.launch[MODE][1] = e.launch[MODE][0];
  } else {
    e.func[MODE][1] = reinterpret_cast<const void*>(&fft_tile_kernel<T, L, P, C, MODE, MODE != M_GEN>);
    e.launch[MODE][1] = &launch_tile<T, L, P, C, MODE, MODE != M_GEN>;
  }
}

// narrow: contiguous batched use; wide: column / multi-pass use
template <typename T, int L, int P, int C, bool NARROW, bool WIDE>
KernelEntry make_entry() {
  using K = TileKernel<T, L, P, C, M_GEN, false>;
  KernelEntry e = {};
  e.L = L;
  e.P = P;
  e.C = C;
  e.threads = K::THREADS;
  e.smem = K::SMEM_BYTES;
  add_mode<T, L, P, C, M_GEN>(e);
  if constexpr (NARROW) {
    add_mode<T, L, P, C, M_ROW>(e);
    add_mode<T, L, P, C, M_ROWDIT>(e);
    if constexpr (L >= 256 && TileKernel<T, L, P, C, M_ROWTMA, false>::SMEM_BYTES <= 200 * 1024)
      add_mode<T, L, P, C, M_ROWTMA>(e);
  }
  if constexpr (WIDE) {
    add_mode<T, L, P, C, M_COL>(e);
    add_mode<T, L, P, C, M_COLTW>(e);
    add_mode<T, L, P, C, M_FIRST>(e);
    if constexpr (C >= 2 && L >= 16) add_mode<T, L, P, C, M_COLTWDIT>(e);
  }
  return e;
}

// defined in kernels_inst.cu (compiled once per GENFFT_KSET)
void register_kernels_f32_small(std::vector<KernelEntry>& v);
void register_kernels_f32_mid(std::vector<KernelEntry>& v);
void register_kernels_f32_large(std::vector<KernelEntry>& v);
void register_kernels_f64_small(std::vector<KernelEntry>& v);
void register_kernels_f64_mid(std::vector<KernelEntry>& v);
void register_kernels_f64_large(std::vector<KernelEntry>& v);

}  // namespace genfft_cuda
