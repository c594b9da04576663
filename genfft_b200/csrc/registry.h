// Registry of compiled tile-kernel instantiations (host side).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "launch.h"
#include "tile_kernel.cuh"
#include "chain_kernel.cuh"

namespace genfft_cuda {

constexpr int kNumModes = 11;

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
#ifndef GENFFT_EMU
  if (prm.pdl) {  // programmatic stream serialization: see fft_tile_kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)K::THREADS);
    cfg.dynamicSmemBytes = K::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, fft_tile_kernel<T, L, P, C, MODE, INV>, prm);
    return;
  }
#endif
  GENFFT_LAUNCH((fft_tile_kernel<T, L, P, C, MODE, INV>), grid, K::THREADS, K::SMEM_BYTES, stream, prm);
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
    if constexpr (P == 16 && L >= 64) {  // last pass of a distributed transform: the store is the all-to-all
      add_mode<T, L, P, C, M_PEER2>(e);
      add_mode<T, L, P, C, M_PEER4>(e);
      add_mode<T, L, P, C, M_PEER8>(e);
    }
  }
  return e;
}

// ---- L2-resident pass chains (chain_kernel.cuh): pairs of wide shapes with equal CTA sizes ----
struct ChainEntry {
  int precision;  // 0: float, 1: double (GENFFT_CUDA_F32 / F64)
  int p;          // points per thread of both passes
  int la, ca, ma, lb, cb, mb, inv;
  int threads;
  size_t smem;
  const void* func;
  void (*launch)(const ChainParams& cp, unsigned grid, cudaStream_t stream);
};

template <typename T, class KA, class KB>
void launch_chain_t(const ChainParams& cp, unsigned grid, cudaStream_t stream) {
  constexpr size_t smem = KA::SMEM_BYTES > KB::SMEM_BYTES ? KA::SMEM_BYTES : KB::SMEM_BYTES;
  GENFFT_LAUNCH((fft_chain_kernel<T, KA, KB>), grid, KA::THREADS, smem, stream, cp);
}

// INV applies to the compile-time-direction modes; M_GEN takes the direction from PassParams::inverse
template <typename T, int LA, int CA, int MA, int LB, int CB, int MB, bool INV, int PP = 16>
void add_chain(std::vector<ChainEntry>& v) {
  using KA = TileKernel<T, LA, PP, CA, MA, MA == M_GEN ? false : INV>;
  using KB = TileKernel<T, LB, PP, CB, MB, MB == M_GEN ? false : INV>;
  ChainEntry e = {};
  e.precision = sizeof(T) == 4 ? 0 : 1;
  e.p = PP;
  e.la = LA; e.ca = CA; e.ma = MA;
  e.lb = LB; e.cb = CB; e.mb = MB;
  e.inv = INV ? 1 : 0;
  e.threads = KA::THREADS;
  e.smem = KA::SMEM_BYTES > KB::SMEM_BYTES ? KA::SMEM_BYTES : KB::SMEM_BYTES;
  e.func = reinterpret_cast<const void*>(&fft_chain_kernel<T, KA, KB>);
  e.launch = &launch_chain_t<T, KA, KB>;
  v.push_back(e);
}

template <typename T, int LA, int CA, int LB, int CB, int MB>
void add_peer_chains(std::vector<ChainEntry>& v) {
  add_chain<T, LA, CA, M_FIRST, LB, CB, MB, false>(v);
  add_chain<T, LA, CA, M_FIRST, LB, CB, MB, true>(v);
  add_chain<T, LA, CA, M_COL, LB, CB, MB, false>(v);
  add_chain<T, LA, CA, M_COL, LB, CB, MB, true>(v);
}

// every mode pair the plans chain, for one pair of shapes
template <typename T, int LA, int CA, int LB, int CB>
void add_chain_shapes(std::vector<ChainEntry>& v) {
  add_chain<T, LA, CA, M_FIRST, LB, CB, M_COLTW, false>(v);  // 2-pass 1D, rows of 2D
  add_chain<T, LA, CA, M_FIRST, LB, CB, M_COLTW, true>(v);
  add_chain<T, LA, CA, M_COL, LB, CB, M_COLTW, false>(v);    // 2-pass columns
  add_chain<T, LA, CA, M_COL, LB, CB, M_COLTW, true>(v);
  add_chain<T, LA, CA, M_COLTW, LB, CB, M_COLTW, false>(v);  // passes 2+3 of a 3-pass transform
  add_chain<T, LA, CA, M_COLTW, LB, CB, M_COLTW, true>(v);
  add_chain<T, LA, CA, M_COLTW, LB, CB, M_COLTWDIT, false>(v);  // ... of a real transform (split fused)
  add_chain<T, LA, CA, M_FIRST, LB, CB, M_COLTWDIT, false>(v);
  add_chain<T, LA, CA, M_FIRST, LB, CB, M_GEN, false>(v);    // distributed 2D: B stores to the peers
  add_chain<T, LA, CA, M_FIRST, LB, CB, M_GEN, true>(v);
  add_chain<T, LA, CA, M_COL, LB, CB, M_GEN, false>(v);
  add_chain<T, LA, CA, M_COL, LB, CB, M_GEN, true>(v);
  add_peer_chains<T, LA, CA, LB, CB, M_PEER2>(v);  // ... through the compile-time peer modes (p2p transport)
  add_peer_chains<T, LA, CA, LB, CB, M_PEER4>(v);
  add_peer_chains<T, LA, CA, LB, CB, M_PEER8>(v);
}

// defined in chains_inst.cu (compiled once per GENFFT_CSET)
void register_chains_0(std::vector<ChainEntry>& v);
void register_chains_1(std::vector<ChainEntry>& v);
void register_chains_2(std::vector<ChainEntry>& v);
void register_chains_3(std::vector<ChainEntry>& v);

// defined in kernels_inst.cu (compiled once per GENFFT_KSET)
void register_kernels_f32_small(std::vector<KernelEntry>& v);
void register_kernels_f32_mid(std::vector<KernelEntry>& v);
void register_kernels_f32_large(std::vector<KernelEntry>& v);
void register_kernels_f64_small(std::vector<KernelEntry>& v);
void register_kernels_f64_mid(std::vector<KernelEntry>& v);
void register_kernels_f64_large(std::vector<KernelEntry>& v);

}  // namespace genfft_cuda
