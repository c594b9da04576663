// Host-side plan structures of libgenfft_cuda (internal).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "registry.h"

namespace genfft_cuda {

enum PlanKind { PLAN_C2C_1D, PLAN_R2C_1D, PLAN_C2C_2D, PLAN_VERT, PLAN_DIT, PLAN_DIST_ROWS, PLAN_DIST_COLS, PLAN_R2C_2D, PLAN_C2R_1D };

// one Stockham pass of a length-N sequence: radix R = kernel length, Ns = product of earlier radices
struct PassSpec {
  const KernelEntry* k = nullptr;
  long long R = 1, Ns = 1;
  const void* tw_L = nullptr;
  const void* tw_hi = nullptr;  // inter-pass twiddle W_{Ns*R} (null for the first pass)
  const void* tw_lo = nullptr;
  int tw_shift = 0;
  const void* tw_b = nullptr;   // W_{P*Ns}^(p*i) laid out [i][p]
};

// decomposition of one length-N transform
struct Seq {
  long long N = 1;
  std::vector<PassSpec> passes;  // empty when N == 1
  bool wide = false;             // built for column (strided) use
};

// A pass launch with every decision taken (mode, grid, tile divisors); only the buffers remain to be filled in.
struct ResolvedLaunch {
  PassParams q;
  int grid = 0;
  int mode = 0;
  void (*launch)(const PassParams& prm, int grid, cudaStream_t stream) = nullptr;
};

// Single-pass contiguous transforms (c2c_1d with n on chip: C1, C2) resolved at plan creation, so that an execution is
// two pointer stores and the launch.  [inverse][input 16-byte aligned (TMA prefetch eligible)]
struct FastPath {
  bool valid = false;
  uint64_t knob_hash = 0;  // of the GENFFT_CUDA_* environment the decisions were taken under
  ResolvedLaunch rl[2][2];
};

struct Plan {
  PlanKind kind;
  int precision;
  int device;
  int num_sms;
  long long n = 0, batch = 1, in_dist = 0, out_dist = 0;
  int half = 0;
  long long width = 0, height = 0;
  int nparts = 1;
  float grid_frac[2] = {1.f, 1.f};  // share of the resident-CTA capacity: {all passes but the last, last pass}
  Seq seq;    // 1D sequence (c2c_1d: n; r2c: n/2; 2d: rows (width); vert: n)
  Seq seq_v;  // 2D: columns (height)
  Seq seq_h;  // r2c_2d: full-width complex rows (RealFFT2D::forward_2x)
  // DIT twiddles (W_n two-level)
  const void* dit_hi = nullptr;
  const void* dit_lo = nullptr;
  int dit_shift = 0;
  const void* dit_a = nullptr;     // W_n^p, p < Ns(last pass)      (fused split, multi-pass plans)
  const void* dit_b = nullptr;     // W_{2L}^k, k < L(last pass)
  const void* dit_full = nullptr;  // W_n^k, k < n/2 (fused split, single-pass plans)
  // scratch (device), lazily grown
  std::mutex mu;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  size_t scratch_need = 0;  // upper bound known at plan time (0 if none / depends on exec args)
  // ticket + per-group counters of the L2-resident pass chains (chain_kernel.cuh), one block PER STREAM: launches on
  // one stream are ordered, launches of the same plan on different streams may overlap and must not share counters
  struct ChainCtr {
    void* ptr = nullptr;
    size_t count = 0;
  };
  std::map<cudaStream_t, ChainCtr> chain_ctrs;
  void* aux = nullptr;  // device staging owned by device-pointer entry points (c2r pre-processed spectrum)
  size_t aux_bytes = 0;
  // host-pointer staging
  void* stage_in = nullptr;
  void* stage_out = nullptr;
  size_t stage_in_bytes = 0, stage_out_bytes = 0;
  std::mutex host_mu;  // host-pointer entry points hold it for the whole call (staging buffers, streams)
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t events[8] = {};
  bool streams_ready = false;
  FastPath fast;
};

}  // namespace genfft_cuda

// the opaque C handle is the plan itself
struct genfft_cuda_plan_s : genfft_cuda::Plan {};

namespace genfft_cuda {

size_t elem_size(int precision);
int set_error(int code, const char* msg);
int exec_c2c_internal(Plan* p, void* out, const void* in, int inverse, cudaStream_t stream, bool brev, bool real_in,
                      long long batch, const void* in2 = nullptr);
int exec_r2c_internal(Plan* p, void* out, const void* in, cudaStream_t stream, long long batch);
int exec_r2c_strided(Plan* p, void* out, const void* in, cudaStream_t stream, long long batch, long long in_dist,
                     long long out_dist);
bool plan_needs_scratch(const Plan* p);

}  // namespace genfft_cuda
