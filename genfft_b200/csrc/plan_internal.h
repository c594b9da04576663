// libgenfft_cuda, internal: what the three host-side translation units share.
//   planner.cu     errors, run-time knobs, kernel registry, twiddle tables, sequence decomposition, pass emission, launches
//   pass_chain.cu  the multi-pass driver: buffer assignment, L2-resident chains of consecutive passes, fused stores
//   abi.cu         the extern "C" boundary (include/genfft_cuda.h): plan creation and device-pointer execution
// (host_exec.cu holds the host-pointer entry points and sees only plan.h.)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/genfft_cuda.h"
#include "aux_kernels.cuh"
#include "launch.h"
#include "plan.h"

namespace genfft_cuda {

int fail(int code, const char* fmt, ...);
#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(GENFFT_CUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                  __FILE__, __LINE__);                                                            \
  } while (0)

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;
extern std::atomic<uint64_t> g_mode_launches[16];

bool is_pow2(long long n);
int ilog2(long long n);

// run-time knobs (GENFFT_CUDA_*): a per-thread snapshot revalidated once per outermost scope
struct KnobScope {
  KnobScope();
  ~KnobScope();
};
int env_int(const char* name, int dflt);
uint64_t current_knob_hash();

int usable_device(int* dev_out, int* sms_out);
const ChainEntry* find_chain(int precision, const KernelEntry* ka, int ma, const KernelEntry* kb, int mb, int inv);
const KernelEntry* find_kernel(int precision, long long L, bool wide);
int kernel_occupancy(const KernelEntry* k, const void* func, size_t smem, int device, int* out);

int twiddle_table(int device, int precision, long long M, long long step, long long count, const void** out);
int two_level_table(int device, int precision, long long M, const void** hi, const void** lo, int* shift);
// max_single > 0 overrides the longest transform done in one pass (the distributed row plans split earlier so that the
// HBM-bound first pass and the NVLink-bound storing pass run as one chain)
int build_seq(Seq* seq, int device, int precision, long long N, bool wide, long long max_single = 0);

PassParams emit_1d(const PassSpec& ps, long long N, const void* in, long long in_dist, void* out, long long out_dist,
                   long long batch, int inverse, bool brev);
PassParams emit_col(const PassSpec& ps, long long N, const void* in, long long in_pitch, void* out, long long out_pitch,
                    long long cols, int inverse, bool brev);
bool strides_fit_32(const PassParams& p);
void set_tile_divisors(PassParams& p);
int resolve_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, ResolvedLaunch* r);
int launch_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, cudaStream_t stream);
void build_fast_path(Plan* p);
int launch_chain(Plan* plan, const ChainEntry* ce, ChainParams& cp, cudaStream_t stream);
int launch_copy(int precision, const CopyParams& cp, long long batch, cudaStream_t stream);
int launch_dit(const Plan* plan, void* out, long long out_dist, const void* in, long long in_dist, int n, int half,
               long long batch, bool real_scalar, cudaStream_t stream);
int ensure_scratch(Plan* plan, size_t bytes);

// ---- the multi-pass driver (pass_chain.cu) ----
// ------------------------------------------------------------------------------------------------
// generic multi-pass driver: a chain of steps over buffers IN -> {OUT, SCRATCH} -> OUT
// ------------------------------------------------------------------------------------------------
struct View {
  void* ptr;
  long long pitch;  // distance between sequences (1D) or row pitch (2D / columns)
};

struct Step {
  const PassSpec* ps;
  long long N;   // sequence length of the Seq this pass belongs to
  bool col;      // column pass (emit_col) or row/1D pass (emit_1d)
  bool safe;     // reads and writes the same positions per tile -> may run in place
  bool brev;
  bool real_in;
  bool c2r = false;  // half-spectrum inverse: the first pass builds the packed spectrum from the n/2+1 input bins while loading
};

// Optional override of the last pass's store: the output bin index is split at 2^part_log2 and the high
// part selects a peer buffer (fused all-to-all over NVLink) or a block at khi*part_stride.
struct FinalStore {
  int part_log2 = -1;
  void* const* peers = nullptr;
  int npeers = 0;
  long long part_stride = 0;
  long long peer_offset = 0;  // elements added to every peer pointer
  // distributed four-step 1D: multiply by W_{tw_n}^(row * (tw_col0 + column)) in the storing pass (compile-time peer
  // modes only); *fused reports whether it was done, otherwise the caller runs the twiddle pass itself
  long long tw_n = 0, tw_col0 = 0;
  bool* fused = nullptr;
};

// Optional fusion of the real-FFT split into the last pass of a multi-pass chain (M_COLTWDIT).
struct DitFuse {
  int half = 0;
  const void* dit_a = nullptr;   // W_n^p, p < Ns of the last pass
  const void* dit_tw = nullptr;  // W_{2L}^k, k < L of the last pass
};

void seq_steps(const Seq& seq, bool col, std::vector<Step>& steps, bool brev_first, bool real_first);
// runs `steps`; count = batch (1D) or rows (row passes of 2D); cols = columns for column passes
int run_chain(Plan* plan, const std::vector<Step>& steps_in, View in, View out, long long scratch_pitch,
              size_t scratch_elems, long long count, long long cols, int inverse, cudaStream_t stream,
              const FinalStore* fs = nullptr, const DitFuse* df = nullptr, const void* in2 = nullptr);

}  // namespace genfft_cuda
