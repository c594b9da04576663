// libgenfft_cuda: plans, pass emission and execution (host side).
//
// The reference builds an "impl" per size from a factory switch (include/genFFT/x86/fft_float_impl_x86.inl:464-497)
// whose constructor chain computes one twiddle table per radix-2 level (include/genFFT/FFTTwiddle.h:44-51).
// Here a plan is a short list of Stockham passes (1 for sizes that fit on chip, 2-3 above), each a
// launch of the tile kernel with its own addressing, plus fp64-computed twiddle tables stored at the
// transform's precision.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/genfft_cuda.h"
#include "aux_kernels.cuh"
#include "launch.h"
#include "plan.h"

namespace genfft_cuda {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_mode_launches[16];  // per compiled addressing mode (tests: "the specialised mode really ran")

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(GENFFT_CUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                  __FILE__, __LINE__);                                                            \
  } while (0)

size_t elem_size(int precision) { return precision == GENFFT_CUDA_F32 ? 8 : 16; }

static bool is_pow2(long long n) { return n >= 1 && (n & (n - 1)) == 0; }
static int ilog2(long long n) {
  int l = 0;
  while ((1LL << l) < n) l++;
  return l;
}

// Run-time knobs (GENFFT_CUDA_*) are consulted on every execution so that a caller can change
// them between calls.  getenv() is a linear scan of the whole environment (~0.3 us with ~130 entries), and one exec call
// consults eight knobs -- more host time than the launch of a small transform costs (C1).  So the GENFFT_CUDA_*
// entries are collected into a per-thread snapshot that is revalidated (identity of the environment's entry
// pointers: setenv / putenv / unsetenv all replace or move them) once per KnobScope, i.e. once per execution.
extern "C" char** environ;
namespace {
struct KnobSnapshot {
  char** env = nullptr;
  size_t count = 0;
  uintptr_t sig = 0;
  std::vector<const char*> entries;  // "GENFFT_CUDA_<NAME>=<value>" strings
  uint64_t knob_hash = 0;            // of the entries' text: what a cached launch decision depends on
  int depth = 0;
  bool valid = false;
};
thread_local KnobSnapshot t_knobs;

void knobs_validate() {
  KnobSnapshot& k = t_knobs;
  char** e = environ;
  size_t n = 0;
  uintptr_t sig = 0;
  if (e)
    for (; e[n]; n++) sig += reinterpret_cast<uintptr_t>(e[n]) ^ (uintptr_t)n;
  if (k.valid && k.env == e && k.count == n && k.sig == sig) return;
  k.entries.clear();
  k.knob_hash = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++)
    if (e[i][0] == 'G' && strncmp(e[i], "GENFFT_CUDA_", 12) == 0) {
      k.entries.push_back(e[i]);
      for (const char* c = e[i]; *c; c++) k.knob_hash = (k.knob_hash ^ (unsigned char)*c) * 1099511628211ull;
      k.knob_hash = (k.knob_hash ^ 0xffu) * 1099511628211ull;
    }
  k.env = e;
  k.count = n;
  k.sig = sig;
  k.valid = true;
}

// every function that reads knobs at execution time opens a scope; only the outermost one revalidates
struct KnobScope {
  KnobScope() {
    if (t_knobs.depth++ == 0) knobs_validate();
  }
  ~KnobScope() { t_knobs.depth--; }
};
}  // namespace

static int env_int(const char* name, int dflt) {
  if (t_knobs.depth == 0) knobs_validate();  // plan-creation-time reads
  const size_t len = strlen(name);
  for (const char* ent : t_knobs.entries)
    if (strncmp(ent, name, len) == 0 && ent[len] == '=') return ent[len + 1] ? atoi(ent + len + 1) : dflt;
  return dflt;
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
static int usable_device(int* dev_out, int* sms_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(GENFFT_CUDA_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0, sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10)
    return fail(GENFFT_CUDA_ERR_CUDA, "device %d has compute capability %d.x; this library is built for sm_100a only",
                dev, major);
  *dev_out = dev;
  *sms_out = sms;
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel registry
// ------------------------------------------------------------------------------------------------
static std::vector<KernelEntry>& registry(int precision) {
  static std::vector<KernelEntry> f32, f64;
  static std::once_flag once;
  std::call_once(once, [] {
    register_kernels_f32_small(f32);
    register_kernels_f32_mid(f32);
    register_kernels_f32_large(f32);
    register_kernels_f64_small(f64);
    register_kernels_f64_mid(f64);
    register_kernels_f64_large(f64);
  });
  return precision == GENFFT_CUDA_F32 ? f32 : f64;
}

static const ChainEntry* find_chain(int precision, const KernelEntry* ka, int ma, const KernelEntry* kb, int mb, int inv) {
  static std::vector<ChainEntry> chains;
  static std::once_flag once;
  std::call_once(once, [] {
    register_chains_0(chains);
    register_chains_1(chains);
    register_chains_2(chains);
    register_chains_3(chains);
  });
  if (ka->P != 16 || kb->P != 16) return nullptr;
  const int prec = precision == GENFFT_CUDA_F32 ? 0 : 1;
  for (auto& e : chains)
    if (e.precision == prec && e.la == ka->L && e.ca == ka->C && e.ma == ma && e.lb == kb->L && e.cb == kb->C &&
        e.mb == mb && e.inv == inv)
      return &e;
  return nullptr;
}

// Kernel shape for a length-L pass.  Narrow (contiguous batched) use takes the fewest sequences per CTA.  Wide (column)
// use wants row segments of at least 128 bytes and CTAs of ~256 threads in float / ~128 in double (128 registers per
// thread there): measured best with one-shot grids (tools/sweep.sh).  Tuning knobs: GENFFT_CUDA_WIDE_C_{F32,F64}
// forces the column count, GENFFT_CUDA_P_{F32,F64} picks the points-per-thread variant where several are compiled.
static const KernelEntry* find_kernel(int precision, long long L, bool wide) {
  const bool f32 = precision == GENFFT_CUDA_F32;
  const int want_p = wide ? env_int(f32 ? "GENFFT_CUDA_P_F32" : "GENFFT_CUDA_P_F64", 16) : 16;
  const int forced = wide ? env_int(f32 ? "GENFFT_CUDA_WIDE_C_F32" : "GENFFT_CUDA_WIDE_C_F64", 0) : 0;
  const long long min_seg = f32 ? 16 : 8, target_threads = f32 ? 256 : 128;
  const long long desired = forced ? forced : std::max(min_seg, target_threads * 16 / std::max(16LL, L));
  const KernelEntry* best = nullptr;
  auto dist = [&](const KernelEntry& e) { return std::abs(ilog2(e.C) - ilog2(desired)); };
  for (int pass = 0; pass < 2 && !best; pass++) {
    for (auto& e : registry(precision)) {
      if (e.L != L) continue;
      if (wide && !e.launch[M_COL][0]) continue;
      if (!wide && !e.launch[M_ROW][0]) continue;
      if (pass == 0 && (e.P != want_p && L >= 16)) continue;  // preferred P first
      if (!best) {
        best = &e;
      } else if (!wide) {
        if (e.C < best->C) best = &e;
      } else if (dist(e) < dist(*best) || (dist(e) == dist(*best) && e.C > best->C)) {
        best = &e;
      }
    }
  }
  return best;
}

static std::mutex g_cfg_mu;
static std::map<std::pair<int, const void*>, int> g_occupancy;  // (device, func) -> CTAs per SM

static int kernel_occupancy(const KernelEntry* k, const void* func, size_t smem, int device, int* out) {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  auto key = std::make_pair(device, func);
  auto it = g_occupancy.find(key);
  if (it != g_occupancy.end()) {
    *out = it->second;
    return GENFFT_CUDA_OK;
  }
  if (smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int n = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, func, k->threads, smem));
  if (n < 1) return fail(GENFFT_CUDA_ERR_CUDA, "kernel L=%d C=%d cannot be resident (smem %zu)", k->L, k->C, smem);
  g_occupancy[key] = n;
  *out = n;
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// twiddle tables: fp64/long-double computed, stored at the transform's precision
// (FFTTwiddle.h:46-50 evaluates cos/sin in double and stores T; same contract, exact octant symmetry)
// ------------------------------------------------------------------------------------------------
static void unit_root(unsigned long long x, unsigned long long M, long double* c, long double* s) {
  // (cos, sin)(2*pi*x/M) with x reduced to the first octant so that symmetric entries are exact mirrors
  x %= M;
  x *= 8;
  M *= 8;
  const unsigned long long q = (4 * x) / M;   // quadrant
  const unsigned long long r = x - q * (M / 4);  // [0, M/4)
  const long double two_pi = 6.283185307179586476925286766559005768L;
  long double cc, ss;
  if (r > M / 8) {
    const long double th = two_pi * (long double)(M / 4 - r) / (long double)M;
    cc = sinl(th);
    ss = cosl(th);
  } else {
    const long double th = two_pi * (long double)r / (long double)M;
    cc = cosl(th);
    ss = sinl(th);
  }
  switch (q) {
    case 0: *c = cc; *s = ss; break;
    case 1: *c = -ss; *s = cc; break;
    case 2: *c = -cc; *s = -ss; break;
    default: *c = ss; *s = -cc; break;
  }
}

static std::mutex g_tw_mu;
// (device, precision, M, step, count) -> device table of W_M^(e*step), e < count
static std::map<std::tuple<int, int, long long, long long, long long>, void*> g_tables;

static int twiddle_table(int device, int precision, long long M, long long step, long long count, const void** out) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(device, precision, M, step, count);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) {
    *out = it->second;
    return GENFFT_CUDA_OK;
  }
  const size_t es = elem_size(precision);
  std::vector<unsigned char> host(es * (size_t)count);
  for (long long e = 0; e < count; e++) {
    long double c, s;
    unit_root((unsigned long long)(e * step), (unsigned long long)M, &c, &s);
    if (precision == GENFFT_CUDA_F32) {
      float* p = reinterpret_cast<float*>(host.data()) + 2 * e;
      p[0] = (float)c;
      p[1] = (float)-s;
    } else {
      double* p = reinterpret_cast<double*>(host.data()) + 2 * e;
      p[0] = (double)c;
      p[1] = (double)-s;
    }
  }
  void* d = nullptr;
  CU_TRY(cudaMalloc(&d, host.size()));
  CU_TRY(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
  g_tables[key] = d;
  *out = d;
  return GENFFT_CUDA_OK;
}

// stage twiddles of one kernel configuration (L, P): for every radix stage s >= 1 a block [q][p] of
// W_{NS*R}^(p*q), q < R, p < NS (see tile_kernel.cuh)
static std::map<std::tuple<int, int, int, int>, void*> g_stage_tables;

static int stage_twiddle_table(int device, int precision, const KernelEntry* k, const void** out) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(device, precision, k->L, k->P);
  auto it = g_stage_tables.find(key);
  if (it != g_stage_tables.end()) {
    *out = it->second;
    return GENFFT_CUDA_OK;
  }
  const int L = k->L, P = k->P;
  const int nst = num_stages(L, P);
  const size_t count = (size_t)std::max(1, stage_tw_size(L, P));
  const size_t es = elem_size(precision);
  std::vector<unsigned char> host(es * count, 0);
  for (int s = 1; s < nst; s++) {
    const int R = stage_radix(L, P, s), NS = stage_ns(L, P, s), off = stage_tw_offset(L, P, s);
    for (int q = 0; q < R; q++)
      for (int p = 0; p < NS; p++) {
        long double c, sn;
        unit_root((unsigned long long)p * q, (unsigned long long)NS * R, &c, &sn);
        const size_t e = (size_t)off + (size_t)q * NS + p;
        if (precision == GENFFT_CUDA_F32) {
          float* t = reinterpret_cast<float*>(host.data()) + 2 * e;
          t[0] = (float)c;
          t[1] = (float)-sn;
        } else {
          double* t = reinterpret_cast<double*>(host.data()) + 2 * e;
          t[0] = (double)c;
          t[1] = (double)-sn;
        }
      }
  }
  void* d = nullptr;
  CU_TRY(cudaMalloc(&d, host.size()));
  CU_TRY(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
  g_stage_tables[key] = d;
  *out = d;
  return GENFFT_CUDA_OK;
}

// inter-pass factor W_{P*Ns}^(p*i) laid out [i][p] (see PassParams::tw_b); with -DGENFFT_TWB_TILED tile-major
// [p / C][i][p % C] for the C-column tiles of the kernel that reads it: a thread's P-1 entries are then immediate
// offsets i*C from one address, and the rows a warp reads are adjacent cache lines
static std::map<std::tuple<int, int, int, long long, int>, void*> g_pass_tables;

static int pass_stage_table(int device, int precision, int P, long long Ns, int C, const void** out) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(device, precision, P, Ns, C);
  auto it = g_pass_tables.find(key);
  if (it != g_pass_tables.end()) {
    *out = it->second;
    return GENFFT_CUDA_OK;
  }
  const size_t es = elem_size(precision);
  const size_t ntiles = (size_t)((Ns + C - 1) / C);
  const size_t count = ntiles * (size_t)P * (size_t)C;
  std::vector<unsigned char> host(es * count, 0);
  // entry (i, p) is W_{P*Ns}^(p*i), by exact index arithmetic (p*i mod P*Ns); columns beyond Ns in the last tile stay 0
  const unsigned long long M = (unsigned long long)P * (unsigned long long)Ns;
  for (int i = 0; i < P; i++)
    for (long long q = 0; q < Ns; q++) {
      long double c, sn;
      unit_root(((unsigned long long)q * (unsigned long long)i) % M, M, &c, &sn);
#ifdef GENFFT_TWB_TILED
      const size_t e = (size_t)(q / C) * (size_t)P * (size_t)C + (size_t)i * (size_t)C + (size_t)(q % C);
#else
      const size_t e = (size_t)i * (size_t)Ns + (size_t)q;
#endif
      if (precision == GENFFT_CUDA_F32) {
        float* t = reinterpret_cast<float*>(host.data()) + 2 * e;
        t[0] = (float)c;
        t[1] = (float)-sn;
      } else {
        double* t = reinterpret_cast<double*>(host.data()) + 2 * e;
        t[0] = (double)c;
        t[1] = (double)-sn;
      }
    }
  void* d = nullptr;
  CU_TRY(cudaMalloc(&d, host.size()));
  CU_TRY(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
  g_pass_tables[key] = d;
  *out = d;
  return GENFFT_CUDA_OK;
}

// whole inter-pass table of a pass with Ns * L = M small enough to stay in L2: entry [k][p] = W_M^(p*k), k < L, p < Ns
static std::map<std::tuple<int, int, long long, long long>, void*> g_direct_tables;

static int direct_pass_table(int device, int precision, long long Ns, long long L, const void** out) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(device, precision, Ns, L);
  auto it = g_direct_tables.find(key);
  if (it != g_direct_tables.end()) {
    *out = it->second;
    return GENFFT_CUDA_OK;
  }
  const size_t es = elem_size(precision);
  const unsigned long long M = (unsigned long long)Ns * (unsigned long long)L;
  std::vector<unsigned char> host(es * (size_t)M);
  for (long long k = 0; k < L; k++)
    for (long long q = 0; q < Ns; q++) {
      long double c, sn;
      unit_root(((unsigned long long)q * (unsigned long long)k) % M, M, &c, &sn);
      const size_t e = (size_t)k * (size_t)Ns + (size_t)q;
      if (precision == GENFFT_CUDA_F32) {
        float* t = reinterpret_cast<float*>(host.data()) + 2 * e;
        t[0] = (float)c;
        t[1] = (float)-sn;
      } else {
        double* t = reinterpret_cast<double*>(host.data()) + 2 * e;
        t[0] = (double)c;
        t[1] = (double)-sn;
      }
    }
  void* d = nullptr;
  CU_TRY(cudaMalloc(&d, host.size()));
  CU_TRY(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
  g_direct_tables[key] = d;
  *out = d;
  return GENFFT_CUDA_OK;
}

// two-level table for W_M^e, e < M: W = hi[e >> shift] * lo[e & (2^shift - 1)]
static int two_level_table(int device, int precision, long long M, const void** hi, const void** lo, int* shift) {
  const int lg = ilog2(M);
  const int sh = lg <= 12 ? lg : (lg + 1) / 2;
  int rc = twiddle_table(device, precision, M, 1LL << sh, M >> sh, hi);
  if (rc) return rc;
  rc = twiddle_table(device, precision, M, 1, 1LL << sh, lo);
  if (rc) return rc;
  *shift = sh;
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// sequence decomposition
// ------------------------------------------------------------------------------------------------
static long long max_single_len(int precision, bool wide) {
  if (wide) return env_int(precision == GENFFT_CUDA_F32 ? "GENFFT_CUDA_WIDE_SINGLE_F32" : "GENFFT_CUDA_WIDE_SINGLE_F64", 2048);
  return precision == GENFFT_CUDA_F32 ? 16384 : 8192;
}
static long long max_pass_len(int precision) {
  return env_int(precision == GENFFT_CUDA_F32 ? "GENFFT_CUDA_MAXLEN_F32" : "GENFFT_CUDA_MAXLEN_F64",
                 512);  // small tiles + one more pass beat 1-CTA-per-SM tiles once grids are one-shot (tools/sweep.sh)
}

static int build_seq(Seq* seq, int device, int precision, long long N, bool wide) {
  seq->N = N;
  seq->wide = wide;
  seq->passes.clear();
  if (N == 1) return GENFFT_CUDA_OK;
  std::vector<long long> lens;
  if (N <= max_single_len(precision, wide)) {
    lens.push_back(N);
  } else {
    const int lg = ilog2(N);
    const int lgmax = ilog2(max_pass_len(precision));
    const int m = (lg + lgmax - 1) / lgmax;
    int rem = lg;
    for (int s = 0; s < m; s++) {  // descending, as even as possible
      int b = (rem + (m - s) - 1) / (m - s);
      lens.push_back(1LL << b);
      rem -= b;
    }
  }
  long long Ns = 1;
  const bool multi = lens.size() > 1;
  for (size_t s = 0; s < lens.size(); s++) {
    PassSpec ps;
    ps.R = lens[s];
    ps.Ns = Ns;
    ps.k = find_kernel(precision, ps.R, wide || multi);
    if (!ps.k) return fail(GENFFT_CUDA_ERR_SIZE, "no kernel for pass length %lld", ps.R);
    int rc = stage_twiddle_table(device, precision, ps.k, &ps.tw_L);
    if (rc) return rc;
    if (Ns > 1) {
      rc = two_level_table(device, precision, Ns * ps.R, &ps.tw_hi, &ps.tw_lo, &ps.tw_shift);
      if (rc) return rc;
      // W_{P*Ns}^(p*i), i < P, p < Ns
      rc = pass_stage_table(device, precision, ps.k->P, Ns, ps.k->C, &ps.tw_b);
      if (rc) return rc;
      // small M: the whole table W_M^(p*k) is read directly (2 MiB at 2^18 single precision: L2-resident)
      if (ilog2(Ns * ps.R) <= env_int("GENFFT_CUDA_DIRECT_TW_LOG2", 18)) {
        rc = direct_pass_table(device, precision, Ns, ps.R, &ps.tw_d);
        if (rc) return rc;
      }
    }
    seq->passes.push_back(ps);
    Ns *= ps.R;
  }
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// pass emission
// ------------------------------------------------------------------------------------------------
static PassParams base_params(const PassSpec& ps, const void* in, void* out, int inverse) {
  PassParams p;
  memset(&p, 0, sizeof p);
  p.in = in;
  p.out = out;
  p.n1 = 1;
  p.n2 = 1;
  p.out_split_log2 = -1;
  p.inverse = inverse;
  p.p_mask = 0xffffffffu;
  p.tw_L = ps.tw_L;
  p.tw_hi = ps.tw_hi;
  p.tw_lo = ps.tw_lo;
  p.tw_shift = ps.tw_shift;
  p.tw_b = ps.tw_b;
  p.tw_b_stride = ps.Ns;
  p.tw_d = ps.tw_d;
  return p;
}

static uint32_t ceil_div(long long a, long long b) { return (uint32_t)((a + b - 1) / b); }

// `batch` length-N sequences, element stride 1, sequence b at b*dist
static PassParams emit_1d(const PassSpec& ps, long long N, const void* in, long long in_dist, void* out,
                          long long out_dist, long long batch, int inverse, bool brev) {
  PassParams p = base_params(ps, in, out, inverse);
  const int C = ps.k->C;
  const long long R = ps.R, Ns = ps.Ns;
  if (N == R) {  // the whole transform on chip: columns are the transforms
    p.ncols = (int)batch;
    p.n2 = ceil_div(batch, C);
    p.ntiles = p.n2;
    p.in_stride_i = 1;
    p.in_stride_c = in_dist;
    p.out_stride_k = 1;
    p.out_stride_c = out_dist;
    p.map_load = p.map_store = 1;
    p.mode = M_ROW;
    if (brev) {
      p.brev_bits = ilog2(N);
      p.g_i = 1;
      p.brev_stride = 1;
      p.mode = M_GEN;
    }
  } else if (Ns == 1) {  // first pass: y[j*R + k] = DFT_R over i of x[j + i*N/R]
    const long long cols = N / R;
    p.ncols = (int)cols;
    p.n2 = ceil_div(cols, C);
    p.ntiles = (uint32_t)(batch * p.n2);
    p.in_t0 = in_dist;
    p.out_t0 = out_dist;
    p.in_stride_i = cols;
    p.in_stride_c = 1;
    p.out_stride_k = 1;
    p.out_stride_c = R;
    p.map_load = 0;
    p.map_store = 1;
    p.mode = M_FIRST;
    if (brev) {
      p.brev_bits = ilog2(N);
      p.g_c = 1;
      p.g_i = cols;
      p.brev_stride = 1;
      p.in_stride_c = 0;
      p.mode = M_GEN;
    }
  } else {  // later pass: j = a*Ns + p;  y[a*Ns*R + p + k*Ns] = DFT_R over i of W^(p*i) x[j + i*N/R]
    const long long a_cnt = N / (R * Ns);
    p.n1 = (uint32_t)a_cnt;
    p.ncols = (int)Ns;
    p.n2 = ceil_div(Ns, C);
    p.ntiles = (uint32_t)(batch * a_cnt * p.n2);
    p.in_t0 = in_dist;
    p.in_t1 = Ns;
    p.out_t0 = out_dist;
    p.out_t1 = Ns * R;
    p.in_stride_i = N / R;
    p.in_stride_c = 1;
    p.out_stride_k = Ns;
    p.out_stride_c = 1;
    p.map_load = p.map_store = 0;
    p.p_c = 1;
    p.p_mask = (uint32_t)(Ns - 1);
    p.mode = M_COLTW;
  }
  return p;
}

// length-N transforms down the columns of an (N x cols) array, row pitches in complex elements
static PassParams emit_col(const PassSpec& ps, long long N, const void* in, long long in_pitch, void* out,
                           long long out_pitch, long long cols, int inverse, bool brev) {
  PassParams p = base_params(ps, in, out, inverse);
  const int C = ps.k->C;
  const long long R = ps.R, Ns = ps.Ns;
  p.ncols = (int)cols;
  p.n2 = ceil_div(cols, C);
  p.in_stride_c = 1;
  p.out_stride_c = 1;
  p.map_load = p.map_store = 0;
  p.mode = (Ns == 1) ? M_COL : M_COLTW;
  if (brev) p.mode = M_GEN;
  if (N == R) {
    p.ntiles = p.n2;
    p.in_stride_i = in_pitch;
    p.out_stride_k = out_pitch;
    if (brev) {
      p.brev_bits = ilog2(N);
      p.g_i = 1;
      p.brev_stride = in_pitch;
    }
  } else if (Ns == 1) {
    p.n1 = (uint32_t)(N / R);
    p.ntiles = p.n1 * p.n2;
    p.in_t1 = in_pitch;
    p.out_t1 = R * out_pitch;
    p.in_stride_i = (N / R) * in_pitch;
    p.out_stride_k = out_pitch;
    if (brev) {
      p.brev_bits = ilog2(N);
      p.g_t1 = 1;
      p.g_i = N / R;
      p.brev_stride = in_pitch;
    }
  } else {
    const long long a_cnt = N / (R * Ns);
    p.n1 = (uint32_t)Ns;
    p.ntiles = (uint32_t)(a_cnt * Ns * p.n2);
    p.in_t0 = Ns * in_pitch;
    p.in_t1 = in_pitch;
    p.out_t0 = Ns * R * out_pitch;
    p.out_t1 = out_pitch;
    p.in_stride_i = (N / R) * in_pitch;
    p.out_stride_k = Ns * out_pitch;
    p.p_t1 = 1;
    p.p_c = 0;
  }
  return p;
}

static bool strides_fit_32(const PassParams& p) {
  auto fits = [](long long v) { return v >= 0 && v < (1LL << 32); };
  return fits(p.in_stride_i) && fits(p.out_stride_k) && fits(p.tw_b_stride);
}
static void set_tile_divisors(PassParams& p) {
  const FastDiv d1 = make_fast_div(p.n1), d2 = make_fast_div(p.n2);
  p.n1_mul = d1.mul;
  p.n1_shr = d1.shr;
  p.n2_mul = d2.mul;
  p.n2_shr = d2.shr;
}

// Everything a pass launch decides before the launch itself: the compiled mode, the grid, the tile divisors.
static int resolve_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, ResolvedLaunch* r) {
  KnobScope knob_scope;
  r->grid = 0;
  r->launch = nullptr;
  if (p.ntiles == 0) return GENFFT_CUDA_OK;
  int mode = p.mode;
  const int inv = p.inverse ? 1 : 0;
  if (mode < 0 || mode >= kNumModes || !ps.k->launch[mode][inv]) mode = M_GEN;
  int occ = 1;
  // TMA prefetch (cp.async.bulk) needs 16-byte aligned rows and pays off with several tiles per CTA
  if (mode == M_ROW && ps.k->launch[M_ROWTMA][inv] && env_int("GENFFT_CUDA_TMA", 1) &&
      ((uintptr_t)p.in % 16 == 0) && ((p.in_stride_c * (long long)elem_size(plan->precision)) % 16 == 0) &&
      p.ntiles >= 4u * (uint32_t)plan->num_sms)
    mode = M_ROWTMA;
  // the compile-time modes address with 32-bit element strides (one multiply-add per access)
  if (mode != M_GEN && !strides_fit_32(p)) mode = M_GEN;
  int rc = kernel_occupancy(ps.k, ps.k->func[mode][inv], ps.k->smem_mode[mode], plan->device, &occ);
  if (rc) return rc;
  // One-shot grids by default: the hardware block scheduler then balances the load dynamically, which on this part
  // streams ~10 % faster than a persistent grid-stride loop (tools/copy_bench.cu: 6.85 vs 6.0-6.3 TB/s).  CTAs own K
  // consecutive tiles (K = 1, or GENFFT_CUDA_TMA_TILES in the TMA mode so that its prefetch pipeline has something to
  // overlap with); the persistent form is kept for SM-fraction launches.
  long long cap = (long long)plan->num_sms * occ;
  const bool frac = p.grid_frac > 0.f && p.grid_frac < 1.f;
  if (frac) cap = std::max<long long>(1, (long long)(cap * p.grid_frac));
  const bool persistent = frac || env_int("GENFFT_CUDA_PERSISTENT", 0);
  r->q = p;
  PassParams& q = r->q;
  set_tile_divisors(q);
  if (persistent) {
    q.tiles_per_cta = 0;
    r->grid = (int)std::min<long long>(p.ntiles, cap);
  } else {
    q.tiles_per_cta = mode == M_ROWTMA ? (uint32_t)std::max(1, env_int("GENFFT_CUDA_TMA_TILES", 8)) : 1u;
    r->grid = (int)std::min<long long>(((long long)p.ntiles + q.tiles_per_cta - 1) / q.tiles_per_cta, 0x7fffffffLL);
  }
  r->launch = ps.k->launch[mode][inv];
  r->mode = mode;
  return GENFFT_CUDA_OK;
}

static int launch_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, cudaStream_t stream) {
  ResolvedLaunch r;
  int rc = resolve_pass(plan, ps, p, &r);
  if (rc) return rc;
  if (!r.launch) return GENFFT_CUDA_OK;
  r.launch(r.q, r.grid, stream);
  g_launches++;
  g_mode_launches[r.mode & 15]++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

// Resolves the launches of a single-pass c2c_1d plan once (FastPath, plan.h): the per-execution host work of the
// general driver below (pass list, buffer assignment, chain search, knob lookups, occupancy lookup) costs more than
// launching a small transform (C1: N = 1024) does.
static void build_fast_path(Plan* p) {
  p->fast.valid = false;
  if (p->kind != PLAN_C2C_1D || p->seq.passes.size() != 1) return;
  if (p->grid_frac[0] < 1.f || p->grid_frac[1] < 1.f) return;
  KnobScope knob_scope;
  const PassSpec& ps = p->seq.passes[0];
  for (int inv = 0; inv < 2; inv++)
    for (int al = 0; al < 2; al++) {
      // stand-in input pointers: only their alignment enters the decisions (TMA prefetch needs 16 bytes)
      const void* in = reinterpret_cast<const void*>((uintptr_t)(al ? 4096 : 4096 + 8));
      PassParams pp = emit_1d(ps, p->n, in, p->in_dist, nullptr, p->out_dist, p->batch, inv, false);
      pp.grid_frac = p->grid_frac[1];
      ResolvedLaunch& r = p->fast.rl[inv][al];
      if (resolve_pass(p, ps, pp, &r) != GENFFT_CUDA_OK || !r.launch) return;
    }
  p->fast.knob_hash = t_knobs.knob_hash;
  p->fast.valid = true;
}

// One launch for two consecutive passes with the intermediate kept in L2 (chain_kernel.cuh).
static int launch_chain(Plan* plan, const ChainEntry* ce, ChainParams& cp, cudaStream_t stream) {
  const size_t need = 1 + (size_t)cp.ngroups;
  void* ctr = nullptr;
  {
    // one counter block per stream: two executions of the plan on different streams may run at the same time
    std::lock_guard<std::mutex> lk(plan->mu);
    Plan::ChainCtr& cc = plan->chain_ctrs[stream];
    if (cc.count < need) {
      if (cc.ptr) {
        CU_TRY(cudaStreamSynchronize(stream));  // earlier launches on this stream are the block's only users
        CU_TRY(cudaFree(cc.ptr));
        cc.ptr = nullptr;
        cc.count = 0;
      }
      const size_t cap = std::max<size_t>(need, 4096);
      if (cudaMalloc(&cc.ptr, cap * sizeof(uint32_t)) != cudaSuccess)
        return fail(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc of %zu chain counters failed", cap);
      cc.count = cap;
    }
    ctr = cc.ptr;
  }
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    auto key = std::make_pair(plan->device, ce->func);
    auto it = g_occupancy.find(key);
    if (it == g_occupancy.end()) {
      if (ce->smem > 48 * 1024)
        CU_TRY(cudaFuncSetAttribute(ce->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ce->smem));
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ce->func, ce->threads, ce->smem));
      if (occ < 1) return fail(GENFFT_CUDA_ERR_CUDA, "chain kernel cannot be resident (smem %zu)", ce->smem);
      g_occupancy[key] = occ;
    } else {
      occ = it->second;
    }
  }
  const unsigned long long total = (unsigned long long)cp.ngroups * (cp.ta + cp.tb);
  if (total == 0 || total > 0x7fffffffULL) return fail(GENFFT_CUDA_ERR_SIZE, "chain of %llu tiles", total);
  const unsigned long long resident = (unsigned long long)plan->num_sms * occ;
  // B(g) is handed out `lag` groups after A(g): far enough that A(g)'s tiles, which the resident CTAs may still be
  // working on, are done by then (no spinning), near enough that A(g)'s output is still in L2
  if (cp.lag == 0) {
    const unsigned long long per = cp.ta + cp.tb;
    cp.lag = (uint32_t)std::min<unsigned long long>(8, 1 + (3 * resident / 2 + per - 1) / per);
    cp.lag = std::max(cp.lag, 2u);
  }
  cp.lag = std::min(cp.lag, cp.ngroups);
  cp.ctr = static_cast<uint32_t*>(ctr);
  CU_TRY(cudaMemsetAsync(ctr, 0, need * sizeof(uint32_t), stream));
  ce->launch(cp, (unsigned)std::min(total, resident), stream);
  g_launches++;
  g_mode_launches[ce->ma & 15]++;
  g_mode_launches[ce->mb & 15]++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

template <typename T>
static int launch_copy_t(const CopyParams& cp, long long batch, cudaStream_t stream) {
  if (cp.rows <= 0 || cp.cols <= 0 || batch <= 0) return GENFFT_CUDA_OK;
  dim3 grid((unsigned)std::min<long long>((cp.cols + 255) / 256, 65535), (unsigned)std::min<long long>(cp.rows, 65535),
            (unsigned)batch);
  GENFFT_LAUNCH((copy_kernel<T>), grid, 256, 0, stream, cp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}
static int launch_copy(int precision, const CopyParams& cp, long long batch, cudaStream_t stream) {
  return precision == GENFFT_CUDA_F32 ? launch_copy_t<float>(cp, batch, stream) : launch_copy_t<double>(cp, batch, stream);
}

static int launch_dit(const Plan* plan, void* out, long long out_dist, const void* in, long long in_dist, int n,
                      int half, long long batch, bool real_scalar, cudaStream_t stream) {
  DitParams d;
  memset(&d, 0, sizeof d);
  d.in = in;
  d.out = out;
  d.in_dist = in_dist;
  d.out_dist = out_dist;
  d.n = n;
  d.half = half;
  d.batch = (int)batch;
  d.in_is_real_scalar = real_scalar ? 1 : 0;
  d.tw_hi = plan->dit_hi;
  d.tw_lo = plan->dit_lo;
  d.tw_shift = plan->dit_shift;
  const int work = n / 4 + 1;
  dim3 grid((unsigned)std::min((work + 255) / 256, 4096), (unsigned)std::min<long long>(batch, 65535));
  if (plan->precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((dit_kernel<float>), grid, 256, 0, stream, d);
  else
    GENFFT_LAUNCH((dit_kernel<double>), grid, 256, 0, stream, d);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// scratch
// ------------------------------------------------------------------------------------------------
static int ensure_scratch(Plan* plan, size_t bytes) {
  std::lock_guard<std::mutex> lk(plan->mu);
  if (plan->scratch_bytes >= bytes) return GENFFT_CUDA_OK;
  if (plan->scratch) {
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaFree(plan->scratch));
    plan->scratch = nullptr;
    plan->scratch_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&plan->scratch, bytes);
  if (e != cudaSuccess) return fail(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc(%zu) for scratch failed: %s", bytes, cudaGetErrorString(e));
  plan->scratch_bytes = bytes;
  return GENFFT_CUDA_OK;
}

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
  bool c2r = false;  // GENFFT_FUSED_C2R: the first pass builds the packed spectrum from the n/2+1 input bins while loading
};

static void seq_steps(const Seq& seq, bool col, std::vector<Step>& steps, bool brev_first, bool real_first) {
  KnobScope knob_scope;
  const size_t m = seq.passes.size();
  for (size_t s = 0; s < m; s++) {
    Step st;
    st.ps = &seq.passes[s];
    st.N = seq.N;
    st.col = col;
    st.safe = (m == 1) || (s == m - 1);
    if (m > 1 && env_int("GENFFT_CUDA_NO_INPLACE", 0)) st.safe = false;
    st.brev = brev_first && s == 0;
    st.real_in = real_first && s == 0;
    if (st.brev || st.real_in) st.safe = st.safe && m == 1;
    steps.push_back(st);
  }
}

// Optional override of the last pass's store: the output bin index is split at 2^part_log2 and the high
// part selects a peer buffer (fused all-to-all over NVLink) or a block at khi*part_stride.
struct FinalStore {
  int part_log2 = -1;
  void* const* peers = nullptr;
  int npeers = 0;
  long long part_stride = 0;
  long long peer_offset = 0;  // elements added to every peer pointer
};

// Optional fusion of the real-FFT split into the last pass of a multi-pass chain (M_COLTWDIT).
struct DitFuse {
  int half = 0;
  const void* dit_a = nullptr;   // W_n^p, p < Ns of the last pass
  const void* dit_tw = nullptr;  // W_{2L}^k, k < L of the last pass
};

// runs `steps`; count = batch (1D) or rows (row passes of 2D); cols = columns for column passes
static int run_chain(Plan* plan, const std::vector<Step>& steps_in, View in, View out, long long scratch_pitch,
                     size_t scratch_elems, long long count, long long cols, int inverse, cudaStream_t stream,
                     const FinalStore* fs = nullptr, const DitFuse* df = nullptr, const void* in2 = nullptr) {
  KnobScope knob_scope;
  std::vector<Step> steps(steps_in);
  if (fs && steps.size() > 1) steps.back().safe = false;  // the last pass writes elsewhere than it reads
  const size_t es = elem_size(plan->precision);
  const size_t n = steps.size();
  // non-in-place steps after the first toggle OUT <-> SCRATCH; choose the first destination so the chain ends in OUT
  int toggles = 0;
  for (size_t s = 1; s < n; s++)
    if (!steps[s].safe) toggles++;
  bool need_scratch = toggles > 0;
  const bool aliased = (in.ptr == out.ptr);
  const bool first_safe = steps[0].safe && in.pitch == out.pitch;
  const bool copy_in = aliased && !(first_safe && toggles == 0);
  // aliased input with an unsafe first pass: stage the input into a second scratch region
  size_t scratch_total = (need_scratch ? scratch_elems : 0) + (copy_in ? scratch_elems : 0);
  if (scratch_total) {
    int rc = ensure_scratch(plan, scratch_total * es);
    if (rc) return rc;
  }
  View scr{plan->scratch, scratch_pitch};
  View cur = in;
  if (copy_in) {
    View stage{(char*)plan->scratch + (need_scratch ? scratch_elems * es : 0), scratch_pitch};
    CopyParams cp;
    memset(&cp, 0, sizeof cp);
    cp.in = in.ptr;
    cp.out = stage.ptr;
    if (steps[0].col) {
      cp.rows = steps[0].N;
      cp.cols = cols;
      cp.in_stride = in.pitch;
      cp.out_stride = stage.pitch;
    } else {
      cp.rows = count;
      cp.cols = steps[0].N;
      cp.in_stride = in.pitch;
      cp.out_stride = stage.pitch;
    }
    int rc = launch_copy(plan->precision, cp, 1, stream);
    if (rc) return rc;
    cur = stage;
  }
  // ---- assign the buffers of every step ----
  struct StepIO {
    View src, dst;
    bool src_scr, dst_scr;
  };
  std::vector<StepIO> io(n);
  bool dst_is_out = (toggles % 2 == 0);
  bool cur_scr = false;
  for (size_t s = 0; s < n; s++) {
    const Step& st = steps[s];
    View dst;
    bool d_scr;
    if (s == 0) {
      dst = dst_is_out ? out : scr;
      d_scr = !dst_is_out;
    } else if (st.safe) {
      dst = cur;  // in place
      d_scr = cur_scr;
    } else {
      dst_is_out = !dst_is_out;
      dst = dst_is_out ? out : scr;
      d_scr = !dst_is_out;
    }
    io[s] = StepIO{cur, dst, cur_scr, d_scr};
    cur = dst;
    cur_scr = d_scr;
  }
  if (cur.ptr != out.ptr) return fail(GENFFT_CUDA_ERR_ARG, "internal: pass chain did not end in the output buffer");

  // parameters of step s for the units [g0, g0 + gn) of its segment (sequences of a batch / rows, or columns)
  auto make_params = [&](size_t s, long long g0, long long gn, PassParams* pp) -> int {
    const Step& st = steps[s];
    const bool col = st.col;
    const size_t src_es = st.real_in ? es / 2 : es;
    // element offset of the group inside a buffer: sequences are `pitch` apart, columns are adjacent
    auto off = [&](const View& v) { return col ? g0 : g0 * v.pitch; };
    const char* src = (const char*)io[s].src.ptr + (size_t)off(io[s].src) * src_es;
    char* dst = (char*)io[s].dst.ptr + (size_t)off(io[s].dst) * es;
    const bool peers_out = fs && fs->peers && s == n - 1;
    PassParams p = st.col ? emit_col(*st.ps, st.N, src, io[s].src.pitch, peers_out ? io[s].dst.ptr : dst, io[s].dst.pitch, gn, inverse, st.brev)
                          : emit_1d(*st.ps, st.N, src, io[s].src.pitch, peers_out ? io[s].dst.ptr : dst, io[s].dst.pitch, gn, inverse, st.brev);
#ifdef GENFFT_FUSED_C2R
    if (st.c2r) {
      p.in_real = 3;
      p.c2r_m = (uint32_t)st.N;
      p.c2r_sc = (st.N == st.ps->R) ? 0 : 1;  // single pass: columns are transforms; first of several: columns are points
      p.c2r_hi = plan->dit_hi;
      p.c2r_lo = plan->dit_lo;
      p.c2r_shift = plan->dit_shift;
      p.mode = M_GEN;
    }
#endif
    if (st.real_in) {
      p.in_real = in2 ? 2 : 1;
      p.in2 = in2 ? (const char*)in2 + (size_t)off(io[s].src) * src_es : nullptr;
      p.mode = M_GEN;
    }
    p.grid_frac = (s == n - 1) ? plan->grid_frac[1] : plan->grid_frac[0];
    if (df && s == n - 1) {
      // pair tiles: Ns/C tiles of {p} U {Ns-p} plus one tile for column 0
      p.mode = M_COLTWDIT;
      p.n2 = (uint32_t)(st.ps->Ns / st.ps->k->C) + 1u;
      p.ntiles = (uint32_t)(gn * p.n2);
      p.dit_half = df->half;
      p.dit_a = df->dit_a;
      p.dit_tw = df->dit_tw;
    }
    if (fs && s == n - 1) {
      const int twiddled_mode = p.mode;  // what the pass would be without the redirected store
      p.mode = M_GEN;
      const long long Ns = (st.N == st.ps->R) ? 1 : st.ps->Ns;
      const int sh = fs->part_log2 - ilog2(Ns);
      if (sh < 0) return fail(GENFFT_CUDA_ERR_SIZE, "part size 2^%d smaller than pass stride %lld", fs->part_log2, Ns);
      p.out_split_log2 = sh;
      if (fs->peers) {
        p.use_peers = 1;
        for (int g = 0; g < fs->npeers && g < kMaxPeers; g++)
          p.out_peer[g] = (char*)fs->peers[g] + (size_t)(fs->peer_offset + off(io[s].dst)) * es;
        // the last pass of a multi-pass transform split over 2 / 4 / 8 ranks: bin k goes to rank k >> log2(L / ranks),
        // which the compile-time peer modes resolve per register (tile_kernel.cuh, M_PEER*)
        const int pm = fs->npeers == 2 ? M_PEER2 : fs->npeers == 4 ? M_PEER4 : fs->npeers == 8 ? M_PEER8 : -1;
        if (pm >= 0 && twiddled_mode == M_COLTW && !st.brev && !st.real_in && st.ps->k->launch[pm][inverse ? 1 : 0] &&
            (1LL << sh) * fs->npeers == st.ps->R && env_int("GENFFT_CUDA_PEER_MODES", 1))
          p.mode = pm;
      } else {
        p.out_stride_khi = fs->part_stride;
      }
    }
    *pp = p;
    return GENFFT_CUDA_OK;
  };

  // L2-resident chain of the last two passes of a segment (chain_kernel.cuh).  The B tiles of a group may depend only
  // on the A tiles of the same group:
  //   * whole units (rows / transforms / column blocks) when the segment has two passes, or B is the pair-tile split
  //     pass of a real transform, or B stores to peers;
  //   * the column classes {p2 in [g*W, (g+1)*W)} of a three-pass transform N = R1*R2*R3: pass 2 (Ns = R1) writes
  //     z[a*R1*R2 + p2 + k*R1], pass 3 (Ns = R1*R2) reads z[p3 + i*R1*R2] with p3 = p2 + k*R1, so all of pass 3's
  //     columns with p3 mod R1 in the class depend exactly on pass 2's tiles of that class.
  // Returns 1 if the pair was launched, 0 if the caller has to run the passes one by one, < 0 on error.
  const long long chain_target = (long long)env_int("GENFFT_CUDA_CHAIN_KB", 4096) << 10;
  const long long chain_max = (long long)env_int("GENFFT_CUDA_CHAIN_MAX_KB", 8192) << 10;
  const bool chain_on = env_int("GENFFT_CUDA_CHAIN", 1) != 0 && plan->grid_frac[0] >= 1.f && plan->grid_frac[1] >= 1.f &&
                        !env_int("GENFFT_CUDA_PERSISTENT", 0);
  auto try_chain = [&](size_t sa, size_t sb, bool three, long long units, int* rc_out) -> bool {
    *rc_out = GENFFT_CUDA_OK;
    const Step &A = steps[sa], &B = steps[sb];
    if (A.brev || A.real_in || A.c2r || B.brev || B.real_in || B.c2r) return false;
    if (io[sb].dst.ptr == io[sa].src.ptr) return false;
    const KernelEntry *ka = A.ps->k, *kb = B.ps->k;
    if (ka->threads != kb->threads) return false;
    const bool col = A.col;
    const int cmax = std::max(ka->C, kb->C);
    ChainParams cp;
    memset(&cp, 0, sizeof cp);
    const bool intra = three && !col && !(df && sb == n - 1) && !(fs && sb == n - 1);
    if (intra) {
      const long long R1 = A.ps->Ns, R2 = A.ps->R, R3 = B.ps->R, N = A.N;
      if (B.ps->Ns != R1 * R2 || R1 * R2 * R3 != N || cmax > R1) return false;
      long long W = cmax;
      while (W * 2 <= R1 && N / R1 * (W * 2) * (long long)es <= chain_target) W *= 2;
      if (N / R1 * W * (long long)es > chain_max) return false;
      int rc = make_params(sa, 0, 1, &cp.a);
      if (!rc) rc = make_params(sb, 0, 1, &cp.b);
      if (rc) { *rc_out = rc; return false; }
      if (cp.a.mode != M_COLTW || cp.b.mode != M_COLTW) return false;
      cp.a.ncols = (int)W;
      cp.a.n2 = (uint32_t)(W / ka->C);
      cp.a.ntiles = cp.a.n1 * cp.a.n2;  // n1 = R3 blocks a
      cp.b.n1 = (uint32_t)R2;            // t1 = k: columns p3 = p2 + k*R1
      cp.b.in_t1 = R1;
      cp.b.out_t1 = R1;
      cp.b.p_t1 = (int)R1;
      cp.b.ncols = (int)W;
      cp.b.n2 = (uint32_t)(W / kb->C);
      cp.b.ntiles = cp.b.n1 * cp.b.n2;
      cp.gdiv = (uint32_t)(R1 / W);
      cp.a_in_hi = io[sa].src.pitch; cp.a_out_hi = io[sa].dst.pitch;
      cp.b_in_hi = io[sb].src.pitch; cp.b_out_hi = io[sb].dst.pitch;
      cp.a_in_lo = cp.a_out_lo = cp.b_in_lo = cp.b_out_lo = W;
      cp.a_p_lo = cp.b_p_lo = (uint32_t)W;
      cp.ngroups = (uint32_t)(units * cp.gdiv);
    } else {
      const long long unit_bytes = A.N * (long long)es;
      long long U = 1;
      while (U * 2 * unit_bytes <= chain_target) U *= 2;
      const long long umin = col ? cmax : 1;
      U = std::max(U, umin);
      while (U > umin && units % U) U /= 2;
      if (units % U || U * unit_bytes > chain_max) return false;
      int rc = make_params(sa, 0, U, &cp.a);
      if (!rc) rc = make_params(sb, 0, U, &cp.b);
      if (rc) { *rc_out = rc; return false; }
      cp.gdiv = 0x7fffffffu;
      if (col) {
        cp.a_in_lo = cp.a_out_lo = cp.b_in_lo = cp.b_out_lo = U;
      } else {
        cp.a_in_lo = U * io[sa].src.pitch; cp.a_out_lo = U * io[sa].dst.pitch;
        cp.b_in_lo = U * io[sb].src.pitch; cp.b_out_lo = U * io[sb].dst.pitch;
      }
      cp.ngroups = (uint32_t)(units / U);
    }
    const ChainEntry* ce = find_chain(plan->precision, ka, cp.a.mode, kb, cp.b.mode, inverse ? 1 : 0);
    if (!ce) return false;
    if (!strides_fit_32(cp.a) || !strides_fit_32(cp.b)) return false;
    set_tile_divisors(cp.a);
    set_tile_divisors(cp.b);
    cp.ta = cp.a.ntiles;
    cp.tb = cp.b.ntiles;
    if (!cp.ta || !cp.tb || !cp.ngroups) return false;
    cp.lag = (uint32_t)std::max(0, env_int("GENFFT_CUDA_CHAIN_LAG", 0));  // 0: chosen by launch_chain
    *rc_out = launch_chain(plan, ce, cp, stream);
    return *rc_out == GENFFT_CUDA_OK;
  };

  // Chain of the FIRST two passes of a three-pass transform N = R1*R2*R3 (knob GENFFT_CUDA_CHAIN12, off by default:
  // written without a GPU at hand, bit-identical to the unchained passes on the emulator, waiting for its measurement).
  // It serves the sequences whose last two passes cannot be chained -- the fused-split last pass of a large real
  // transform (C4), whose groups would be whole 16 MiB transforms.  Pass 1 writes y[j*R1 + k] for column j < R2*R3;
  // pass 2's block a < R3 reads y[(a + i*R3)*R1 + p], i < R2: it depends on pass 1's columns j = a + i*R3.  A group is
  // CA adjacent blocks a: pass 1's R2 tiles of columns [a0, a0 + CA) + i*R3 and pass 2's CA*(R1/CB) tiles of those
  // blocks -- R1*R2*CA points (C4: 2 MiB).
  auto try_chain_first_two = [&](size_t sa, size_t sb, long long units, int* rc_out) -> bool {
    *rc_out = GENFFT_CUDA_OK;
    const Step &A = steps[sa], &B = steps[sb];
    if (A.col || A.brev || A.real_in || A.c2r || B.brev || B.real_in || B.c2r) return false;
    if (io[sb].dst.ptr == io[sa].src.ptr || io[sb].dst.ptr == io[sb].src.ptr) return false;
    const KernelEntry *ka = A.ps->k, *kb = B.ps->k;
    if (ka->threads != kb->threads) return false;
    const long long R1 = A.ps->R, R2 = B.ps->R, N = A.N;
    if (A.ps->Ns != 1 || B.ps->Ns != R1 || N % (R1 * R2)) return false;
    const long long R3 = N / (R1 * R2);
    const long long CA = ka->C, CB = kb->C;
    if (R3 < CA || R3 % CA || R1 % CB) return false;
    if (R1 * R2 * CA * (long long)es > chain_max) return false;
    ChainParams cp;
    memset(&cp, 0, sizeof cp);
    int rc = make_params(sa, 0, 1, &cp.a);
    if (!rc) rc = make_params(sb, 0, 1, &cp.b);
    if (rc) { *rc_out = rc; return false; }
    if (cp.a.mode != M_FIRST || cp.b.mode != M_COLTW) return false;
    // A within a group: tile i < R2 is the column block a0 + i*R3 (t1 = i), one block of CA columns (t2 = 0)
    cp.a.n1 = (uint32_t)R2;
    cp.a.n2 = 1;
    cp.a.ntiles = (uint32_t)R2;
    cp.a.ncols = (int)CA;
    cp.a.in_t1 = R3;        // columns are adjacent input elements
    cp.a.out_t1 = R3 * R1;  // column j's bins start at y[j*R1]
    // B within a group: t1 = a - a0 < CA, t2 = the column blocks of p < R1 (the pass's own enumeration)
    cp.b.n1 = (uint32_t)CA;
    cp.b.ntiles = cp.b.n1 * cp.b.n2;
    cp.gdiv = (uint32_t)(R3 / CA);  // groups per transform
    cp.a_in_hi = io[sa].src.pitch; cp.a_out_hi = io[sa].dst.pitch;
    cp.b_in_hi = io[sb].src.pitch; cp.b_out_hi = io[sb].dst.pitch;
    cp.a_in_lo = CA;
    cp.a_out_lo = CA * R1;
    cp.b_in_lo = CA * cp.b.in_t1;    // in_t1 = Ns = R1 per block a
    cp.b_out_lo = CA * cp.b.out_t1;  // out_t1 = R1*R2
    cp.a_p_lo = cp.b_p_lo = 0;
    cp.ngroups = (uint32_t)(units * cp.gdiv);
    const ChainEntry* ce = find_chain(plan->precision, ka, cp.a.mode, kb, cp.b.mode, inverse ? 1 : 0);
    if (!ce) return false;
    if (!strides_fit_32(cp.a) || !strides_fit_32(cp.b)) return false;
    set_tile_divisors(cp.a);
    set_tile_divisors(cp.b);
    cp.ta = cp.a.ntiles;
    cp.tb = cp.b.ntiles;
    if (!cp.ta || !cp.tb || !cp.ngroups) return false;
    cp.lag = (uint32_t)std::max(0, env_int("GENFFT_CUDA_CHAIN_LAG", 0));
    *rc_out = launch_chain(plan, ce, cp, stream);
    return *rc_out == GENFFT_CUDA_OK;
  };
  const bool chain12_on = chain_on && env_int("GENFFT_CUDA_CHAIN12", 0) != 0;

  // ---- execute.  Consecutive passes along the same dimension form a segment.  The last two passes of a segment run
  // as one L2-resident chain when a chain kernel exists for their shapes.  (The older experiment GENFFT_CUDA_L2_GROUP_MB
  // ran a segment group by group with one launch per pass and group; it LOST -- C5 10.7 -> 13.4 ms at 64 MiB groups --
  // because those launches are single-wave and latency-bound.  It is kept for reference, off by default.) ----
  const long long l2_group_bytes = (long long)env_int("GENFFT_CUDA_L2_GROUP_MB", 0) << 20;
  size_t seg_begin = 0;
  while (seg_begin < n) {
    size_t seg_end = seg_begin + 1;
    while (seg_end < n && steps[seg_end].col == steps[seg_begin].col) seg_end++;
    const bool col = steps[seg_begin].col;
    const long long units = col ? cols : count;
    const long long unit_bytes = steps[seg_begin].N * (long long)es;
    size_t run_end = seg_end;  // passes [seg_begin, run_end) are launched one by one
    bool chained = false;
    int crc = GENFFT_CUDA_OK;
    if (seg_end - seg_begin == 3 && chain12_on && l2_group_bytes == 0 && !col &&
        steps[seg_end - 1].N * (long long)es > chain_max &&
        (((df || fs) && seg_end == n) || env_int("GENFFT_CUDA_CHAIN12", 0) >= 2)) {  // 2: also instead of the class-wise chain of passes 2+3
      // the last two passes cannot be chained (their groups would be whole transforms beyond the L2 budget): chain
      // the first two instead and stream the last one
      chained = try_chain_first_two(seg_begin, seg_begin + 1, units, &crc);
      if (crc) return crc;
      if (chained) {
        PassParams p;
        int rc = make_params(seg_end - 1, 0, units, &p);
        if (!rc) rc = launch_pass(plan, *steps[seg_end - 1].ps, p, stream);
        if (rc) return rc;
        seg_begin = seg_end;
        continue;
      }
    }
    if (seg_end - seg_begin >= 2 && chain_on && l2_group_bytes == 0) {
      // passes before the pair first, over all units
      for (size_t s = seg_begin; s + 2 < seg_end; s++) {
        PassParams p;
        int rc = make_params(s, 0, units, &p);
        if (!rc) rc = launch_pass(plan, *steps[s].ps, p, stream);
        if (rc) return rc;
      }
      chained = try_chain(seg_end - 2, seg_end - 1, seg_end - seg_begin == 3, units, &crc);
      if (crc) return crc;
      if (chained) {
        seg_begin = seg_end;
        continue;
      }
      // no chain for this pair: run the two remaining passes plainly
      for (size_t s = seg_end - 2; s < seg_end; s++) {
        PassParams p;
        int rc = make_params(s, 0, units, &p);
        if (!rc) rc = launch_pass(plan, *steps[s].ps, p, stream);
        if (rc) return rc;
      }
      seg_begin = seg_end;
      continue;
    }
    long long group = units;
    if (seg_end - seg_begin >= 2 && l2_group_bytes > 0 && !copy_in) {
      group = std::max<long long>(1, l2_group_bytes / unit_bytes);
      if (col) {  // whole column tiles
        int cmax = 1;
        for (size_t s = seg_begin; s < seg_end; s++) cmax = std::max(cmax, steps[s].ps->k->C);
        group = std::max<long long>(cmax, group / cmax * cmax);
      }
      group = std::min(group, units);
    }
    for (long long g0 = 0; g0 < units; g0 += group) {
      const long long gn = std::min(group, units - g0);
      for (size_t s = seg_begin; s < run_end; s++) {
        PassParams p;
        int rc = make_params(s, g0, gn, &p);
        if (!rc) rc = launch_pass(plan, *steps[s].ps, p, stream);
        if (rc) return rc;
      }
    }
    seg_begin = seg_end;
  }
  return GENFFT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// plan creation
// ------------------------------------------------------------------------------------------------
// A plan's tables, scratch and launch attributes live on the device it was created on.
static int enter_exec(const Plan* p) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != p->device)
    return fail(GENFFT_CUDA_ERR_ARG, "plan was created on device %d but device %d is current", p->device, dev);
  return GENFFT_CUDA_OK;
}

static int new_plan(Plan** out, PlanKind kind, int precision) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision %d", precision);
  int dev, sms;
  int rc = usable_device(&dev, &sms);
  if (rc) return rc;
  Plan* p = new genfft_cuda_plan_s();
  p->kind = kind;
  p->precision = precision;
  p->device = dev;
  p->num_sms = sms;
  *out = p;
  return GENFFT_CUDA_OK;
}

static const long long kMaxN = 1LL << 27;
// Tile counts are 32-bit and the smallest tile holds 256 points; column / batch counts are ints.  2^38 points is
// 2 TiB in single precision, far beyond the device's memory, so this only rejects nonsense before it wraps.
static const long long kMaxPoints = 1LL << 38;
static int check_volume(long long n, long long batch) {
  if (batch > 0x7fffff00LL || n * batch > kMaxPoints)
    return fail(GENFFT_CUDA_ERR_SIZE, "batch %lld x %lld points exceeds the supported volume", batch, n);
  return GENFFT_CUDA_OK;
}
// consecutive transforms of a batch must not overlap: distances below the transform's extent (or negative) would
// give overlapping or out-of-bounds stores
static int check_dist(long long batch, long long dist, long long extent, const char* what) {
  if (dist < 0 || (batch > 1 && dist < extent))
    return fail(GENFFT_CUDA_ERR_ARG, "%s = %lld is smaller than the %lld elements of one transform", what, dist, extent);
  return GENFFT_CUDA_OK;
}

// sequence + twiddles of an n-point real transform (n/2-point packed complex transform + split)
static int setup_r2c_tables(Plan* p, int precision, long long n) {
  int rc = build_seq(&p->seq, p->device, precision, n >= 2 ? n / 2 : 1, false);
  if (!rc) rc = two_level_table(p->device, precision, n >= 8 ? n : 8, &p->dit_hi, &p->dit_lo, &p->dit_shift);
  if (!rc && n >= 4 && p->seq.passes.size() == 1)  // W_n^k, k < n/2, for the fused split
    rc = twiddle_table(p->device, precision, n, 1, n / 2, &p->dit_full);
  if (!rc && p->seq.passes.size() > 1) {
    const PassSpec& last = p->seq.passes.back();
    if (last.k->launch[M_COLTWDIT][0] && last.Ns % last.k->C == 0 && last.Ns * last.R == n / 2) {
      rc = twiddle_table(p->device, precision, n, 1, last.Ns, &p->dit_a);
      if (!rc) rc = twiddle_table(p->device, precision, 2 * last.R, 1, last.R, &p->dit_b);
    }
  }
  return rc;
}

}  // namespace genfft_cuda

using namespace genfft_cuda;


extern "C" {

const char* genfft_cuda_last_error_string(void) { return g_last_error.c_str(); }

int genfft_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int d = 0; d < n; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

uint64_t genfft_cuda_launch_count(void) { return g_launches.load(); }
uint64_t genfft_cuda_debug_mode_launch_count(int mode) { return mode >= 0 && mode < 16 ? g_mode_launches[mode].load() : 0; }

// host evaluation of the kernels' division-free tile decode (tile_kernel.cuh: fast_div), for the CPU test suite
uint32_t genfft_cuda_debug_fast_div(uint32_t x, uint32_t d) {
  const FastDiv f = make_fast_div(d);
  return fast_div(x, f.mul, f.shr);
}

int genfft_cuda_plan_c2c_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch, int64_t in_dist,
                            int64_t out_dist) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(n) || n > kMaxN) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld (power of two <= 2^27 required)", (long long)n);
  if (batch < 1) return fail(GENFFT_CUDA_ERR_ARG, "batch must be >= 1");
  int rc = check_volume(n, batch);
  if (!rc) rc = check_dist(batch, in_dist ? in_dist : n, n, "in_dist");
  if (!rc) rc = check_dist(batch, out_dist ? out_dist : n, n, "out_dist");
  if (rc) return rc;
  Plan* p;
  rc = new_plan(&p, PLAN_C2C_1D, precision);
  if (rc) return rc;
  p->n = n;
  p->batch = batch;
  p->in_dist = in_dist ? in_dist : n;
  p->out_dist = out_dist ? out_dist : n;
  rc = build_seq(&p->seq, p->device, precision, n, false);
  if (rc) {
    delete p;
    return rc;
  }
  build_fast_path(p);
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_r2c_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch, int half,
                            int64_t in_dist, int64_t out_dist) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(n) || n > 2 * kMaxN) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld", (long long)n);
  if (batch < 1) return fail(GENFFT_CUDA_ERR_ARG, "batch must be >= 1");
  int rc = check_volume(n, batch);
  if (!rc) rc = check_dist(batch, in_dist ? in_dist : n, n, "in_dist");
  if (!rc) rc = check_dist(batch, out_dist ? out_dist : (half ? n / 2 + 1 : n), n == 1 ? 1 : (half ? n / 2 + 1 : n), "out_dist");
  if (rc) return rc;
  Plan* p;
  rc = new_plan(&p, PLAN_R2C_1D, precision);
  if (rc) return rc;
  p->n = n;
  p->batch = batch;
  p->half = half != 0;
  p->in_dist = in_dist ? in_dist : n;
  p->out_dist = out_dist ? out_dist : (half ? n / 2 + 1 : n);
  if (n >= 2 && (p->in_dist & 1)) {
    delete p;
    return fail(GENFFT_CUDA_ERR_ARG, "in_dist must be even (the real input is read as packed complex pairs)");
  }
  rc = setup_r2c_tables(p, precision, n);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_c2c_2d(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t height) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(width) || !is_pow2(height) || width > kMaxN || height > kMaxN)
    return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld x %lld", (long long)width, (long long)height);
  int rc = check_volume(width, height);
  if (rc) return rc;
  Plan* p;
  rc = new_plan(&p, PLAN_C2C_2D, precision);
  if (rc) return rc;
  p->width = width;
  p->height = height;
  p->n = width * height;
  rc = build_seq(&p->seq, p->device, precision, width, false);
  if (!rc) rc = build_seq(&p->seq_v, p->device, precision, height, true);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_vert(genfft_cuda_plan_t* plan, int precision, int64_t n) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(n) || n > kMaxN) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld", (long long)n);
  Plan* p;
  int rc = new_plan(&p, PLAN_VERT, precision);
  if (rc) return rc;
  p->n = n;
  rc = build_seq(&p->seq, p->device, precision, n, true);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_dit(genfft_cuda_plan_t* plan, int precision, int64_t n) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  // GetDITImpl asserts n even or n in {1, 2} (generic/fft_dit_impl_generic.inl:120); powers of two here
  if (!is_pow2(n) || n > 2 * kMaxN) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld", (long long)n);
  Plan* p;
  int rc = new_plan(&p, PLAN_DIT, precision);
  if (rc) return rc;
  p->n = n;
  rc = two_level_table(p->device, precision, n >= 8 ? n : 8, &p->dit_hi, &p->dit_lo, &p->dit_shift);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

// Launch only a fraction of the resident-CTA capacity for the passes of this plan (frac_other: all passes but the
// last; frac_last: the last pass), so that kernels of two plans running on different streams share the SMs --
// used to overlap a link-bound remote-store pass with the HBM-bound local pass of the next chunk.
int genfft_cuda_plan_set_grid_fraction(genfft_cuda_plan_t plan, double frac_other, double frac_last) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!(frac_other > 0 && frac_other <= 1 && frac_last > 0 && frac_last <= 1))
    return fail(GENFFT_CUDA_ERR_ARG, "fractions must be in (0, 1]");
  plan->grid_frac[0] = (float)frac_other;
  plan->grid_frac[1] = (float)frac_last;
  build_fast_path(plan);  // cached launches were resolved for full grids
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_destroy(genfft_cuda_plan_t plan) {
  if (!plan) return GENFFT_CUDA_OK;
  Plan* p = plan;
  if (p->scratch) cudaFree(p->scratch);
  for (auto& kv : p->chain_ctrs)
    if (kv.second.ptr) cudaFree(kv.second.ptr);
  if (p->aux) cudaFree(p->aux);
  if (p->stage_in) cudaFree(p->stage_in);
  if (p->stage_out) cudaFree(p->stage_out);
  if (p->streams_ready) {
    for (auto& s : p->streams)
      if (s) cudaStreamDestroy(s);
    for (auto& e : p->events)
      if (e) cudaEventDestroy(e);
  }
  delete plan;
  return GENFFT_CUDA_OK;
}

int64_t genfft_cuda_plan_size(genfft_cuda_plan_t plan) { return plan ? plan->n : 0; }

int genfft_cuda_plan_num_passes(genfft_cuda_plan_t plan) {
  if (!plan) return 0;
  int n = (int)plan->seq.passes.size() + (int)plan->seq_v.passes.size();
  if (plan->kind == PLAN_R2C_1D || plan->kind == PLAN_DIT) n += 1;
  return n;
}

size_t genfft_cuda_plan_scratch_bytes(genfft_cuda_plan_t plan) { return plan ? plan->scratch_bytes : 0; }

int genfft_cuda_plan_describe(genfft_cuda_plan_t plan, char* buf, size_t buflen) {
  if (!plan || !buf || !buflen) return fail(GENFFT_CUDA_ERR_ARG, "null argument");
  std::string s;
  char tmp[160];
  auto add_seq = [&](const char* name, const Seq& q) {
    snprintf(tmp, sizeof tmp, "%s N=%lld:", name, q.N);
    s += tmp;
    for (auto& ps : q.passes) {
      snprintf(tmp, sizeof tmp, " [L=%lld Ns=%lld C=%d thr=%d smem=%zu]", ps.R, ps.Ns, ps.k->C, ps.k->threads, ps.k->smem);
      s += tmp;
    }
    s += ";";
  };
  add_seq(plan->kind == PLAN_C2C_2D ? "rows" : "seq", plan->seq);
  if (plan->kind == PLAN_C2C_2D) add_seq(" cols", plan->seq_v);
  if (plan->kind == PLAN_R2C_1D)
    s += (plan->dit_full || plan->dit_a) ? " +split fused into the last pass" : " +split kernel";
  snprintf(buf, buflen, "%s", s.c_str());
  return GENFFT_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------
// device-pointer execution
// ---------------------------------------------------------------------------------------------------
}  // extern "C"

namespace genfft_cuda {
int set_error(int code, const char* msg) { return fail(code, "%s", msg); }

int exec_c2c_internal(Plan* p, void* out, const void* in, int inverse, cudaStream_t stream, bool brev, bool real_in,
                      long long batch, const void* in2) {
  KnobScope knob_scope;
  if (!p || p->kind != PLAN_C2C_1D) return fail(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (batch < 0) batch = p->batch;
  if (p->seq.passes.empty()) {  // n == 1: the transform is the identity
    if (out == in && !real_in) return GENFFT_CUDA_OK;
    CopyParams cp;
    memset(&cp, 0, sizeof cp);
    cp.in = in;
    cp.out = out;
    cp.rows = batch;
    cp.cols = 1;
    cp.in_stride = p->in_dist;
    cp.out_stride = p->out_dist;
    cp.in_real = real_in;
    if (in2) return fail(GENFFT_CUDA_ERR_SIZE, "transform_interleave needs n >= 2");
    return launch_copy(p->precision, cp, 1, stream);
  }
  // single pass, plain complex input, the plan's own batch, knobs unchanged since plan creation: the launch was
  // resolved when the plan was made (identical to what the general driver below would decide)
  if (p->fast.valid && !brev && !real_in && !in2 && batch == p->batch && (out != in || p->in_dist == p->out_dist) &&
      p->fast.knob_hash == t_knobs.knob_hash) {
    const ResolvedLaunch& r = p->fast.rl[inverse ? 1 : 0][(reinterpret_cast<uintptr_t>(in) % 16 == 0) ? 1 : 0];
    PassParams q = r.q;
    q.in = in;
    q.out = out;
    r.launch(q, r.grid, stream);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return GENFFT_CUDA_OK;
  }
  std::vector<Step> steps;
  seq_steps(p->seq, false, steps, brev, real_in);
  View vin{const_cast<void*>(in), p->in_dist}, vout{out, p->out_dist};
  return run_chain(p, steps, vin, vout, p->n, (size_t)p->n * batch, batch, 0, inverse, stream, nullptr, nullptr, in2);
}

int exec_r2c_internal(Plan* p, void* out, const void* in, cudaStream_t st, long long batch) {
  return exec_r2c_strided(p, out, in, st, batch, 0, 0);
}

// in_dist (real scalars) / out_dist (complex elements) override the plan's when non-zero (rows of a real image)
int exec_r2c_strided(Plan* pl, void* out, const void* in, cudaStream_t st, long long batch, long long in_dist_o,
                     long long out_dist_o) {
  KnobScope knob_scope;
  if (!pl || (pl->kind != PLAN_R2C_1D && pl->kind != PLAN_R2C_2D)) return fail(GENFFT_CUDA_ERR_ARG, "not an r2c plan");
  if (int rc_dev = enter_exec(pl)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (batch < 0) batch = pl->batch;
  Plan* p = pl;
  const long long P_IN = in_dist_o ? in_dist_o : pl->in_dist, P_OUT = out_dist_o ? out_dist_o : pl->out_dist;
  if (pl->n >= 2 && (P_IN & 1)) return fail(GENFFT_CUDA_ERR_ARG, "real input distance must be even");
  const long long n = p->n;
  if (n <= 2) {  // no complex sub-transform: the split reads the input directly (FFTReal.h:206-207)
    return launch_dit(p, out, P_OUT, in, n == 1 ? P_IN : P_IN / 2, (int)n, p->half, batch, n == 1, st);
  }
  // n/2 fits on chip: one kernel does the packed complex transform and the split (fused real-FFT post-process)
  if (p->seq.passes.size() == 1 && p->dit_full && p->seq.passes[0].k->launch[M_ROWDIT][0] &&
      env_int("GENFFT_CUDA_FUSED_DIT", 1) && out != in) {
    const PassSpec& ps = p->seq.passes[0];
    PassParams pp = emit_1d(ps, n / 2, in, P_IN / 2, out, P_OUT, batch, 0, false);
    pp.mode = M_ROWDIT;
    pp.dit_tw = p->dit_full;
    pp.dit_half = p->half;
    return launch_pass(p, ps, pp, st);
  }
  std::vector<Step> steps;
  seq_steps(p->seq, false, steps, false, false);
  View vin{const_cast<void*>(in), P_IN / 2}, vout{out, P_OUT};
  // multi-pass: the split is fused into the last pass, which then works on pairs of column groups
  if (p->seq.passes.size() > 1 && p->dit_a && env_int("GENFFT_CUDA_FUSED_DIT", 1)) {
    DitFuse df;
    df.half = p->half;
    df.dit_a = p->dit_a;
    df.dit_tw = p->dit_b;
    return run_chain(p, steps, vin, vout, n / 2, (size_t)(n / 2) * batch, batch, 0, 0, st, nullptr, &df);
  }
  int rc = run_chain(p, steps, vin, vout, n / 2, (size_t)(n / 2) * batch, batch, 0, 0, st);
  if (rc) return rc;
  return launch_dit(p, out, P_OUT, out, P_OUT, (int)n, p->half, batch, false, st);
}

bool plan_needs_scratch(const Plan* p) {
  return p->seq.passes.size() > 2 || p->seq_v.passes.size() > 1;
}
}  // namespace genfft_cuda

extern "C" {

int genfft_cuda_exec_c2c_dev(genfft_cuda_plan_t plan, void* out, const void* in, int inverse, void* stream) {
  return exec_c2c_internal(plan, out, in, inverse, (cudaStream_t)stream, false, false, -1, nullptr);
}

int genfft_cuda_debug_time_c2c_pairs(genfft_cuda_plan_t plan, void* out, void* mid, const void* in, int iters,
                                     void* stream, double* us_per_pair) {
  if (!us_per_pair || iters < 1) return fail(GENFFT_CUDA_ERR_ARG, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int warm = 0; warm < 2; warm++) {
    int rc = exec_c2c_internal(plan, mid, in, 0, st, false, false, -1, nullptr);
    if (!rc) rc = exec_c2c_internal(plan, out, mid, 1, st, false, false, -1, nullptr);
    if (rc) return rc;
  }
  CU_TRY(cudaStreamSynchronize(st));
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) {
    int rc = exec_c2c_internal(plan, mid, in, 0, st, false, false, -1, nullptr);
    if (!rc) rc = exec_c2c_internal(plan, out, mid, 1, st, false, false, -1, nullptr);
    if (rc) return rc;
  }
  CU_TRY(cudaStreamSynchronize(st));
  *us_per_pair = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_c2c_no_scramble_dev(genfft_cuda_plan_t plan, void* inout, int inverse, void* stream) {
  return exec_c2c_internal(plan, inout, inout, inverse, (cudaStream_t)stream, true, false, -1, nullptr);
}

int genfft_cuda_exec_c2c_real_in_dev(genfft_cuda_plan_t plan, void* out, const void* in_real, void* stream) {
  if (out == in_real) return fail(GENFFT_CUDA_ERR_ARG, "transform_real requires out != in");
  return exec_c2c_internal(plan, out, in_real, 0, (cudaStream_t)stream, false, true, -1, nullptr);
}

int genfft_cuda_exec_c2c_interleave_dev(genfft_cuda_plan_t plan, void* out, const void* in1, const void* in2,
                                        void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_1D) return fail(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in1 || !in2) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in1 || out == in2) return fail(GENFFT_CUDA_ERR_ARG, "transform_interleave requires out != in");
  return exec_c2c_internal(p, out, in1, 0, (cudaStream_t)stream, false, true, -1, in2);
}

int genfft_cuda_separate_2x_real_dev(int precision, void* out1, void* out2, const void* in, int64_t n, void* stream) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!out1 || !out2 || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (n < 1 || n > kMaxN) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld", (long long)n);
  SeparateParams sp;
  sp.in = in;
  sp.out1 = out1;
  sp.out2 = out2;
  sp.n = (int)n;
  const unsigned grid = (unsigned)std::min<long long>((n / 2 + 1 + 255) / 256, 8192);
  if (precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((separate_kernel<float>), grid, 256, 0, (cudaStream_t)stream, sp);
  else
    GENFFT_LAUNCH((separate_kernel<double>), grid, 256, 0, (cudaStream_t)stream, sp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

// plan-owned device staging of the device-pointer entry points that need a dense temporary
static int ensure_aux(Plan* p, size_t need) {
  std::lock_guard<std::mutex> lk(p->mu);
  if (p->aux_bytes >= need) return GENFFT_CUDA_OK;
  if (p->aux) {
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaFree(p->aux));
    p->aux = nullptr;
    p->aux_bytes = 0;
  }
  if (cudaMalloc(&p->aux, need) != cudaSuccess) return fail(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc(%zu) failed", need);
  p->aux_bytes = need;
  return GENFFT_CUDA_OK;
}

int genfft_cuda_plan_r2c_2d(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t height) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(width) || !is_pow2(height) || width > kMaxN || height > kMaxN || width < 2)
    return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld x %lld", (long long)width, (long long)height);
  int rc = check_volume(width, height);
  if (rc) return rc;
  Plan* p;
  rc = new_plan(&p, PLAN_R2C_2D, precision);
  if (rc) return rc;
  p->width = width;
  p->height = height;
  p->n = width;
  p->batch = height;
  p->half = 1;
  p->in_dist = width;
  p->out_dist = width;
  rc = setup_r2c_tables(p, precision, width);
  if (!rc) rc = build_seq(&p->seq_v, p->device, precision, height, true);
  if (!rc) rc = build_seq(&p->seq_h, p->device, precision, width, false);  // forward_2x: full-width complex rows
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

// RealFFT2D<T>::forward (include/genFFT/FFTReal.h:83-104): real rows -> half spectra (the reference packs row pairs
// into one complex transform and separates them, :144-165; here each row is a real transform with the split fused),
// column transforms on the width/2+1 independent columns, Hermitian completion of the remaining columns.
int genfft_cuda_exec_r2c_2d_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                                int64_t in_stride, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_R2C_2D) return fail(GENFFT_CUDA_ERR_ARG, "not an r2c_2d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return fail(GENFFT_CUDA_ERR_ARG, "RealFFT2D::forward requires out != in");
  if (out_stride < p->width || in_stride < p->width) return fail(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = exec_r2c_strided(p, out, in, st, p->height, in_stride, out_stride);
  if (rc) return rc;
  const long long vcols = p->width / 2 + 1;
  if (!p->seq_v.passes.empty()) {
    std::vector<Step> steps;
    seq_steps(p->seq_v, true, steps, false, false);
    View v{out, out_stride};
    rc = run_chain(p, steps, v, v, vcols, (size_t)vcols * p->height, 0, vcols, 0, st);
    if (rc) return rc;
  }
  if (p->width > 2) {
    MirrorParams mp;
    mp.data = out;
    mp.stride = out_stride;
    mp.w = (int)p->width;
    mp.h = (int)p->height;
    dim3 grid((unsigned)std::min<long long>((p->width / 2 + 255) / 256, 1024), (unsigned)std::min<long long>(p->height, 65535));
    if (p->precision == GENFFT_CUDA_F32)
      GENFFT_LAUNCH((mirror2d_kernel<float>), grid, 256, 0, st, mp);
    else
      GENFFT_LAUNCH((mirror2d_kernel<double>), grid, 256, 0, st, mp);
    g_launches++;
    CU_TRY(cudaGetLastError());
  }
  return GENFFT_CUDA_OK;
}

// RealFFT2D<T>::forward_2x (include/genFFT/FFTReal.h:106-118): the 2D transform of the complex image in1 + i*in2,
// i.e. two real images at the price of one complex transform.  The reference builds the rows with
// scramble_row_x2 (:167-180: every row is transform_interleave of a row of in1 and a row of in2; the offset of in2's
// second half is written with in_stride1 there -- a typo, in_stride2 is what is meant and what is done here) and
// finishes with the column transform on all `width` columns.  Here: the rows' first pass reads its real and imaginary
// parts from the two images (no interleaved copy is materialised) when both have the same row stride; otherwise, or
// with GENFFT_CUDA_2X_FUSED=0, one interleaving copy into plan-owned memory precedes the ordinary 2D pass chain.
int genfft_cuda_exec_r2c_2d_2x_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in1,
                                   int64_t in_stride1, const void* in2, int64_t in_stride2, void* stream) {
  KnobScope knob_scope;
  Plan* p = plan;
  if (!p || p->kind != PLAN_R2C_2D) return fail(GENFFT_CUDA_ERR_ARG, "not an r2c_2d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in1 || !in2) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in1 || out == in2) return fail(GENFFT_CUDA_ERR_ARG, "RealFFT2D::forward_2x requires out != in");
  if (out_stride < p->width || in_stride1 < p->width || in_stride2 < p->width)
    return fail(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t elems = (size_t)p->width * p->height;
  const bool fused = in_stride1 == in_stride2 && env_int("GENFFT_CUDA_2X_FUSED", 1);
  std::vector<Step> steps, cols;
  seq_steps(p->seq_h, false, steps, false, fused);
  seq_steps(p->seq_v, true, cols, false, false);
  steps.insert(steps.end(), cols.begin(), cols.end());
  View vout{out, out_stride};
  if (fused) {
    View vin{const_cast<void*>(in1), in_stride1};  // pitch in real scalars (the first pass reads scalars)
    return run_chain(p, steps, vin, vout, p->width, elems, p->height, p->width, 0, st, nullptr, nullptr, in2);
  }
  int rc = ensure_aux(p, elems * elem_size(p->precision));
  if (rc) return rc;
  CopyParams cp;
  memset(&cp, 0, sizeof cp);
  cp.in = in1;
  cp.in2 = in2;
  cp.out = p->aux;
  cp.rows = p->height;
  cp.cols = p->width;
  cp.in_stride = in_stride1;
  cp.in2_stride = in_stride2;
  cp.out_stride = p->width;
  cp.in_real = 2;
  rc = launch_copy(p->precision, cp, 1, st);
  if (rc) return rc;
  View vin{p->aux, p->width};
  return run_chain(p, steps, vin, vout, p->width, elems, p->height, p->width, 0, st);
}

int genfft_cuda_plan_c2r_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch, int64_t in_dist,
                            int64_t out_dist) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(n) || n < 2 || n > 2 * kMaxN)
    return fail(GENFFT_CUDA_ERR_SIZE, "unsupported size %lld (power of two >= 2 required)", (long long)n);
  if (batch < 1) return fail(GENFFT_CUDA_ERR_ARG, "batch must be >= 1");
  int rc = check_volume(n, batch);
  if (!rc) rc = check_dist(batch, in_dist ? in_dist : n / 2 + 1, n / 2 + 1, "in_dist");
  if (!rc) rc = check_dist(batch, out_dist ? out_dist : n, n, "out_dist");
  if (rc) return rc;
  Plan* p;
  rc = new_plan(&p, PLAN_C2R_1D, precision);
  if (rc) return rc;
  p->n = n;
  p->batch = batch;
  p->half = 1;
  p->in_dist = in_dist ? in_dist : n / 2 + 1;
  p->out_dist = out_dist ? out_dist : n;
  if (n >= 2 && (p->out_dist & 1)) {
    delete p;
    return fail(GENFFT_CUDA_ERR_ARG, "out_dist must be even (the real output is written as packed complex pairs)");
  }
  rc = build_seq(&p->seq, p->device, precision, n / 2, false);
  if (!rc) rc = two_level_table(p->device, precision, n, &p->dit_hi, &p->dit_lo, &p->dit_shift);  // W_n^k, exact n
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

// Unscaled inverse of RealFFT<T>::forward(half = true): out[j] = n * x[j].  `in` holds n/2+1 bins per transform.
int genfft_cuda_exec_c2r_dev(genfft_cuda_plan_t plan, void* out, const void* in, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2R_1D) return fail(GENFFT_CUDA_ERR_ARG, "not a c2r_1d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return fail(GENFFT_CUDA_ERR_ARG, "c2r requires out != in");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = p->n, M = n / 2;
  const size_t es = elem_size(p->precision);
#ifdef GENFFT_FUSED_C2R
  // variant: no pre-process pass and no staging -- the first butterfly pass reads X[s] and X[M - s] itself
  if (!p->seq.passes.empty() && env_int("GENFFT_CUDA_FUSED_C2R", 1)) {
    std::vector<Step> steps;
    seq_steps(p->seq, false, steps, false, false);
    steps[0].c2r = true;
    steps[0].safe = steps[0].safe && steps.size() == 1;
    View vin{const_cast<void*>(in), p->in_dist}, vout{out, p->out_dist / 2};
    return run_chain(p, steps, vin, vout, M, (size_t)M * p->batch, p->batch, 0, 1, st);
  }
#endif
  // stage the pre-processed spectrum Z' (M complex per transform) in plan-owned memory
  {
    int rc = ensure_aux(p, (size_t)M * p->batch * es);
    if (rc) return rc;
  }
  C2rParams cp;
  memset(&cp, 0, sizeof cp);
  cp.in = in;
  cp.out = p->aux;
  cp.in_dist = p->in_dist;
  cp.out_dist = M;
  cp.n = (int)n;
  cp.batch = (int)p->batch;
  cp.tw_hi = p->dit_hi;
  cp.tw_lo = p->dit_lo;
  cp.tw_shift = p->dit_shift;
  dim3 grid((unsigned)std::min<long long>((M + 255) / 256, 4096), (unsigned)std::min<long long>(p->batch, 65535));
  if (p->precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((c2r_pre_kernel<float>), grid, 256, 0, st, cp);
  else
    GENFFT_LAUNCH((c2r_pre_kernel<double>), grid, 256, 0, st, cp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  if (p->seq.passes.empty()) {  // M == 1: the inverse transform is the identity
    CopyParams c2;
    memset(&c2, 0, sizeof c2);
    c2.in = p->aux;
    c2.out = out;
    c2.rows = p->batch;
    c2.cols = 1;
    c2.in_stride = 1;
    c2.out_stride = p->out_dist / 2;
    return launch_copy(p->precision, c2, 1, st);
  }
  std::vector<Step> steps;
  seq_steps(p->seq, false, steps, false, false);
  View vin{p->aux, M}, vout{out, p->out_dist / 2};
  return run_chain(p, steps, vin, vout, M, (size_t)M * p->batch, p->batch, 0, 1, st);
}

int genfft_cuda_exec_r2c_dev(genfft_cuda_plan_t plan, void* out, const void* in, void* stream) {
  return exec_r2c_internal(plan, out, in, (cudaStream_t)stream, -1);
}

int genfft_cuda_exec_c2c_2d_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                                int64_t in_stride, int inverse, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_2D) return fail(GENFFT_CUDA_ERR_ARG, "not a c2c_2d plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return fail(GENFFT_CUDA_ERR_ARG, "FFT2D::transform requires out != in (fft.h:209)");
  if (out_stride < p->width || in_stride < p->width) return fail(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<Step> rows, cols;
  seq_steps(p->seq, false, rows, false, false);
  seq_steps(p->seq_v, true, cols, false, false);
  View vin{const_cast<void*>(in), in_stride}, vout{out, out_stride};
  const size_t elems = (size_t)p->width * p->height;
  if (rows.empty() && cols.empty()) {
    CopyParams cp;
    memset(&cp, 0, sizeof cp);
    cp.in = in;
    cp.out = out;
    cp.rows = 1;
    cp.cols = 1;
    return launch_copy(p->precision, cp, 1, st);
  }
  // one chain over both dimensions: row passes (count = height sequences), then column passes
  std::vector<Step> steps(rows);
  steps.insert(steps.end(), cols.begin(), cols.end());
  return run_chain(p, steps, vin, vout, p->width, elems, p->height, p->width, inverse, st);
}

int genfft_cuda_exec_vert_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                              int64_t in_stride, int64_t cols, int inverse, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_VERT) return fail(GENFFT_CUDA_ERR_ARG, "not a vert plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return fail(GENFFT_CUDA_ERR_ARG, "FFTVert::transform requires out != in (fft.h:141)");
  if (cols < 0 || out_stride < cols || in_stride < cols) return fail(GENFFT_CUDA_ERR_ARG, "bad cols/stride");
  if (cols == 0) return GENFFT_CUDA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->seq.passes.empty()) {
    CopyParams cp;
    memset(&cp, 0, sizeof cp);
    cp.in = in;
    cp.out = out;
    cp.rows = 1;
    cp.cols = cols;
    return launch_copy(p->precision, cp, 1, st);
  }
  std::vector<Step> steps;
  seq_steps(p->seq, true, steps, false, false);
  View vin{const_cast<void*>(in), in_stride}, vout{out, out_stride};
  return run_chain(p, steps, vin, vout, cols, (size_t)cols * p->n, 0, cols, inverse, st);
}

int genfft_cuda_exec_vert_no_scramble_dev(genfft_cuda_plan_t plan, void* data, int64_t stride, int64_t cols,
                                          int inverse, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_VERT) return fail(GENFFT_CUDA_ERR_ARG, "not a vert plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!data) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (cols < 0 || stride < cols) return fail(GENFFT_CUDA_ERR_ARG, "bad cols/stride");
  if (cols == 0 || p->seq.passes.empty()) return GENFFT_CUDA_OK;
  std::vector<Step> steps;
  seq_steps(p->seq, true, steps, true, false);
  View v{data, stride};
  return run_chain(p, steps, v, v, stride, (size_t)stride * p->n, 0, cols, inverse, (cudaStream_t)stream);
}

int genfft_cuda_exec_dit_dev(genfft_cuda_plan_t plan, void* out, const void* in, int half, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_DIT) return fail(GENFFT_CUDA_ERR_ARG, "not a dit plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  return launch_dit(p, out, 0, in, 0, (int)p->n, half != 0, 1, false, (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// distributed 2D building blocks (slab decomposition; the process group lives in genfft_b200/dist.py)
// ---------------------------------------------------------------------------------------------------
extern "C" {

static void* const kPeerSentinel = (void*)(uintptr_t)16;

int genfft_cuda_plan_dist_rows(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t rows, int nparts) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(width) || width > kMaxN || width < 2) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported width %lld", (long long)width);
  if (!is_pow2(nparts) || nparts > kMaxPeers || nparts > width) return fail(GENFFT_CUDA_ERR_ARG, "nparts must be a power of two <= %d", kMaxPeers);
  if (rows < 1) return fail(GENFFT_CUDA_ERR_ARG, "rows must be >= 1");
  Plan* p;
  int rc = new_plan(&p, PLAN_DIST_ROWS, precision);
  if (rc) return rc;
  p->width = width;
  p->height = rows;
  p->n = width;
  p->nparts = nparts;
  rc = build_seq(&p->seq, p->device, precision, width, false);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_dist_rows_dev(genfft_cuda_plan_t plan, void* out, void* const* out_peers, int64_t part_stride,
                                   int64_t row0, const void* in, int64_t in_stride, int inverse, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_DIST_ROWS) return fail(GENFFT_CUDA_ERR_ARG, "not a dist_rows plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!in || (!out && !out_peers)) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  const long long Wp = p->width / p->nparts;
  const size_t es = elem_size(p->precision);
  FinalStore fs;
  fs.part_log2 = ilog2(Wp);
  fs.peers = out_peers;
  fs.npeers = p->nparts;
  fs.part_stride = part_stride;
  fs.peer_offset = row0 * Wp;
  std::vector<Step> steps;
  seq_steps(p->seq, false, steps, false, false);
  View vin{const_cast<void*>(in), in_stride};
  View vout{out_peers ? kPeerSentinel : (void*)((char*)out + (size_t)(row0 * Wp) * es), Wp};
  return run_chain(p, steps, vin, vout, p->width, (size_t)p->width * p->height, p->height, 0, inverse,
                   (cudaStream_t)stream, &fs);
}

int genfft_cuda_plan_dist_cols(genfft_cuda_plan_t* plan, int precision, int64_t height, int64_t cols, int nparts) {
  if (!plan) return fail(GENFFT_CUDA_ERR_ARG, "plan is null");
  if (!is_pow2(height) || height > kMaxN || height < 2) return fail(GENFFT_CUDA_ERR_SIZE, "unsupported height %lld", (long long)height);
  if (!is_pow2(nparts) || nparts > kMaxPeers || nparts > height) return fail(GENFFT_CUDA_ERR_ARG, "nparts must be a power of two <= %d", kMaxPeers);
  if (cols < 1) return fail(GENFFT_CUDA_ERR_ARG, "cols must be >= 1");
  Plan* p;
  int rc = new_plan(&p, PLAN_DIST_COLS, precision);
  if (rc) return rc;
  p->width = cols;
  p->height = height;
  p->n = height;
  p->nparts = nparts;
  rc = build_seq(&p->seq, p->device, precision, height, true);
  if (rc) {
    delete p;
    return rc;
  }
  *plan = static_cast<genfft_cuda_plan_t>(p);
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_dist_cols_dev(genfft_cuda_plan_t plan, void* out, void* const* out_peers, int64_t out_stride,
                                   int64_t col0, void* data, int64_t stride, int inverse, void* stream) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_DIST_COLS) return fail(GENFFT_CUDA_ERR_ARG, "not a dist_cols plan");
  if (int rc_dev = enter_exec(p)) return rc_dev;
  if (!data || (!out && !out_peers)) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  std::vector<Step> steps;
  seq_steps(p->seq, true, steps, false, false);
  View vin{data, stride};
  const size_t elems = (size_t)p->width * p->height;
  if (!out_peers) {
    View vout{out, out_stride};
    return run_chain(p, steps, vin, vout, p->width, elems, 0, p->width, inverse, (cudaStream_t)stream);
  }
  FinalStore fs;
  fs.part_log2 = ilog2(p->height / p->nparts);
  fs.peers = out_peers;
  fs.npeers = p->nparts;
  fs.peer_offset = col0;
  View vout{kPeerSentinel, out_stride};
  return run_chain(p, steps, vin, vout, p->width, elems, 0, p->width, inverse, (cudaStream_t)stream, &fs);
}

// out[b*out_dist + r*out_stride + c] = in[b*in_dist + r*in_stride + c], complex elements (unpack after ncclRecv)
int genfft_cuda_copy2d_dev(int precision, void* out, int64_t out_stride, int64_t out_dist, const void* in,
                           int64_t in_stride, int64_t in_dist, int64_t rows, int64_t cols, int64_t batch,
                           void* stream) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!out || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  CopyParams cp;
  memset(&cp, 0, sizeof cp);
  cp.in = in;
  cp.out = out;
  cp.in_stride = in_stride;
  cp.out_stride = out_stride;
  cp.in_dist = in_dist;
  cp.out_dist = out_dist;
  cp.rows = rows;
  cp.cols = cols;
  return launch_copy(precision, cp, batch, (cudaStream_t)stream);
}

// data[r][c] *= W_N^((row0 + r) * c), conjugated for the inverse: the twiddle between the column and the row transforms
// of the distributed four-step 1D transform (genfft_b200/dist.py::DistFFT1D).  n_total = N = H * W.
int genfft_cuda_twiddle2d_dev(int precision, void* data, int64_t stride, int64_t rows, int64_t cols, int64_t row0,
                              int64_t n_total, int inverse, void* stream) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!data) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (rows < 0 || cols < 0 || row0 < 0 || stride < cols) return fail(GENFFT_CUDA_ERR_ARG, "bad rows/cols/stride");
  if (!is_pow2(n_total) || n_total < 8 || n_total > (1LL << 40))
    return fail(GENFFT_CUDA_ERR_SIZE, "n_total must be a power of two in [8, 2^40]");
  if (cols > n_total || row0 + rows > n_total / std::max<int64_t>(cols, 1))  // exponents (row0 + r) * c stay below N
    return fail(GENFFT_CUDA_ERR_SIZE, "(row0 + rows) * cols exceeds n_total");
  if (!rows || !cols) return GENFFT_CUDA_OK;
  int dev, sms;
  int rc = usable_device(&dev, &sms);
  if (rc) return rc;
  Twiddle2dParams tp;
  memset(&tp, 0, sizeof tp);
  rc = two_level_table(dev, precision, n_total, &tp.tw_hi, &tp.tw_lo, &tp.tw_shift);
  if (rc) return rc;
  tp.data = data;
  tp.stride = stride;
  tp.rows = rows;
  tp.cols = cols;
  tp.row0 = row0;
  tp.inverse = inverse ? 1 : 0;
  dim3 grid((unsigned)std::min<long long>((cols + 255) / 256, 1024), (unsigned)std::min<long long>(rows, 65535));
  if (precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((twiddle2d_kernel<float>), grid, 256, 0, (cudaStream_t)stream, tp);
  else
    GENFFT_LAUNCH((twiddle2d_kernel<double>), grid, 256, 0, (cudaStream_t)stream, tp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

// out[c * out_stride + r] = in[r * in_stride + c] (complex elements); out != in.
int genfft_cuda_transpose_dev(int precision, void* out, int64_t out_stride, const void* in, int64_t in_stride,
                              int64_t rows, int64_t cols, void* stream) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!out || !in || out == in) return fail(GENFFT_CUDA_ERR_ARG, "null or aliased buffer");
  if (rows < 0 || cols < 0 || in_stride < cols || out_stride < rows) return fail(GENFFT_CUDA_ERR_ARG, "bad rows/cols/stride");
  if (!rows || !cols) return GENFFT_CUDA_OK;
  TransposeParams tp;
  tp.in = in;
  tp.out = out;
  tp.in_stride = in_stride;
  tp.out_stride = out_stride;
  tp.rows = rows;
  tp.cols = cols;
  const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  const unsigned grid = (unsigned)std::min<long long>(tiles, 1LL << 20);
  if (precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((transpose_kernel<float>), grid, 256, 0, (cudaStream_t)stream, tp);
  else
    GENFFT_LAUNCH((transpose_kernel<double>), grid, 256, 0, (cudaStream_t)stream, tp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

// Stream-ordered barrier across the ranks of a process group over IPC-mapped flag arrays (aux_kernels.cuh).
int genfft_cuda_peer_barrier_dev(void* const* peer_flags, int rank, int world, uint32_t epoch, void* stream) {
  if (!peer_flags || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(GENFFT_CUDA_ERR_ARG, "bad peer barrier arguments");
  static_assert(sizeof(PeerBarrierParams{}.peer_flags) / sizeof(void*) == kMaxPeers, "flag arrays per barrier");
  PeerBarrierParams p;
  memset(&p, 0, sizeof p);
  for (int r = 0; r < world; r++) {
    if (!peer_flags[r]) return fail(GENFFT_CUDA_ERR_ARG, "null flag array");
    p.peer_flags[r] = static_cast<uint32_t*>(peer_flags[r]);
  }
  p.rank = rank;
  p.world = world;
  p.epoch = epoch;
  GENFFT_LAUNCH((peer_barrier_kernel), 1, 32, 0, (cudaStream_t)stream, p);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

int genfft_cuda_memset_dev(void* ptr, int value, size_t bytes) {
  CU_TRY(cudaMemset(ptr, value, bytes));
  return GENFFT_CUDA_OK;
}

int genfft_cuda_malloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(GENFFT_CUDA_ERR_ARG, "null");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return GENFFT_CUDA_OK;
}
int genfft_cuda_free(void* ptr) {
  CU_TRY(cudaFree(ptr));
  return GENFFT_CUDA_OK;
}
int genfft_cuda_ipc_get_handle(void* ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, 64);
  return GENFFT_CUDA_OK;
}
int genfft_cuda_ipc_open_handle(void** ptr, const unsigned char handle[64]) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GENFFT_CUDA_OK;
}
int genfft_cuda_ipc_close_handle(void* ptr) {
  CU_TRY(cudaIpcCloseMemHandle(ptr));
  return GENFFT_CUDA_OK;
}

}  // extern "C"
