// libgenfft_cuda: the planner (host side) -- errors, knobs, kernel registry, twiddle tables, sequence decomposition,
// pass emission and launches.
//
// The reference builds an "impl" per size from a factory switch (include/genFFT/x86/fft_float_impl_x86.inl:464-497)
// whose constructor chain computes one twiddle table per radix-2 level (include/genFFT/FFTTwiddle.h:44-51).
// Here a plan is a short list of Stockham passes (1 for sizes that fit on chip, 2-3 above), each a
// launch of the tile kernel with its own addressing, plus fp64-computed twiddle tables stored at the
// transform's precision.
#include <cuda_runtime.h>

#include <algorithm>
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

#include "plan_internal.h"

namespace genfft_cuda {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_mode_launches[16];  // per compiled addressing mode (tests: "the specialised mode really ran")

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}


size_t elem_size(int precision) { return precision == GENFFT_CUDA_F32 ? 8 : 16; }

bool is_pow2(long long n) { return n >= 1 && (n & (n - 1)) == 0; }
int ilog2(long long n) {
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

}  // namespace

// every function that reads knobs at execution time opens a scope; only the outermost one revalidates
KnobScope::KnobScope() {
  if (t_knobs.depth++ == 0) knobs_validate();
}
KnobScope::~KnobScope() { t_knobs.depth--; }
uint64_t current_knob_hash() { return t_knobs.knob_hash; }

int env_int(const char* name, int dflt) {
  if (t_knobs.depth == 0) knobs_validate();  // plan-creation-time reads
  const size_t len = strlen(name);
  for (const char* ent : t_knobs.entries)
    if (strncmp(ent, name, len) == 0 && ent[len] == '=') return ent[len + 1] ? atoi(ent + len + 1) : dflt;
  return dflt;
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
int usable_device(int* dev_out, int* sms_out) {
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

const ChainEntry* find_chain(int precision, const KernelEntry* ka, int ma, const KernelEntry* kb, int mb, int inv) {
  static std::vector<ChainEntry> chains;
  static std::once_flag once;
  std::call_once(once, [] {
    register_chains_0(chains);
    register_chains_1(chains);
    register_chains_2(chains);
    register_chains_3(chains);
  });
  if (ka->P != kb->P) return nullptr;
  const int prec = precision == GENFFT_CUDA_F32 ? 0 : 1;
  for (auto& e : chains)
    if (e.precision == prec && e.p == ka->P && e.la == ka->L && e.ca == ka->C && e.ma == ma && e.lb == kb->L && e.cb == kb->C &&
        e.mb == mb && e.inv == inv)
      return &e;
  return nullptr;
}

// Kernel shape for a length-L pass.  Narrow (contiguous batched) use takes the fewest sequences per CTA.  Wide (column)
// use wants row segments of at least 128 bytes and CTAs of ~256 threads in float / ~128 in double (128 registers per
// thread there): measured best with one-shot grids (tools/sweep.sh).  Tuning knobs: GENFFT_CUDA_WIDE_C_{F32,F64}
// forces the column count.
const KernelEntry* find_kernel(int precision, long long L, bool wide) {
  const bool f32 = precision == GENFFT_CUDA_F32;
  const int want_p = 16;  // every compiled shape holds 16 points per thread (8-point double shapes lost twice)
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

int kernel_occupancy(const KernelEntry* k, const void* func, size_t smem, int device, int* out) {
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

int twiddle_table(int device, int precision, long long M, long long step, long long count, const void** out) {
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

int stage_twiddle_table(int device, int precision, const KernelEntry* k, const void** out) {
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

// inter-pass factor W_{P*Ns}^(p*i) laid out [i][p] (see PassParams::tw_b): lanes read consecutive p
static std::map<std::tuple<int, int, int, long long, int>, void*> g_pass_tables;

int pass_stage_table(int device, int precision, int P, long long Ns, int C, const void** out) {
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
      const size_t e = (size_t)i * (size_t)Ns + (size_t)q;
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

// two-level table for W_M^e, e < M: W = hi[e >> shift] * lo[e & (2^shift - 1)]
int two_level_table(int device, int precision, long long M, const void** hi, const void** lo, int* shift) {
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

int build_seq(Seq* seq, int device, int precision, long long N, bool wide, long long max_single) {
  seq->N = N;
  seq->wide = wide;
  seq->passes.clear();
  if (N == 1) return GENFFT_CUDA_OK;
  std::vector<long long> lens;
  if (N <= (max_single > 0 ? max_single : max_single_len(precision, wide))) {
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
  return p;
}

static uint32_t ceil_div(long long a, long long b) { return (uint32_t)((a + b - 1) / b); }

// `batch` length-N sequences, element stride 1, sequence b at b*dist
PassParams emit_1d(const PassSpec& ps, long long N, const void* in, long long in_dist, void* out,
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
PassParams emit_col(const PassSpec& ps, long long N, const void* in, long long in_pitch, void* out,
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

bool strides_fit_32(const PassParams& p) {
  auto fits = [](long long v) { return v >= 0 && v < (1LL << 32); };
  return fits(p.in_stride_i) && fits(p.out_stride_k) && fits(p.tw_b_stride);
}
void set_tile_divisors(PassParams& p) {
  const FastDiv d1 = make_fast_div(p.n1), d2 = make_fast_div(p.n2);
  p.n1_mul = d1.mul;
  p.n1_shr = d1.shr;
  p.n2_mul = d2.mul;
  p.n2_shr = d2.shr;
}

// Everything a pass launch decides before the launch itself: the compiled mode, the grid, the tile divisors.
int resolve_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, ResolvedLaunch* r) {
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
  // Programmatic dependent launch (fft_tile_kernel) for grids of at most one wave, where the launch latency is the
  // cost: a forward + inverse pair of N = 1024 from a C loop 8.19 -> 5.15 us, N = 16384 24.6 -> 20.3 us; large grids
  // gain nothing and C2 measured 4 % slower with it (profiles/r02_ab_programmatic_dependent_launch.log).
  // GENFFT_CUDA_PDL: 0 never, 1 (default) single-wave grids, 2 always.
  {
    const int pdl = env_int("GENFFT_CUDA_PDL", 1);
    q.pdl = pdl >= 2 || (pdl == 1 && (long long)r->grid <= (long long)plan->num_sms * occ) ? 1 : 0;
  }
  return GENFFT_CUDA_OK;
}

int launch_pass(const Plan* plan, const PassSpec& ps, const PassParams& p, cudaStream_t stream) {
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
void build_fast_path(Plan* p) {
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
int launch_chain(Plan* plan, const ChainEntry* ce, ChainParams& cp, cudaStream_t stream) {
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
    // B stores to the peers (distributed transforms): its tiles are NVLink-bound and hold their CTAs several times
    // longer than A's, so A needs more slack to stay ahead -- otherwise B tiles spin on their group counter instead of
    // feeding the link.  Measured on 2 B200s (profiles/r02_c5_dist_2gpu_knob_sweep*.jsonl): lag 5 (the formula) 6.96 ms,
    // 6 6.37, 8 6.26, 12 6.27, 16 6.29, 24 6.34 for C5 in natural order; 8 groups of ~4 MiB stay L2-resident.
    if (cp.b.use_peers) cp.lag = std::max(cp.lag, 8u);
  }
  cp.lag = std::min(cp.lag, cp.ngroups);
  // One counter block per stream: two executions of the plan on different streams may run at the same time.  The block
  // is picked, zeroed and handed to the launch under the plan's lock: another thread that recycles the blocks (below)
  // must not free one between these steps.
  const size_t need = 1 + (size_t)cp.ngroups;
  void* ctr = nullptr;
  std::lock_guard<std::mutex> lk(plan->mu);
  if (plan->chain_ctrs.size() >= 32 && !plan->chain_ctrs.count(stream)) {
    // a plan that has seen many (short-lived) streams: drop the blocks of the others once the device is idle
    CU_TRY(cudaDeviceSynchronize());
    for (auto& kv : plan->chain_ctrs)
      if (kv.second.ptr) cudaFree(kv.second.ptr);
    plan->chain_ctrs.clear();
  }
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
  cp.ctr = static_cast<uint32_t*>(ctr);
  CU_TRY(cudaMemsetAsync(ctr, 0, need * sizeof(uint32_t), stream));
  // GENFFT_CUDA_CHAIN_GRID_PCT: experiment knob, share of the resident-CTA capacity the persistent chain grid uses
  const unsigned long long cap = std::max<unsigned long long>(1, resident * (unsigned)std::min(100, std::max(1, env_int("GENFFT_CUDA_CHAIN_GRID_PCT", 100))) / 100);
  ce->launch(cp, (unsigned)std::min(total, cap), stream);
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
int launch_copy(int precision, const CopyParams& cp, long long batch, cudaStream_t stream) {
  return precision == GENFFT_CUDA_F32 ? launch_copy_t<float>(cp, batch, stream) : launch_copy_t<double>(cp, batch, stream);
}

int launch_dit(const Plan* plan, void* out, long long out_dist, const void* in, long long in_dist, int n,
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
int ensure_scratch(Plan* plan, size_t bytes) {
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


}  // namespace genfft_cuda
