// libgenfft_cuda: the extern "C" boundary (include/genfft_cuda.h) -- plan creation and execution on device pointers.
// The entry points replace the reference's back-end factories genfft::impl_x86_dispatch::Get{,Vert,DIT}Impl
// (include/genFFT/x86/fft_x86_dispatch.h:33-40, src/fft_x86_dispatch.cpp:64-139) and the virtual calls on the objects
// they return (include/genFFT/FFTLevel.h:43-59,102-118; include/genFFT/FFTDIT.h:43-49).
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
      p->fast.knob_hash == current_knob_hash()) {
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
  // no pre-process pass and no staging: the first butterfly pass reads X[s] and X[M - s] itself (in_real == 3)
  if (!p->seq.passes.empty()) {
    std::vector<Step> steps;
    seq_steps(p->seq, false, steps, false, false);
    steps[0].c2r = true;
    steps[0].safe = steps[0].safe && steps.size() == 1;
    View vin{const_cast<void*>(in), p->in_dist}, vout{out, p->out_dist / 2};
    return run_chain(p, steps, vin, vout, M, (size_t)M * p->batch, p->batch, 0, 1, st);
  }
  // n == 2 (M == 1, no butterfly pass): the pre-process kernel alone, staged in plan-owned memory
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
  // M == 1: the inverse transform is the identity
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
  // Rows that would fit on chip in one pass are still split in two when they are sent to peers: the one-pass kernels
  // of 4096+ points are one or two big CTAs per SM whose loads and NVLink stores do not overlap, while a two-pass chain
  // runs the HBM-bound pass and the link-bound storing pass (compile-time peer mode) side by side
  // (GENFFT_CUDA_DIST_ROWS_SINGLE: longest row still done in one pass).
  rc = build_seq(&p->seq, p->device, precision, width, false,
                 nparts > 1 ? std::max(2, env_int("GENFFT_CUDA_DIST_ROWS_SINGLE", 16384)) : 0);
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

static int exec_dist_cols(genfft_cuda_plan_t plan, void* out, void* const* out_peers, int64_t out_stride, int64_t col0,
                          void* data, int64_t stride, int inverse, int64_t tw_n, int* fused, void* stream);

int genfft_cuda_exec_dist_cols_dev(genfft_cuda_plan_t plan, void* out, void* const* out_peers, int64_t out_stride,
                                   int64_t col0, void* data, int64_t stride, int inverse, void* stream) {
  return exec_dist_cols(plan, out, out_peers, out_stride, col0, data, stride, inverse, 0, nullptr, stream);
}

// The column transforms of the distributed four-step 1D transform with the twiddle W_n^(kr*c) that follows them fused
// into the peer-storing pass.  *fused = 1 when the stores carried it, 0 when the caller still has to run
// genfft_cuda_twiddle2d_dev on the receiving slab (shapes the compile-time peer modes do not cover).
int genfft_cuda_exec_dist_cols_tw_dev(genfft_cuda_plan_t plan, void* const* out_peers, int64_t out_stride, int64_t col0,
                                      void* data, int64_t stride, int inverse, int64_t n_total, int* fused, void* stream) {
  if (!out_peers || !fused) return fail(GENFFT_CUDA_ERR_ARG, "null argument");
  if (!is_pow2(n_total)) return fail(GENFFT_CUDA_ERR_SIZE, "n_total must be a power of two");
  return exec_dist_cols(plan, nullptr, out_peers, out_stride, col0, data, stride, inverse, n_total, fused, stream);
}

// rows x width slab -> the peers' (H x width/nparts) column blocks at rows [row0, row0 + rows): the first global
// transpose of the distributed four-step 1D transform in one launch (16-byte accesses)
int genfft_cuda_scatter_cols_dev(int precision, void* const* peers, int nparts, int64_t row0, const void* in,
                                 int64_t in_stride, int64_t rows, int64_t width, void* stream) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return fail(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!peers || !in) return fail(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (!is_pow2(nparts) || nparts > kMaxPeers || !is_pow2(width) || width < 2 * nparts || rows < 0 || row0 < 0 ||
      in_stride < width)
    return fail(GENFFT_CUDA_ERR_ARG, "bad scatter arguments");
  if (precision == GENFFT_CUDA_F32 && (((uintptr_t)in & 15) || (in_stride & 1)))
    return fail(GENFFT_CUDA_ERR_ARG, "scatter needs 16-byte aligned rows");
  if (!rows) return GENFFT_CUDA_OK;
  ScatterParams sp;
  memset(&sp, 0, sizeof sp);
  sp.in = in;
  for (int g = 0; g < nparts; g++) {
    if (!peers[g] || ((uintptr_t)peers[g] & 15)) return fail(GENFFT_CUDA_ERR_ARG, "null or misaligned peer block");
    sp.peer[g] = peers[g];
  }
  sp.in_stride = in_stride;
  sp.rows = rows;
  sp.width = width;
  sp.row0 = row0;
  sp.wp_log2 = ilog2(width / nparts);
  const long long vecs = rows * (width / (precision == GENFFT_CUDA_F32 ? 2 : 1));
  const unsigned grid = (unsigned)std::min<long long>((vecs + 255) / 256, 148LL * 64);
  if (precision == GENFFT_CUDA_F32)
    GENFFT_LAUNCH((scatter_cols_kernel<float>), grid, 256, 0, (cudaStream_t)stream, sp);
  else
    GENFFT_LAUNCH((scatter_cols_kernel<double>), grid, 256, 0, (cudaStream_t)stream, sp);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return GENFFT_CUDA_OK;
}

static int exec_dist_cols(genfft_cuda_plan_t plan, void* out, void* const* out_peers, int64_t out_stride, int64_t col0,
                          void* data, int64_t stride, int inverse, int64_t tw_n, int* fused, void* stream) {
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
  bool did = false;
  fs.tw_n = tw_n;
  fs.tw_col0 = col0;
  fs.fused = &did;
  View vout{kPeerSentinel, out_stride};
  int rc = run_chain(p, steps, vin, vout, p->width, elems, 0, p->width, inverse, (cudaStream_t)stream, &fs);
  if (fused) *fused = did ? 1 : 0;
  return rc;
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
