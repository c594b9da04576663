// libgenfft_cuda: the multi-pass driver (host side) -- a chain of Stockham passes over the buffers
// IN -> {OUT, SCRATCH} -> OUT, with the last two passes of a segment run as one L2-resident launch (chain_kernel.cuh)
// and the optional fused stores of the last pass (peer buffers of the distributed transforms, the real-FFT split).
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



void seq_steps(const Seq& seq, bool col, std::vector<Step>& steps, bool brev_first, bool real_first) {
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



// runs `steps`; count = batch (1D) or rows (row passes of 2D); cols = columns for column passes
int run_chain(Plan* plan, const std::vector<Step>& steps_in, View in, View out, long long scratch_pitch,
                     size_t scratch_elems, long long count, long long cols, int inverse, cudaStream_t stream,
                     const FinalStore* fs, const DitFuse* df, const void* in2) {
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
    if (st.c2r) {
      p.in_real = 3;
      p.c2r_m = (uint32_t)st.N;
      p.c2r_sc = (st.N == st.ps->R) ? 0 : 1;  // single pass: columns are transforms; first of several: columns are points
      p.c2r_hi = plan->dit_hi;
      p.c2r_lo = plan->dit_lo;
      p.c2r_shift = plan->dit_shift;
      p.mode = M_GEN;
    }
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
            (1LL << sh) * fs->npeers == st.ps->R && env_int("GENFFT_CUDA_PEER_MODES", 1)) {
          p.mode = pm;
          if (fs->tw_n >= 8 && fs->tw_n <= (1LL << 31) && st.col && env_int("GENFFT_CUDA_FUSED_1D_TWIDDLE", 1)) {
            int rc = two_level_table(plan->device, plan->precision, fs->tw_n, &p.tw2_hi, &p.tw2_lo, &p.tw2_shift);
            if (rc) return rc;
            p.tw2_col0 = (uint32_t)(fs->tw_col0 + g0);
            if (fs->fused) *fs->fused = true;
          }
        }
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


}  // namespace genfft_cuda
