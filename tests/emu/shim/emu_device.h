// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT (see cuda_runtime.h in this directory).
// Host restatement of the inline-PTX helpers of genfft_b200/csrc/tile_kernel.cuh and chain_kernel.cuh.  Included from
// inside namespace genfft_cuda, at the place of the PTX block, when GENFFT_EMU is defined.

// mbarrier + cp.async.bulk.  The barrier word holds the phase bit (bit 0) and the bytes still expected (upper half);
// a bulk copy is a memcpy that completes at issue and flips the phase when nothing is pending any more; a waiter polls,
// giving the other threads of the CTA a turn (in reverse / shuffled thread order the issuing thread may run last).
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  *bar = (*bar & 1ull) | ((uint64_t)bytes << 32);
  if (bytes == 0) *bar ^= 1ull;
}
inline void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  memcpy(smem_dst, gmem_src, bytes);
  const uint64_t pending = (*bar >> 32) - bytes;
  *bar = (*bar & 1ull) | (pending << 32);
  if (pending == 0) *bar ^= 1ull;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while ((uint32_t)(*bar & 1ull) == parity) ::genfft_emu::spin_yield();
}

// base[stride * k]: `stride` is a 32-bit element stride, as in the PTX form (mad.wide.u32)
template <int OP, typename V>
inline V ld_strided(const V* base, uint32_t stride, uint32_t k) { return base[(unsigned long long)stride * k]; }
template <int OP, typename V>
inline void st_strided(V* base, uint32_t stride, uint32_t k, const V& v) { base[(unsigned long long)stride * k] = v; }
template <typename V>
inline V ldg_strided(const V* base, uint32_t stride, uint32_t k) { return base[(unsigned long long)stride * k]; }

// pass chains: CTAs run one after another, so the counters are plain memory
inline uint32_t ld_acquire_gpu(const uint32_t* p) { return *p; }
inline uint32_t ld_relaxed_gpu(const uint32_t* p) { return *p; }
inline void red_release_gpu_add(uint32_t* p, uint32_t v) { *p += v; }
