// Scratch microbenchmark: can tensor-map TMA (cp.async.bulk.tensor.2d) stage the STRIDED tiles of a column pass
// faster / cheaper than per-thread LDG?  The tile is `rows` row segments of `seg` bytes, `pitch` bytes apart -- the
// access pattern of M_FIRST / M_COLTW / M_COL in genfft_b200/csrc/tile_kernel.cuh.  Copy only, no butterflies:
//   ldg       : one-shot grid, a CTA loads its tile with 16-byte LDGs into registers and stores it (what the pass
//               kernels do today, minus the math)
//   tma_ld    : persistent CTAs, tile i+1.. fetched by ONE cp.async.bulk.tensor per tile into a shared-memory ring
//               (mbarrier complete_tx), threads read the tile from shared memory and store it with 16-byte STGs
//   tma_ld_st : the same ring, the tile is written back by a bulk tensor store (no thread touches the data)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tma_tile_bench.cu -o tools/tma_tile_bench
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- baseline: per-thread 16-byte loads, one tile per CTA (one-shot grid) ----
template <int VEC>
__global__ void __launch_bounds__(256) ldg_tile_copy(const char* __restrict__ in, char* __restrict__ out, int rows, int seg,
                                                     long long pitch, int tiles_per_row) {
  const int lanes = seg / 16;
  const int lr = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = blockDim.x / lanes;
  const long long tile = blockIdx.x;
  const long long blk = tile / tiles_per_row, t = tile % tiles_per_row;
  const long long base = blk * (long long)rows * pitch + t * seg + lr * 16;
  int4 v[VEC];
  for (int r = r0; r < rows; r += rstep * VEC) {
#pragma unroll
    for (int k = 0; k < VEC; k++)
      if (r + k * rstep < rows) v[k] = *reinterpret_cast<const int4*>(in + base + (long long)(r + k * rstep) * pitch);
#pragma unroll
    for (int k = 0; k < VEC; k++)
      if (r + k * rstep < rows) *reinterpret_cast<int4*>(out + base + (long long)(r + k * rstep) * pitch) = v[k];
  }
}

// ---- TMA ring: STAGES tiles in flight per CTA ----
template <int STAGES, bool TMA_STORE>
__global__ void __launch_bounds__(256) tma_tile_copy(const __grid_constant__ CUtensorMap in_map,
                                                     const __grid_constant__ CUtensorMap out_map, char* __restrict__ out,
                                                     int rows, int seg, long long pitch, int tiles_per_row,
                                                     long long ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  const uint32_t tile_bytes = (uint32_t)rows * (uint32_t)seg;
  const int seg_el = seg / 8;  // tensor-map elements are 8 bytes (one complex64)
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long tile, int s) {
    const long long blk = tile / tiles_per_row, t = tile % tiles_per_row;
    mbar_arrive_expect_tx(&full[s], tile_bytes);
    tma_load_2d(smem + (size_t)s * tile_bytes, &in_map, (int)(t * seg_el), (int)(blk * rows), &full[s]);
  };
  // prologue: fill the ring
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; s++) {
      const long long tile = blockIdx.x + (long long)s * gridDim.x;
      if (tile < ntiles) issue(tile, s);
    }
  const int lanes = seg / 16;
  const int lr = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = blockDim.x / lanes;
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int s = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1u;
    mbar_wait(&full[s], parity);
    const long long blk = tile / tiles_per_row, t = tile % tiles_per_row;
    unsigned char* buf = smem + (size_t)s * tile_bytes;
    if constexpr (TMA_STORE) {
      if (threadIdx.x == 0) {
        tma_store_2d(&out_map, (int)(t * seg_el), (int)(blk * rows), buf);
        bulk_commit();
        bulk_wait_read<0>();  // the slot may be overwritten once the store has READ it
      }
    } else {
      const long long base = blk * (long long)rows * pitch + t * seg + lr * 16;
      for (int r = r0; r < rows; r += rstep) {
        const int4 v = *reinterpret_cast<const int4*>(buf + (size_t)r * seg + lr * 16);
        *reinterpret_cast<int4*>(out + base + (long long)r * pitch) = v;
      }
    }
    __syncthreads();  // every thread is done with slot s
    const long long next = tile + (long long)STAGES * gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) {
      fence_proxy_async();  // order the generic-proxy reads of the slot before the async-proxy overwrite
      issue(next, s);
    }
  }
  if (TMA_STORE && threadIdx.x == 0) bulk_wait_all();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) {
    printf("cuTensorMapEncodeTiled not available\n");
    exit(1);
  }
  return (EncodeFn)fn;
}

static CUtensorMap make_map(EncodeFn enc, void* base, long long pitch, long long total_rows, int seg, int rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)(pitch / 8), (cuuint64_t)total_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {(cuuint32_t)(seg / 8), (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d (pitch %lld rows %lld seg %d box rows %d)\n", (int)r, pitch, total_rows, seg, rows);
    exit(1);
  }
  return m;
}

__global__ void fill_pattern(unsigned long long* p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1);
}

// resident CTAs per SM of a ring of STAGES tiles (at most 4, shared memory permitting)
template <int STAGES, bool TMA_STORE>
static float run_tma(const CUtensorMap& im, const CUtensorMap& om, char* b, int rows, int seg, long long pitch, int tpr,
                     long long ntiles, int sms, cudaEvent_t e0, cudaEvent_t e1) {
  const size_t smem = (size_t)STAGES * rows * seg;
  if (smem > 227 * 1024) return -1.f;  // ring does not fit (prints a negative rate)
  const int ctas_per_sm = (int)((220 * 1024) / smem) < 1 ? 1 : ((220 * 1024) / smem > 4 ? 4 : (int)((220 * 1024) / smem));
  CK(cudaFuncSetAttribute(tma_tile_copy<STAGES, TMA_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float ms = 0;
  for (int it = 0; it < 3; it++) {
    CK(cudaEventRecord(e0));
    tma_tile_copy<STAGES, TMA_STORE><<<sms * ctas_per_sm, 256, smem>>>(im, om, b, rows, seg, pitch, tpr, ntiles);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  CK(cudaGetLastError());
  return ms;
}

int main() {
  const long long bytes = 1LL << 30;
  char *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes));
  CK(cudaMemset(b, 0, bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  EncodeFn enc = get_encode();
  // fill `a` with a position-dependent pattern so that misplaced tiles are caught
  fill_pattern<<<sms * 8, 256>>>((unsigned long long*)a, bytes / 8);
  CK(cudaDeviceSynchronize());
  auto verify = [&](const char* what) {
    // compare 4 MiB at the start, middle and end
    static unsigned char ha[1 << 22], hb[1 << 22];
    for (long long off : {0LL, bytes / 2, bytes - (1LL << 22)}) {
      CK(cudaMemcpy(ha, a + off, 1 << 22, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hb, b + off, 1 << 22, cudaMemcpyDeviceToHost));
      for (int i = 0; i < (1 << 22); i++)
        if (ha[i] != hb[i]) {
          printf("  MISMATCH (%s) at byte %lld\n", what, off + i);
          return;
        }
    }
  };
  printf("rows  seg    pitch | ldg one-shot |  tma_ld 2 stages | tma_ld 3 stages | tma_ld 4 stages | tma_ld_st 3 stages   (GB/s, read+write)\n");
  for (int rows : {128, 256})
    for (int seg : {128, 256})
      for (long long pitch : {32768LL, 262144LL, 2097152LL}) {
        const int tpr = (int)(pitch / seg);
        const long long total_rows = bytes / pitch;
        const long long nblk = total_rows / rows;
        const long long ntiles = nblk * tpr;
        CUtensorMap im = make_map(enc, a, pitch, total_rows, seg, rows);
        CUtensorMap om = make_map(enc, b, pitch, total_rows, seg, rows);
        float ms = 0;
        for (int it = 0; it < 3; it++) {
          CK(cudaEventRecord(e0));
          ldg_tile_copy<8><<<(unsigned)ntiles, 256>>>(a, b, rows, seg, pitch, tpr);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double gb = 2.0 * bytes / 1e9;
        printf("%4d %4d %8lld | %12.0f |", rows, seg, pitch, gb / (ms * 1e-3));
        CK(cudaMemset(b, 0, bytes));
        ms = run_tma<2, false>(im, om, b, rows, seg, pitch, tpr, ntiles, sms, e0, e1);
        printf(" %15.0f |", gb / (ms * 1e-3));
        verify("tma_ld x2");
        ms = run_tma<3, false>(im, om, b, rows, seg, pitch, tpr, ntiles, sms, e0, e1);
        printf(" %15.0f |", gb / (ms * 1e-3));
        ms = run_tma<4, false>(im, om, b, rows, seg, pitch, tpr, ntiles, sms, e0, e1);
        printf(" %15.0f |", gb / (ms * 1e-3));
        CK(cudaMemset(b, 0, bytes));
        ms = run_tma<3, true>(im, om, b, rows, seg, pitch, tpr, ntiles, sms, e0, e1);
        printf(" %15.0f\n", gb / (ms * 1e-3));
        verify("tma_ld_st");
        fflush(stdout);
      }
  CK(cudaDeviceSynchronize());
  return 0;
}
