// Scratch: what the L2 of this part delivers to the SMs when the data is L2-RESIDENT -- the ceiling of a pass of the
// L2-resident chains (chain_kernel.cuh), whose intermediate never goes to HBM.  Every CTA owns one contiguous chunk
// and streams it `reps` times inside one launch; loads go past L1 (ld.global.cg), so every byte crosses the
// SM <-> L2 fabric.  Working sets from 4 MiB to 64 MiB per buffer (L2: 126 MB) and 1 GiB (HBM) for comparison.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/l2_bw_bench.cu -o tools/l2_bw_bench
#include <cstdio>
#include <cuda_runtime.h>

// volatile asm: the loads of every repetition must really be issued (the compiler would hoist a plain load of the
// loop-invariant address out of the repetition loop)
__device__ __forceinline__ float2 ld_cg(const float2* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_cg(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <typename V, int VEC>
__global__ void copy_rep(const V* __restrict__ in, V* __restrict__ out, int reps) {
  const long long base = (long long)blockIdx.x * blockDim.x * VEC + threadIdx.x;
  for (int r = 0; r < reps; r++) {
    V v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) v[k] = ld_cg(in + base + k * blockDim.x);
#pragma unroll
    for (int k = 0; k < VEC; k++) out[base + k * blockDim.x] = v[k];
  }
}
template <typename V, int VEC>
__global__ void write_rep(V* __restrict__ out, int reps) {
  const long long base = (long long)blockIdx.x * blockDim.x * VEC + threadIdx.x;
  for (int r = 0; r < reps; r++) {
    V v;
#pragma unroll
    for (int j = 0; j < (int)(sizeof(V) / sizeof(float)); j++) reinterpret_cast<float*>(&v)[j] = (float)(r + j);
#pragma unroll
    for (int k = 0; k < VEC; k++) out[base + k * blockDim.x] = v;
  }
}

int main() {
  const size_t maxb = 1ull << 30;
  void *a, *b;
  cudaMalloc(&a, maxb); cudaMalloc(&b, maxb); cudaMemset(a, 1, maxb); cudaMemset(b, 2, maxb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time_ms = [&](auto launch) {
    float best = 1e9;
    for (int it = 0; it < 4; it++) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
  };
  printf("%-10s %-6s %12s %12s   (GB/s; copy counts read + write)\n", "buffer", "access", "copy", "write");
  const size_t sizes[] = {4ull << 20, 8ull << 20, 16ull << 20, 32ull << 20, 48ull << 20, 64ull << 20, 1ull << 30};
  for (size_t bytes : sizes) {
    const int reps = (int)((bytes >= (1ull << 30)) ? 1 : (4ull << 30) / bytes);  // ~4 GiB through the fabric per launch
    {
      using V = float2; constexpr int VEC = 16;  // 8-byte accesses, 16 per thread: what the pass kernels issue
      const unsigned grid = (unsigned)(bytes / sizeof(V) / (256 * VEC));
      const float c = time_ms([&] { copy_rep<V, VEC><<<grid, 256>>>((const V*)a, (V*)b, reps); });
      const float w = time_ms([&] { write_rep<V, VEC><<<grid, 256>>>((V*)b, reps); });
      const double tot = (double)bytes * reps;
      printf("%6zu MiB %-6s %12.0f %12.0f\n", bytes >> 20, "8 B", 2 * tot / c / 1e6, tot / w / 1e6);
    }
    {
      using V = float4; constexpr int VEC = 8;  // 16-byte accesses
      const unsigned grid = (unsigned)(bytes / sizeof(V) / (256 * VEC));
      const float c = time_ms([&] { copy_rep<V, VEC><<<grid, 256>>>((const V*)a, (V*)b, reps); });
      const float w = time_ms([&] { write_rep<V, VEC><<<grid, 256>>>((V*)b, reps); });
      const double tot = (double)bytes * reps;
      printf("%6zu MiB %-6s %12.0f %12.0f\n", bytes >> 20, "16 B", 2 * tot / c / 1e6, tot / w / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
