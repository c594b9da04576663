// Scratch microbenchmark (2 GPUs, one process): what shape of NVLink peer STORES reaches the link rate?
// The distributed transforms' last pass stores its tile straight into the peers' buffers: per warp instruction 32 lanes
// x 8 bytes = one 256-byte run, the runs of a thread 2-32 KiB apart.  Measured 640-700 GB/s per direction against the
// 770 GB/s of a peer copy.  This compares, with both GPUs storing to each other at the same time:
//   run8   : 8-byte stores, 256-byte runs scattered `gap` bytes apart      (what the FFT kernels do)
//   run16  : 16-byte stores, 512-byte runs scattered                       (what a staged, re-laid-out store would do)
//   seq8 / seq16 : the same store widths over one contiguous range         (a copy)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/peer_store_bench.cu -o /tmp/peer_store_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// every thread stores `per` vectors; a warp's lanes are contiguous; consecutive stores of a warp are `gap_vecs` apart
template <typename V>
__global__ void __launch_bounds__(256) store_kernel(V* dst, size_t total_vecs, size_t gap_vecs, int per, int scattered) {
  const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  V v;
  float* f = reinterpret_cast<float*>(&v);
  for (int k = 0; k < (int)(sizeof(V) / 4); k++) f[k] = (float)(warp + k);
  if (scattered) {
    // warp w owns runs  w*32 + j*gap  (mod total): a thread's stores are `gap` apart, like bins k = u + j*TN of a tile
    const size_t base = (warp * 32) % gap_vecs + (warp * 32 / gap_vecs) * gap_vecs * per;
#pragma unroll 8
    for (int j = 0; j < per; j++) {
      const size_t idx = base + (size_t)j * gap_vecs + lane;
      if (idx < total_vecs) dst[idx] = v;
    }
  } else {
    const size_t base = warp * 32 * per;
#pragma unroll 8
    for (int j = 0; j < per; j++) {
      const size_t idx = base + (size_t)j * 32 + lane;
      if (idx < total_vecs) dst[idx] = v;
    }
  }
}

template <typename V>
static double run(V* dst_on_1, V* dst_on_0, size_t bytes, size_t gap_bytes, int scattered, cudaStream_t s0, cudaStream_t s1) {
  const size_t total = bytes / sizeof(V), gap = gap_bytes / sizeof(V);
  const int per = 16;
  const size_t warps = total / (32 * per);
  const unsigned grid = (unsigned)(warps / 8);
  cudaEvent_t a, b;
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaSetDevice(0));
    if (rep == 2) { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventRecord(a, s0)); }
    store_kernel<V><<<grid, 256, 0, s0>>>(dst_on_1, total, gap, per, scattered);
    if (rep == 2) CK(cudaEventRecord(b, s0));
    CK(cudaSetDevice(1));
    store_kernel<V><<<grid, 256, 0, s1>>>(dst_on_0, total, gap, per, scattered);
    CK(cudaSetDevice(0));
    CK(cudaStreamSynchronize(s0));
    CK(cudaSetDevice(1));
    CK(cudaStreamSynchronize(s1));
  }
  CK(cudaSetDevice(0));
  CK(cudaEventElapsedTime(&ms, a, b));
  return bytes / (ms * 1e-3) / 1e9;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = (size_t)1 << 30;
  void *on0, *on1;
  cudaStream_t s0, s1;
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&on0, bytes)); CK(cudaStreamCreate(&s0));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&on1, bytes)); CK(cudaStreamCreate(&s1));
  printf("peer stores, both directions at once, 1 GiB per direction (GB/s per direction, timed on GPU 0):\n");
  for (size_t gap : {(size_t)2048, (size_t)32768, (size_t)1 << 20}) {
    printf("  run8   (8-byte stores, 256-byte runs %7zu B apart): %.0f\n", gap, run<float2>((float2*)on1, (float2*)on0, bytes, gap, 1, s0, s1));
    printf("  run16  (16-byte stores, 512-byte runs %7zu B apart): %.0f\n", gap, run<float4>((float4*)on1, (float4*)on0, bytes, gap, 1, s0, s1));
  }
  printf("  seq8   (8-byte stores, contiguous):  %.0f\n", run<float2>((float2*)on1, (float2*)on0, bytes, 0, 0, s0, s1));
  printf("  seq16  (16-byte stores, contiguous): %.0f\n", run<float4>((float4*)on1, (float4*)on0, bytes, 0, 0, s0, s1));
  // cudaMemcpyPeer for reference (one direction)
  CK(cudaSetDevice(0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaMemcpyPeerAsync(on1, 1, on0, 0, bytes, s0));
  CK(cudaEventRecord(a, s0));
  CK(cudaMemcpyPeerAsync(on1, 1, on0, 0, bytes, s0));
  CK(cudaEventRecord(b, s0));
  CK(cudaStreamSynchronize(s0));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("  cudaMemcpyPeer, one direction: %.0f\n", bytes / (ms * 1e-3) / 1e9);
  return 0;
}
