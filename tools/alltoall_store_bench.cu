// Scratch microbenchmark (N GPUs, one process): the ceiling of an all-to-all done with SM-issued peer stores -- every GPU
// streams an equal share of a buffer to every other GPU at the same time, 8-byte stores in 256-byte runs (the shape of
// the distributed FFT's last pass) or 16-byte stores.  The distributed 2D transform's all-to-all phases are judged
// against this figure (and against the nominal 900 GB/s per direction).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/alltoall_store_bench.cu -o /tmp/a2a_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Peers { void* p[8]; };

// `vecs_per_peer` vectors go to each of the n-1 peers; consecutive warp-instructions of a warp rotate over the peers
template <typename V>
__global__ void __launch_bounds__(256) a2a_kernel(Peers dst, int self, int n, size_t vecs_per_peer) {
  const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x / 32);
  const int lane = threadIdx.x & 31;
  V v;
  float* f = reinterpret_cast<float*>(&v);
  for (int k = 0; k < (int)(sizeof(V) / 4); k++) f[k] = (float)(warp + k);
  for (size_t base = warp * 32; base < vecs_per_peer; base += nwarps * 32) {
#pragma unroll 1
    for (int g = 1; g < n; g++) {
      const int peer = (self + g) % n;
      reinterpret_cast<V*>(dst.p[peer])[(size_t)self * vecs_per_peer + base + lane] = v;
    }
  }
}

template <typename V>
static void run(const char* name, int n, std::vector<void*>& buf, std::vector<cudaStream_t>& st, size_t bytes_per_peer) {
  Peers pp;
  for (int g = 0; g < 8; g++) pp.p[g] = g < n ? buf[g] : nullptr;
  const size_t vecs = bytes_per_peer / sizeof(V);
  std::vector<cudaEvent_t> a(n), b(n);
  for (int rep = 0; rep < 3; rep++) {
    for (int d = 0; d < n; d++) {
      CK(cudaSetDevice(d));
      if (rep == 2) { CK(cudaEventCreate(&a[d])); CK(cudaEventCreate(&b[d])); CK(cudaEventRecord(a[d], st[d])); }
      a2a_kernel<V><<<148 * 8, 256, 0, st[d]>>>(pp, d, n, vecs);
      if (rep == 2) CK(cudaEventRecord(b[d], st[d]));
    }
    for (int d = 0; d < n; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
  }
  float worst = 0;
  for (int d = 0; d < n; d++) {
    float ms;
    CK(cudaSetDevice(d));
    CK(cudaEventElapsedTime(&ms, a[d], b[d]));
    worst = ms > worst ? ms : worst;
  }
  const double sent = (double)bytes_per_peer * (n - 1);
  printf("  %-34s %d GPUs: %.1f MB sent per GPU in %.3f ms (slowest GPU) = %.0f GB/s per direction per GPU\n", name, n,
         sent / 1e6, worst, sent / (worst * 1e-3) / 1e9);
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  for (int n : {2, 4, 8}) {
    if (n > ndev) break;
    std::vector<void*> buf(n);
    std::vector<cudaStream_t> st(n);
    const size_t per_peer = (size_t)128 << 20;  // 128 MiB to every peer: the size of C5's blocks on 8 GPUs
    for (int d = 0; d < n; d++) {
      CK(cudaSetDevice(d));
      for (int e = 0; e < n; e++)
        if (e != d) { cudaError_t r = cudaDeviceEnablePeerAccess(e, 0); if (r != cudaSuccess && r != cudaErrorPeerAccessAlreadyEnabled) CK(r); else cudaGetLastError(); }
      CK(cudaMalloc(&buf[d], per_peer * n));
      CK(cudaStreamCreate(&st[d]));
    }
    run<float2>("8-byte stores, 256-byte runs", n, buf, st, per_peer);
    run<float4>("16-byte stores, 512-byte runs", n, buf, st, per_peer);
    for (int d = 0; d < n; d++) { CK(cudaSetDevice(d)); CK(cudaFree(buf[d])); CK(cudaStreamDestroy(st[d])); }
  }
  return 0;
}
