// Scratch: what limits a streaming copy on this part?  Variants of a 2 GiB -> 2 GiB copy:
//   persistent grid-stride vs one-shot grid, bytes in flight per thread, cache hints.
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC, int HINT>
__global__ void copy_k(const int4* __restrict__ in, int4* __restrict__ out, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * VEC) {
    int4 v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) if (i + k * stride < n) v[k] = HINT ? __ldcs(in + i + k * stride) : in[i + k * stride];
#pragma unroll
    for (int k = 0; k < VEC; k++) if (i + k * stride < n) { if (HINT) __stcs(out + i + k * stride, v[k]); else out[i + k * stride] = v[k]; }
  }
}
// block-contiguous: each CTA copies a contiguous chunk of `chunk` int4 per iteration
template <int VEC, int HINT>
__global__ void copy_blk(const int4* __restrict__ in, int4* __restrict__ out, long long n) {
  const long long chunk = (long long)blockDim.x * VEC;
  for (long long base = (long long)blockIdx.x * chunk; base < n; base += (long long)gridDim.x * chunk) {
    int4 v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) v[k] = HINT ? __ldcs(in + base + threadIdx.x + k * blockDim.x) : in[base + threadIdx.x + k * blockDim.x];
#pragma unroll
    for (int k = 0; k < VEC; k++) { if (HINT) __stcs(out + base + threadIdx.x + k * blockDim.x, v[k]); else out[base + threadIdx.x + k * blockDim.x] = v[k]; }
  }
}
int main() {
  const long long bytes = 2LL << 30; const long long n = bytes / 16;
  int4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMemset(a, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto run = [&](const char* name, auto launch) {
    float best = 1e9;
    for (int it = 0; it < 5; it++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-48s %7.0f GB/s\n", name, 2.0 * bytes / (best * 1e-3) / 1e9);
  };
  run("cudaMemcpy D2D", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
  run("grid-stride persistent 4x256/SM VEC4", [&] { copy_k<4, 0><<<sms * 4, 256>>>(a, b, n); });
  run("grid-stride persistent 8x256/SM VEC4", [&] { copy_k<4, 0><<<sms * 8, 256>>>(a, b, n); });
  run("grid-stride persistent 8x256/SM VEC8", [&] { copy_k<8, 0><<<sms * 8, 256>>>(a, b, n); });
  run("grid-stride persistent 8x256/SM VEC4 ldcs/stcs", [&] { copy_k<4, 1><<<sms * 8, 256>>>(a, b, n); });
  run("one-shot grid VEC4 (n/1024 CTAs)", [&] { copy_k<4, 0><<<(unsigned)(n / 1024), 256>>>(a, b, n); });
  run("block-contig persistent 4x256/SM VEC8 (32 KiB)", [&] { copy_blk<8, 0><<<sms * 4, 256>>>(a, b, n); });
  run("block-contig persistent 4x256/SM VEC8 ldcs/stcs", [&] { copy_blk<8, 1><<<sms * 4, 256>>>(a, b, n); });
  run("block-contig one-shot VEC8 (n/2048 CTAs)", [&] { copy_blk<8, 0><<<(unsigned)(n / 2048), 256>>>(a, b, n); });
  run("block-contig one-shot VEC4 (n/1024 CTAs)", [&] { copy_blk<4, 0><<<(unsigned)(n / 1024), 256>>>(a, b, n); });
  run("block-contig one-shot VEC4 128thr", [&] { copy_blk<4, 0><<<(unsigned)(n / 512), 128>>>(a, b, n); });
  run("block-contig one-shot VEC8 stcs", [&] { copy_blk<8, 1><<<(unsigned)(n / 2048), 256>>>(a, b, n); });
  cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
