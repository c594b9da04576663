// Scratch microbenchmark: distributed-shared-memory (DSMEM) exchange bandwidth of a 2-CTA cluster on sm_100a.
// Question it answers ("one-pass 32768-point rows"): a 32768-point fp32 row is 256 KiB, more than one CTA's
// shared memory, so a one-pass kernel would split it over a CTA pair and move HALF of the row across the SM-to-SM
// network at each of the three exchanges between its four radix stages.  Is that network fast enough to stay under
// the HBM time of the row (2 x 256 KiB at ~6.5 TB/s / 74 pairs)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dsmem_bench.cu -o /tmp/dsmem_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

constexpr int kThreads = 1024;
constexpr int kElems = 16384;  // float2 per CTA: half of a 32768-point row (128 KiB)

// mode 0: every thread reads its 8 remote elements (half of its 16 points) per exchange from the partner CTA
// mode 1: ... writes them to the partner CTA;  mode 2: the same traffic but local (baseline)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) exchange_kernel(float2* out, int reps) {
  extern __shared__ __align__(16) float2 sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  float2* remote = cluster.map_shared_rank(sm, rank ^ 1u);
  float2* target = MODE == 2 ? sm : remote;
  const int t = threadIdx.x;
  for (int i = t; i < kElems; i += kThreads) sm[i] = make_float2((float)i, (float)rank);
  cluster.sync();
  float2 acc = make_float2(0.f, 0.f);
  for (int r = 0; r < reps; r++) {
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) target[t + i * kThreads + ((r & 1) ? 8192 : 0)] = make_float2(acc.x + i, acc.y + r);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float2 v = target[t + i * kThreads + ((r & 1) ? 8192 : 0)];
        acc.x += v.x;
        acc.y += v.y;
      }
    }
    cluster.sync();  // an exchange of a Stockham stage ends in a barrier across the pair
  }
  if (acc.x == 123.456f) out[blockIdx.x * kThreads + t] = acc;
  if (MODE == 1 && sm[t].x == 123.456f) out[t] = sm[t];
}

template <int MODE>
static void run(const char* name, int sms, float2* out) {
  const size_t smem = sizeof(float2) * kElems;
  cudaFuncSetAttribute(exchange_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 2000;
  const int grid = sms / 2 * 2;
  exchange_kernel<MODE><<<grid, kThreads, smem>>>(out, 10);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  exchange_kernel<MODE><<<grid, kThreads, smem>>>(out, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  const double bytes_per_cta = 8.0 * kThreads * 8 * reps;  // 8 float2 per thread per exchange
  const double us_per_exchange = ms * 1e3 / reps;
  printf("%-28s %s  %.3f us per exchange of 64 KiB per CTA (incl. cluster.sync)  %.1f GB/s per SM  %.2f TB/s chip\n", name,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e), us_per_exchange, bytes_per_cta / (ms * 1e-3) / 1e9,
         bytes_per_cta * grid / (ms * 1e-3) / 1e12);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float2* out;
  cudaMalloc(&out, sizeof(float2) * kThreads * sms);
  run<2>("local smem reads (baseline)", sms, out);
  run<0>("remote (DSMEM) reads", sms, out);
  run<1>("remote (DSMEM) writes", sms, out);
  // the budget: one 32768-point row per CTA pair = 512 KiB of HBM traffic; at 6.5 TB/s over 74 pairs that is
  const double row_us = 512.0 * 1024 / (6.5e12 / (sms / 2)) * 1e6;
  printf("HBM time of one 32768-point fp32 row per CTA pair at 6.5 TB/s: %.2f us; a one-pass kernel needs 3 exchanges per row\n", row_us);
  return 0;
}
