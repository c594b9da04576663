// Scratch microbenchmark: achievable HBM bandwidth of the tile access pattern of a Stockham pass.
// A CTA copies a tile of R rows x SEG bytes (rows `stride` bytes apart) from in to out; tiles are adjacent
// segments, then the next row-block.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stride_bench.cu -o tools/stride_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC>
__global__ void tile_copy(const char* __restrict__ in, char* __restrict__ out, long long total_bytes, int rows,
                          int seg, long long stride, long long ntiles, int tiles_per_row, int lanes_per_row) {
  // threads: lane-in-row (seg/16 lanes of 16 B) x rows-per-iteration
  const int lr = threadIdx.x % lanes_per_row;
  const int r0 = threadIdx.x / lanes_per_row;
  const int rstep = blockDim.x / lanes_per_row;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long blk = tile / tiles_per_row, t = tile % tiles_per_row;
    const long long base = blk * (long long)rows * stride + t * seg + lr * 16;
    int4 v[VEC];
    for (int r = r0; r < rows; r += rstep * VEC) {
#pragma unroll
      for (int k = 0; k < VEC; k++)
        if (r + k * rstep < rows) v[k] = *reinterpret_cast<const int4*>(in + base + (long long)(r + k * rstep) * stride);
#pragma unroll
      for (int k = 0; k < VEC; k++)
        if (r + k * rstep < rows) *reinterpret_cast<int4*>(out + base + (long long)(r + k * rstep) * stride) = v[k];
    }
  }
}

int main() {
  const long long bytes = 1LL << 30;
  char *a, *b;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes);
  cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("rows seg stride inplace ctas/sm  GB/s\n");
  for (int inplace = 0; inplace < 2; inplace++)
    for (int rows : {128, 256, 1024})
      for (int seg : {32, 64, 128, 256, 512})
        for (long long stride : {2048LL, 131072LL, 1048576LL, 4194304LL}) {
          if ((long long)rows * stride > bytes) continue;
          if (stride < seg) continue;
          const int tiles_per_row = (int)(stride / seg);
          const long long nblk = bytes / ((long long)rows * stride);
          const long long ntiles = nblk * tiles_per_row;
          const int lanes = seg / 16;
          const int threads = 256;
          const int ctas = 4;
          for (int it = 0; it < 3; it++) {
            cudaEventRecord(e0);
            tile_copy<8><<<sms * ctas, threads>>>(a, inplace ? a : b, bytes, rows, seg, stride, ntiles, tiles_per_row, lanes);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
          }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          printf("%4d %4d %8lld %d %d  %7.0f\n", rows, seg, stride, inplace, ctas, 2.0 * bytes / (ms * 1e-3) / 1e9);
        }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
