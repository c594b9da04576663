// libgenfft_cuda: host-pointer entry points -- the literal drop-in for the reference's CPU callers
// (FFT<T>::transform(std::complex<T>* out, const std::complex<T>* in), include/genFFT/fft.h:80-85, etc.).
// The library stages host<->device copies itself.  Batched 1D plans that need no scratch are cut into
// chunks that alternate between two streams, so chunk i's D2H, chunk i+1's kernels and chunk i+2's H2D
// overlap (PCIe is full duplex and the copy engines are independent of the SMs).
#include <cuda_runtime.h>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>

#include "../../include/genfft_cuda.h"
#include "plan.h"

using namespace genfft_cuda;


#define HX_TRY(expr)                                                        \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) {                                                \
      std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e);  \
      return set_error(GENFFT_CUDA_ERR_CUDA, m.c_str());                    \
    }                                                                       \
  } while (0)

static int ensure_stage(Plan* p, size_t in_bytes, size_t out_bytes) {
  std::lock_guard<std::mutex> lk(p->mu);
  if (p->stage_in_bytes < in_bytes) {
    if (p->stage_in) cudaFree(p->stage_in);
    p->stage_in = nullptr;
    p->stage_in_bytes = 0;
    if (cudaMalloc(&p->stage_in, in_bytes) != cudaSuccess) return set_error(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc of input staging failed");
    p->stage_in_bytes = in_bytes;
  }
  if (p->stage_out_bytes < out_bytes) {
    if (p->stage_out) cudaFree(p->stage_out);
    p->stage_out = nullptr;
    p->stage_out_bytes = 0;
    if (cudaMalloc(&p->stage_out, out_bytes) != cudaSuccess) return set_error(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc of output staging failed");
    p->stage_out_bytes = out_bytes;
  }
  if (!p->streams_ready) {
    for (auto& s : p->streams) HX_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    p->streams_ready = true;
  }
  return GENFFT_CUDA_OK;
}

static size_t chunk_bytes_target() {
  const char* s = getenv("GENFFT_CUDA_HOST_CHUNK_MB");
  size_t mb = s && *s ? (size_t)atol(s) : 64;
  return std::max<size_t>(mb, 1) << 20;
}

// batched 1D (c2c or r2c): in/out element sizes and distances differ between the two
static int exec_batched_host(Plan* p, void* out, const void* in, size_t in_elem, long long in_dist, long long in_len,
                             size_t out_elem, long long out_dist, long long out_len, bool r2c, int inverse,
                             bool brev, bool real_in) {
  const long long batch = p->batch;
  const size_t in_row = (size_t)in_dist * in_elem, out_row = (size_t)out_dist * out_elem;
  const size_t in_last = (size_t)in_len * in_elem, out_last = (size_t)out_len * out_elem;
  const bool pipelined = batch > 1 && !plan_needs_scratch(p) && !brev;
  long long chunk = batch;
  int nbuf = 1;
  if (pipelined) {
    chunk = std::max<long long>(1, (long long)(chunk_bytes_target() / std::max(in_row, out_row)));
    chunk = std::min(chunk, batch);
    if (chunk < batch) nbuf = 2;
  }
  const size_t in_chunk = (size_t)(chunk - 1) * in_row + in_last, out_chunk = (size_t)(chunk - 1) * out_row + out_last;
  // 256-byte aligned sub-buffers
  const size_t in_slot = (in_chunk + 255) & ~(size_t)255, out_slot = (out_chunk + 255) & ~(size_t)255;
  int rc = ensure_stage(p, in_slot * nbuf, out_slot * nbuf);
  if (rc) return rc;
  int slot = 0;
  for (long long b0 = 0; b0 < batch; b0 += chunk, slot ^= (nbuf - 1)) {
    const long long nb = std::min(chunk, batch - b0);
    cudaStream_t st = p->streams[slot];
    char* din = (char*)p->stage_in + slot * in_slot;
    char* dout = (char*)p->stage_out + slot * out_slot;
    const size_t ib = (size_t)(nb - 1) * in_row + in_last, ob = (size_t)(nb - 1) * out_row + out_last;
    // Rows are `dist` apart on both sides.  Without gaps (dist == length) a chunk is one linear copy; with gaps only
    // the transforms themselves move: the caller's memory between two outputs is not ours to write (and the staging
    // buffer's gaps hold nothing), the gaps of the input need not cross the bus.
    if (in_row == in_last || nb == 1)
      HX_TRY(cudaMemcpyAsync(din, (const char*)in + (size_t)b0 * in_row, ib, cudaMemcpyHostToDevice, st));
    else
      HX_TRY(cudaMemcpy2DAsync(din, in_row, (const char*)in + (size_t)b0 * in_row, in_row, in_last, (size_t)nb,
                               cudaMemcpyHostToDevice, st));
    rc = r2c ? exec_r2c_internal(p, dout, din, st, nb)
             : exec_c2c_internal(p, brev ? din : dout, din, inverse, st, brev, real_in, nb);
    if (rc) return rc;
    if (out_row == out_last || nb == 1)
      HX_TRY(cudaMemcpyAsync((char*)out + (size_t)b0 * out_row, brev ? din : dout, ob, cudaMemcpyDeviceToHost, st));
    else
      HX_TRY(cudaMemcpy2DAsync((char*)out + (size_t)b0 * out_row, out_row, brev ? din : dout, out_row, out_last,
                               (size_t)nb, cudaMemcpyDeviceToHost, st));
  }
  for (int s = 0; s < nbuf; s++) HX_TRY(cudaStreamSynchronize(p->streams[s]));
  return GENFFT_CUDA_OK;
}

// ---- page-locked host buffers on the NUMA node of the current device ------------------------------------------------
// The host-pointer path moves every byte across PCIe twice, so where the CALLER's buffers live decides its rate: on a
// two-socket multi-GPU host a buffer on the other socket also crosses the socket interconnect, and with one process
// per GPU all links are busy at once.  genfft_cuda_host_alloc maps anonymous memory, binds it to the GPU's node
// (mbind), touches it and registers it with CUDA -- no CPU of that node is needed in the process's cpuset.
namespace {
std::mutex g_host_mu;
std::map<void*, size_t> g_host_blocks;

int numa_node_of_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
  for (char* c = bus; *c; c++)
    if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
}  // namespace

extern "C" int genfft_cuda_host_alloc(void** ptr, size_t bytes, int numa_local, int* numa_node_out) {
  if (!ptr || !bytes) return set_error(GENFFT_CUDA_ERR_ARG, "null pointer or zero size");
  const size_t len = (bytes + 4095) & ~(size_t)4095;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return set_error(GENFFT_CUDA_ERR_ALLOC, "mmap of the host buffer failed");
  int node = numa_local ? numa_node_of_current_device() : -1;
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / 64] = 1ul << (node % 64);
    // MPOL_PREFERRED (1): falls back to another node instead of failing when the node is not in the cpuset's mems
    if (syscall(SYS_mbind, p, len, 1, mask, sizeof mask * 8 + 1, 0) != 0) node = -1;
  } else {
    node = -1;
  }
  for (size_t off = 0; off < len; off += 4096) static_cast<volatile char*>(p)[off] = 0;  // commit the pages there
  cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    munmap(p, len);
    return set_error(GENFFT_CUDA_ERR_CUDA, cudaGetErrorString(e));
  }
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[p] = len;
  }
  if (numa_node_out) *numa_node_out = node;
  *ptr = p;
  return GENFFT_CUDA_OK;
}

extern "C" int genfft_cuda_host_free(void* ptr) {
  if (!ptr) return GENFFT_CUDA_OK;
  size_t len = 0;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_blocks.find(ptr);
    if (it == g_host_blocks.end()) return set_error(GENFFT_CUDA_ERR_ARG, "not a genfft_cuda_host_alloc block");
    len = it->second;
    g_host_blocks.erase(it);
  }
  cudaHostUnregister(ptr);
  munmap(ptr, len);
  return GENFFT_CUDA_OK;
}

extern "C" {

int genfft_cuda_exec_c2c(genfft_cuda_plan_t plan, void* out, const void* in, int inverse) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return set_error(GENFFT_CUDA_ERR_ARG, "FFT::transform requires out != in");
  const size_t es = elem_size(p->precision);
  return exec_batched_host(p, out, in, es, p->in_dist, p->n, es, p->out_dist, p->n, false, inverse, false, false);
}

int genfft_cuda_exec_c2c_no_scramble(genfft_cuda_plan_t plan, void* inout, int inverse) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!inout) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (p->in_dist != p->out_dist) return set_error(GENFFT_CUDA_ERR_ARG, "in-place transform needs in_dist == out_dist");
  const size_t es = elem_size(p->precision);
  return exec_batched_host(p, inout, inout, es, p->in_dist, p->n, es, p->out_dist, p->n, false, inverse, true, false);
}

int genfft_cuda_exec_c2c_real_in(genfft_cuda_plan_t plan, void* out, const void* in_real) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in_real) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  const size_t es = elem_size(p->precision);
  return exec_batched_host(p, out, in_real, es / 2, p->in_dist, p->n, es, p->out_dist, p->n, false, 0, false, true);
}

// FFT<T>::transform_interleave(out, in1, in2) (fft.h:100-105): host pointers
int genfft_cuda_exec_c2c_interleave(genfft_cuda_plan_t plan, void* out, const void* in1, const void* in2) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2c_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in1 || !in2) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (p->batch != 1) return set_error(GENFFT_CUDA_ERR_ARG, "transform_interleave on host pointers supports batch 1");
  const size_t es = elem_size(p->precision);
  const size_t rbytes = (size_t)p->n * es / 2, cbytes = (size_t)p->n * es;
  int rc = ensure_stage(p, 2 * ((rbytes + 255) & ~(size_t)255), cbytes);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  char* d1 = (char*)p->stage_in;
  char* d2 = d1 + ((rbytes + 255) & ~(size_t)255);
  HX_TRY(cudaMemcpyAsync(d1, in1, rbytes, cudaMemcpyHostToDevice, st));
  HX_TRY(cudaMemcpyAsync(d2, in2, rbytes, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_c2c_interleave_dev(plan, p->stage_out, d1, d2, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpyAsync(out, p->stage_out, cbytes, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

// RealFFT2D<T>::forward(out, out_stride, in, in_stride) (FFTReal.h:83-104): host pointers
int genfft_cuda_exec_r2c_2d(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in, int64_t in_stride) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_R2C_2D) return set_error(GENFFT_CUDA_ERR_ARG, "not an r2c_2d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out_stride < p->width || in_stride < p->width) return set_error(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  const size_t es = elem_size(p->precision);
  const size_t in_bytes = (size_t)p->width * p->height * es / 2, out_bytes = (size_t)p->width * p->height * es;
  int rc = ensure_stage(p, in_bytes, out_bytes);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpy2DAsync(p->stage_in, p->width * es / 2, in, in_stride * es / 2, p->width * es / 2, p->height,
                           cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_r2c_2d_dev(plan, p->stage_out, p->width, p->stage_in, p->width, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpy2DAsync(out, out_stride * es, p->stage_out, p->width * es, p->width * es, p->height,
                           cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

// RealFFT2D<T>::forward_2x(out, out_stride, in1, in_stride1, in2, in_stride2) (FFTReal.h:106-118): host pointers.
// Both images are staged densely, so the device side always takes the path that reads them without interleaving.
int genfft_cuda_exec_r2c_2d_2x(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in1,
                               int64_t in_stride1, const void* in2, int64_t in_stride2) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_R2C_2D) return set_error(GENFFT_CUDA_ERR_ARG, "not an r2c_2d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in1 || !in2) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out_stride < p->width || in_stride1 < p->width || in_stride2 < p->width)
    return set_error(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  const size_t es = elem_size(p->precision);
  const size_t img_bytes = ((size_t)p->width * p->height * es / 2 + 255) & ~(size_t)255;
  const size_t out_bytes = (size_t)p->width * p->height * es;
  int rc = ensure_stage(p, 2 * img_bytes, out_bytes);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  char* d1 = (char*)p->stage_in;
  char* d2 = d1 + img_bytes;
  const size_t row = p->width * es / 2;
  HX_TRY(cudaMemcpy2DAsync(d1, row, in1, in_stride1 * es / 2, row, p->height, cudaMemcpyHostToDevice, st));
  HX_TRY(cudaMemcpy2DAsync(d2, row, in2, in_stride2 * es / 2, row, p->height, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_r2c_2d_2x_dev(plan, p->stage_out, p->width, d1, p->width, d2, p->width, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpy2DAsync(out, out_stride * es, p->stage_out, p->width * es, p->width * es, p->height,
                           cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

// half-spectrum inverse on host pointers
int genfft_cuda_exec_c2r(genfft_cuda_plan_t plan, void* out, const void* in) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2R_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2r_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  const size_t es = elem_size(p->precision);
  const size_t in_bytes = ((size_t)(p->batch - 1) * p->in_dist + p->n / 2 + 1) * es;
  const size_t out_row = (size_t)p->out_dist * es / 2, out_last = (size_t)p->n * es / 2;
  const size_t out_bytes = (size_t)(p->batch - 1) * out_row + out_last;
  int rc = ensure_stage(p, in_bytes, out_bytes);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpyAsync(p->stage_in, in, in_bytes, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_c2r_dev(plan, p->stage_out, p->stage_in, st);
  if (rc) return rc;
  // only the transforms themselves come back: the caller's memory between two outputs (out_dist > n) is not ours to write
  if (out_row == out_last || p->batch == 1)
    HX_TRY(cudaMemcpyAsync(out, p->stage_out, out_bytes, cudaMemcpyDeviceToHost, st));
  else
    HX_TRY(cudaMemcpy2DAsync(out, out_row, p->stage_out, out_row, out_last, (size_t)p->batch, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_r2c(genfft_cuda_plan_t plan, void* out, const void* in) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_R2C_1D) return set_error(GENFFT_CUDA_ERR_ARG, "not an r2c_1d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  const size_t es = elem_size(p->precision);
  const long long out_len = p->n == 1 ? 1 : (p->half ? p->n / 2 + 1 : p->n);
  return exec_batched_host(p, out, in, es / 2, p->in_dist, p->n, es, p->out_dist, out_len, true, 0, false, false);
}

int genfft_cuda_exec_c2c_2d(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in, int64_t in_stride,
                            int inverse) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_C2C_2D) return set_error(GENFFT_CUDA_ERR_ARG, "not a c2c_2d plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return set_error(GENFFT_CUDA_ERR_ARG, "FFT2D::transform requires out != in");
  if (out_stride < p->width || in_stride < p->width) return set_error(GENFFT_CUDA_ERR_ARG, "stride smaller than width");
  const size_t es = elem_size(p->precision);
  const size_t dense = (size_t)p->width * p->height * es;
  int rc = ensure_stage(p, dense, dense);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpy2DAsync(p->stage_in, p->width * es, in, in_stride * es, p->width * es, p->height, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_c2c_2d_dev(plan, p->stage_out, p->width, p->stage_in, p->width, inverse, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpy2DAsync(out, out_stride * es, p->stage_out, p->width * es, p->width * es, p->height, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_vert(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in, int64_t in_stride,
                          int64_t cols, int inverse) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_VERT) return set_error(GENFFT_CUDA_ERR_ARG, "not a vert plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (out == in) return set_error(GENFFT_CUDA_ERR_ARG, "FFTVert::transform requires out != in");
  if (cols < 0 || out_stride < cols || in_stride < cols) return set_error(GENFFT_CUDA_ERR_ARG, "bad cols/stride");
  if (cols == 0) return GENFFT_CUDA_OK;
  const size_t es = elem_size(p->precision);
  const size_t dense = (size_t)cols * p->n * es;
  int rc = ensure_stage(p, dense, dense);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpy2DAsync(p->stage_in, cols * es, in, in_stride * es, cols * es, p->n, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_vert_dev(plan, p->stage_out, cols, p->stage_in, cols, cols, inverse, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpy2DAsync(out, out_stride * es, p->stage_out, cols * es, cols * es, p->n, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_vert_no_scramble(genfft_cuda_plan_t plan, void* data, int64_t stride, int64_t cols, int inverse) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_VERT) return set_error(GENFFT_CUDA_ERR_ARG, "not a vert plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!data) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (cols < 0 || stride < cols) return set_error(GENFFT_CUDA_ERR_ARG, "bad cols/stride");
  if (cols == 0) return GENFFT_CUDA_OK;
  const size_t es = elem_size(p->precision);
  const size_t dense = (size_t)cols * p->n * es;
  int rc = ensure_stage(p, dense, 256);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpy2DAsync(p->stage_in, cols * es, data, stride * es, cols * es, p->n, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_vert_no_scramble_dev(plan, p->stage_in, cols, cols, inverse, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpy2DAsync(data, stride * es, p->stage_in, cols * es, cols * es, p->n, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

// separate_2x_real_FFT(out1, out2, in, N) (include/genFFT/FFTReal.h:35-66) on host pointers.  No plan exists for it
// (the reference's is a free function), so the staging is per call; out1 or out2 may alias in, as in the reference.
int genfft_cuda_separate_2x_real(int precision, void* out1, void* out2, const void* in, int64_t n) {
  if (precision != GENFFT_CUDA_F32 && precision != GENFFT_CUDA_F64) return set_error(GENFFT_CUDA_ERR_ARG, "bad precision");
  if (!out1 || !out2 || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  if (n < 1 || n > (1LL << 27)) return set_error(GENFFT_CUDA_ERR_SIZE, "unsupported size");
  const size_t bytes = (size_t)n * elem_size(precision);
  const size_t slot = (bytes + 255) & ~(size_t)255;
  char* d = nullptr;
  if (cudaMalloc(&d, 3 * slot) != cudaSuccess) return set_error(GENFFT_CUDA_ERR_ALLOC, "cudaMalloc of staging failed");
  int rc = GENFFT_CUDA_OK;
  cudaError_t e = cudaMemcpy(d, in, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = genfft_cuda_separate_2x_real_dev(precision, d + slot, d + 2 * slot, d, n, nullptr);
    if (!rc) e = cudaMemcpy(out1, d + slot, bytes, cudaMemcpyDeviceToHost);
    if (!rc && e == cudaSuccess) e = cudaMemcpy(out2, d + 2 * slot, bytes, cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  if (rc) return rc;
  if (e != cudaSuccess) return set_error(GENFFT_CUDA_ERR_CUDA, cudaGetErrorString(e));
  return GENFFT_CUDA_OK;
}

int genfft_cuda_exec_dit(genfft_cuda_plan_t plan, void* out, const void* in, int half) {
  Plan* p = plan;
  if (!p || p->kind != PLAN_DIT) return set_error(GENFFT_CUDA_ERR_ARG, "not a dit plan");
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  if (!out || !in) return set_error(GENFFT_CUDA_ERR_ARG, "null buffer");
  const size_t es = elem_size(p->precision);
  const long long n = p->n;
  const size_t in_bytes = (size_t)std::max<long long>(1, n / 2) * es;
  const size_t out_bytes = (size_t)(n <= 1 ? 1 : (half ? n / 2 + 1 : n)) * es;
  // the split runs in place on the device copy of the input (as RealFFT::forward does, FFTReal.h:211)
  int rc = ensure_stage(p, std::max(in_bytes, out_bytes), 256);
  if (rc) return rc;
  cudaStream_t st = p->streams[0];
  HX_TRY(cudaMemcpyAsync(p->stage_in, in, in_bytes, cudaMemcpyHostToDevice, st));
  rc = genfft_cuda_exec_dit_dev(plan, p->stage_in, p->stage_in, half, st);
  if (rc) return rc;
  HX_TRY(cudaMemcpyAsync(out, p->stage_in, out_bytes, cudaMemcpyDeviceToHost, st));
  HX_TRY(cudaStreamSynchronize(st));
  return GENFFT_CUDA_OK;
}

}  // extern "C"
