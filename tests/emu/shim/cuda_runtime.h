// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//
// A fake <cuda_runtime.h> that lets g++ compile genfft_b200/csrc/*.cu unchanged into tests/emu/_build/libgenfft_emu.so,
// so that the "-m 'not gpu'" tests can run the REAL kernel sources (tile addressing, Stockham stages, fused twiddles,
// fused real-FFT split, pass chains and their ticket order, plan construction) on a machine without a GPU and compare
// them with the oracle.  Device memory is host memory, a kernel launch runs every CTA in turn with its threads as
// fibers (emu_runtime.cpp), __syncthreads() switches fibers.  Nothing in genfft_b200 ever loads this library: the
// product path has no CPU fallback (tests/test_abi.py::test_product_never_loads_the_emulator).
//
// What the emulator can and cannot show: it checks the kernels' LOGIC (indices, tables, barriers that are missing
// between a write and a read by a lower-numbered thread, plan/pass bookkeeping).  It says nothing about performance,
// PTX spelling, memory-model races between CTAs, or anything in the `#ifndef GENFFT_EMU` blocks (inline PTX), which
// only the -m gpu tests exercise.
#pragma once
#define GENFFT_EMU 1

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>

// ---- language extensions ------------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct alignas(8) float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace genfft_emu {
// set by the scheduler before a fiber resumes
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_body);
unsigned char* dyn_smem();
void barrier();
void spin_yield();
}  // namespace genfft_emu
#define threadIdx (::genfft_emu::g_threadIdx)
#define blockIdx (::genfft_emu::g_blockIdx)
#define blockDim (::genfft_emu::g_blockDim)
#define gridDim (::genfft_emu::g_gridDim)

// ---- device intrinsics the kernels use -------------------------------------------------------------------------------
inline void __syncthreads() { ::genfft_emu::barrier(); }
template <typename V> inline V __ldg(const V* p) { return *p; }
template <typename V> inline V __ldcs(const V* p) { return *p; }
template <typename V> inline V __ldcg(const V* p) { return *p; }
template <typename V> inline void __stcs(V* p, const V& v) { *p = v; }
inline uint32_t __brev(uint32_t v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
inline void __nanosleep(unsigned) {}
inline void __trap() { abort(); }
inline void __threadfence_system() {}
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
  const uint32_t old = *p;
  *p = old + v;
  return old;
}
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// ---- runtime API (the subset planner.cu / pass_chain.cu / abi.cu / host_exec.cu call) ---------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
inline const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory (emulated)" : "invalid value (emulated)";
}
struct CUstream_st;
struct CUevent_st;
typedef CUstream_st* cudaStream_t;
typedef CUevent_st* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace genfft_emu {
int num_sms();  // GENFFT_EMU_SMS (default 3): small, so that persistent grids have several CTAs but stay cheap
}

// Device allocations sit between two inaccessible guard pages with their END on the upper guard (16-byte granular),
// so a kernel that runs past (or before) an allocation it was given faults immediately -- a poor man's memcheck.
namespace genfft_emu {
void* guarded_alloc(size_t bytes);
void guarded_free(void* p);
}
inline cudaError_t cudaMalloc(void** p, size_t bytes) {
  *p = ::genfft_emu::guarded_alloc(bytes);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t cudaFree(void* p) { ::genfft_emu::guarded_free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrComputeCapabilityMajor ? 10 : ::genfft_emu::num_sms();
  return cudaSuccess;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
// "IPC": every emulated rank lives in this process, so a handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
enum { cudaHostRegisterDefault = 0 };
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetPCIBusId(char* s, int n, int) { snprintf(s, n, "0000:00:00.0"); return cudaSuccess; }
