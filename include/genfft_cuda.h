/* libgenfft_cuda -- C ABI of the B200 (sm_100a) FFT engine that drops in for genFFT's transform path.
 *
 * This boundary replaces the reference's run-time back-end selection: the six link-time symbols
 *   genfft::impl_x86_dispatch::GetImpl / GetVertImpl / GetDITImpl (int n, float|double)
 * declared in include/genFFT/x86/fft_x86_dispatch.h:33-40 and defined in src/fft_x86_dispatch.cpp:64-139,
 * and the objects they return (impl::FFTBase<T>, impl::FFTVertBase<T>, impl::FFTDITBase<T>,
 * include/genFFT/FFTLevel.h:43-59,102-118, include/genFFT/FFTDIT.h:43-49).  The C++ headers
 * include/genfft_cuda/fft.h (class mirror of genfft::FFT / FFTVert / DIT / FFT2D / RealFFT) and
 * include/genfft_cuda/backend.h (factories with the reference's FFTImplFactory signature,
 * include/genFFT/fft.h:41-52) sit on top of these entry points.
 *
 * Conventions (identical to the reference):
 *   - interleaved complex (re, im), layout-compatible with std::complex<T>;
 *   - forward X[k] = sum x[n] exp(-2*pi*i*n*k/N); inverse uses +, and is NOT scaled (README.txt:31);
 *   - power-of-two sizes only; strides / distances are in complex elements unless stated.
 * Differences: sizes up to 2^27 (the reference stops at 2^23), batched and device-pointer entry
 * points, and errors are returned as status codes instead of assert()
 * (include/genFFT/x86/fft_float_impl_x86.inl:494-495).
 *
 * There is no CPU fallback: every entry point fails with GENFFT_CUDA_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef GENFFT_CUDA_H
#define GENFFT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct genfft_cuda_plan_s* genfft_cuda_plan_t;

enum {
  GENFFT_CUDA_OK = 0,
  GENFFT_CUDA_ERR_SIZE = 1,    /* not a power of two / out of range (reference: assert(!"unsupported size")) */
  GENFFT_CUDA_ERR_ARG = 2,     /* null pointer, out == in where the reference forbids it, bad enum */
  GENFFT_CUDA_ERR_CUDA = 3,    /* CUDA runtime error; see genfft_cuda_last_error_string() */
  GENFFT_CUDA_ERR_ALLOC = 4
};

enum { GENFFT_CUDA_F32 = 0, GENFFT_CUDA_F64 = 1 };

/* thread-local description of the last failure */
const char* genfft_cuda_last_error_string(void);
/* number of usable sm_100 devices (0 if none) */
int genfft_cuda_device_count(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
uint64_t genfft_cuda_launch_count(void);

/* ---- plans --------------------------------------------------------------------------------------
 * A plan is immutable after creation and may be executed concurrently from several host threads on
 * different streams as long as the executions do not share the plan's internal scratch (plans that
 * need scratch serialise on the stream they are given; the counters of the L2-resident pass chains
 * are kept per stream).  Plans are created on the current CUDA device and must be executed with that
 * device current (GENFFT_CUDA_ERR_ARG otherwise).  The HOST-pointer entry points share one set of
 * staging buffers per plan and therefore serialise per plan (an internal mutex is held for the whole
 * call): concurrent host-pointer calls on one plan are safe but do not overlap -- use one plan per
 * thread for that.  Batched plans reject distances smaller than one transform (GENFFT_CUDA_ERR_ARG).
 * Replaces FFT<T>::FFT(int n) -> factory(n, T()) (include/genFFT/fft.h:59-64). */

/* 1D complex, n points, `batch` transforms; in_dist/out_dist between consecutive transforms
 * (0 = n).  FFT<T> (fft.h:54-113). */
int genfft_cuda_plan_c2c_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch,
                            int64_t in_dist, int64_t out_dist);
/* 1D real input, n real points -> n/2+1 (half) or n complex bins.  RealFFT<T> (FFTReal.h:186-221).
 * in_dist in real scalars (0 = n), out_dist in complex elements (0 = n/2+1 if half else n). */
int genfft_cuda_plan_r2c_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch, int half,
                            int64_t in_dist, int64_t out_dist);
/* Half-spectrum INVERSE of the real transform: n/2+1 bins -> n real points, unscaled (out = n * x).  An addition:
 * the reference has no inverse real FFT (README.txt:51-52).  in_dist in complex elements (0 = n/2+1), out_dist in
 * real scalars (0 = n). */
int genfft_cuda_plan_c2r_1d(genfft_cuda_plan_t* plan, int precision, int64_t n, int64_t batch, int64_t in_dist,
                            int64_t out_dist);
/* 2D complex width x height (row-major, width contiguous).  FFT2D<T>(width, height) (fft.h:198-245). */
int genfft_cuda_plan_c2c_2d(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t height);
/* 2D real input width x height.  RealFFT2D<T>(width, height) (FFTReal.h:71-184); forward only, as in the reference
 * (its inverse is assert(!"TODO"), FFTReal.h:127-130). */
int genfft_cuda_plan_r2c_2d(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t height);
/* n-point FFT along axis 0 of an (n x cols) row-major array.  FFTVert<T> (fft.h:115-171). */
int genfft_cuda_plan_vert(genfft_cuda_plan_t* plan, int precision, int64_t n);
/* real-FFT split / post-process of size n.  DIT<T> (fft.h:173-196), adjust_DIT_impl
 * (include/genFFT/generic/fft_dit_impl_generic.inl:27-61). */
int genfft_cuda_plan_dit(genfft_cuda_plan_t* plan, int precision, int64_t n);
int genfft_cuda_plan_destroy(genfft_cuda_plan_t plan);
/* Launch only a fraction of the resident-CTA capacity (frac_other: every pass but the last; frac_last: the last
 * pass) so that two plans on different streams share the SMs -- used to overlap an NVLink-bound remote-store pass
 * with the HBM-bound local pass of the next chunk in the distributed 2D transform. */
int genfft_cuda_plan_set_grid_fraction(genfft_cuda_plan_t plan, double frac_other, double frac_last);

/* introspection (used by the benchmark to compute roofline figures) */
int64_t genfft_cuda_plan_size(genfft_cuda_plan_t plan);          /* FFT<T>::size(), fft.h:107 */
int genfft_cuda_plan_num_passes(genfft_cuda_plan_t plan);        /* kernel launches per execution */
size_t genfft_cuda_plan_scratch_bytes(genfft_cuda_plan_t plan);
/* writes a human readable description of the pass decomposition into buf */
int genfft_cuda_plan_describe(genfft_cuda_plan_t plan, char* buf, size_t buflen);

/* ---- execution on DEVICE pointers (the measured path) ---------------------------------------------
 * `stream` is a cudaStream_t (NULL = legacy default stream).  All calls are asynchronous. */

/* FFT<T>::transform<inv>(out, in) (fft.h:80-85).  out == in is accepted (goes through scratch when
 * the decomposition is not in-place safe). */
int genfft_cuda_exec_c2c_dev(genfft_cuda_plan_t plan, void* out, const void* in, int inverse, void* stream);
/* FFT<T>::transform_no_scramble<inv>(inout) (fft.h:69-73): input is in bit-reversed order, in place. */
int genfft_cuda_exec_c2c_no_scramble_dev(genfft_cuda_plan_t plan, void* inout, int inverse, void* stream);
/* FFT<T>::transform_real(out, in) (fft.h:90-94): n real scalars -> n complex bins. */
int genfft_cuda_exec_c2c_real_in_dev(genfft_cuda_plan_t plan, void* out, const void* in_real, void* stream);
/* FFT<T>::transform_interleave(out, in1, in2) (fft.h:100-105): transform of in1 + i*in2 (two real signals). */
int genfft_cuda_exec_c2c_interleave_dev(genfft_cuda_plan_t plan, void* out, const void* in1, const void* in2,
                                        void* stream);
/* separate_2x_real_FFT(out1, out2, in, N) (FFTReal.h:35-66); out1 or out2 may alias in. */
int genfft_cuda_separate_2x_real_dev(int precision, void* out1, void* out2, const void* in, int64_t n, void* stream);
/* RealFFT2D<T>::forward(out, out_stride, in, in_stride) (FFTReal.h:83-104): real width x height image -> full
 * width x height complex spectrum; out_stride in complex elements, in_stride in real scalars; out != in. */
int genfft_cuda_exec_r2c_2d_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                                int64_t in_stride, void* stream);
/* RealFFT2D<T>::forward_2x(out, out_stride, in1, in_stride1, in2, in_stride2) (FFTReal.h:106-118): full
 * width x height spectrum of the complex image in1 + i*in2 (two real images in one transform); strides of the inputs
 * in real scalars.  The reference offsets in2's lower half with in_stride1 (FFTReal.h:178), a typo: in_stride2 is
 * honoured here. */
int genfft_cuda_exec_r2c_2d_2x_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in1,
                                   int64_t in_stride1, const void* in2, int64_t in_stride2, void* stream);
/* inverse of RealFFT<T>::forward(half = true), unscaled; out != in; n >= 2. */
int genfft_cuda_exec_c2r_dev(genfft_cuda_plan_t plan, void* out, const void* in, void* stream);
/* RealFFT<T>::forward(out, in, half) (FFTReal.h:204-213); half is fixed at plan time.  `out` is also the
 * workspace and must hold n/2+1 (half) or n complex elements per transform. */
int genfft_cuda_exec_r2c_dev(genfft_cuda_plan_t plan, void* out, const void* in, void* stream);
/* FFT2D<T>::transform<inv>(out, out_stride, in, in_stride) (fft.h:213-218); out != in. */
int genfft_cuda_exec_c2c_2d_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                                int64_t in_stride, int inverse, void* stream);
/* FFTVert<T>::transform<inv>(out, out_stride, in, in_stride, cols) (fft.h:145-150); out != in. */
int genfft_cuda_exec_vert_dev(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                              int64_t in_stride, int64_t cols, int inverse, void* stream);
/* FFTVert<T>::transform_no_scramble<inv>(data, stride, cols) (fft.h:132-136): rows bit-reversed, in place. */
int genfft_cuda_exec_vert_no_scramble_dev(genfft_cuda_plan_t plan, void* data, int64_t stride, int64_t cols,
                                          int inverse, void* stream);
/* DIT<T>::apply(out, in, half) (fft.h:181-189); out may equal in. */
int genfft_cuda_exec_dit_dev(genfft_cuda_plan_t plan, void* out, const void* in, int half, void* stream);

/* ---- execution on HOST pointers (the literal drop-in for the reference's CPU callers) -----------------
 * The library stages host<->device copies itself (chunked and overlapped with compute for batched
 * plans) and returns when `out` is complete. */
int genfft_cuda_exec_c2c(genfft_cuda_plan_t plan, void* out, const void* in, int inverse);
int genfft_cuda_exec_c2c_no_scramble(genfft_cuda_plan_t plan, void* inout, int inverse);
int genfft_cuda_exec_c2c_real_in(genfft_cuda_plan_t plan, void* out, const void* in_real);
int genfft_cuda_exec_r2c(genfft_cuda_plan_t plan, void* out, const void* in);
int genfft_cuda_exec_c2r(genfft_cuda_plan_t plan, void* out, const void* in);
int genfft_cuda_exec_c2c_interleave(genfft_cuda_plan_t plan, void* out, const void* in1, const void* in2);
int genfft_cuda_exec_r2c_2d(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in, int64_t in_stride);
int genfft_cuda_exec_r2c_2d_2x(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in1,
                               int64_t in_stride1, const void* in2, int64_t in_stride2);
int genfft_cuda_exec_c2c_2d(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                            int64_t in_stride, int inverse);
int genfft_cuda_exec_vert(genfft_cuda_plan_t plan, void* out, int64_t out_stride, const void* in,
                          int64_t in_stride, int64_t cols, int inverse);
int genfft_cuda_exec_vert_no_scramble(genfft_cuda_plan_t plan, void* data, int64_t stride, int64_t cols,
                                      int inverse);
int genfft_cuda_exec_dit(genfft_cuda_plan_t plan, void* out, const void* in, int half);
/* separate_2x_real_FFT(out1, out2, in, N) (FFTReal.h:35-66) on host pointers; out1 or out2 may alias in. */
int genfft_cuda_separate_2x_real(int precision, void* out1, void* out2, const void* in, int64_t n);

/* Page-locked host buffers for the entry points above, placed on the NUMA node of the current device when
 * numa_local != 0 (mmap + mbind + cudaHostRegister; *numa_node_out = the node, or -1 when it could not be bound).
 * The host-pointer path is PCIe-bound, so on a two-socket multi-GPU host the placement of the caller's buffers
 * decides its rate.  Nothing in the reference corresponds to this (genFFT never leaves the CPU). */
int genfft_cuda_host_alloc(void** ptr, size_t bytes, int numa_local, int* numa_node_out);
int genfft_cuda_host_free(void* ptr);

/* ---- distributed 2D building blocks (slab decomposition; one process per GPU) ---------------------------
 * The host side (genfft_b200/dist.py) owns the process group; these run the local passes.
 *
 * rows_pass: `rows` independent length-`width` row FFTs of a (rows x width) slab.  The output is
 * written block-transposed for the all-to-all: bin k of local row r goes to
 *     dst[k / (width/nparts)] + (row0 + r) * (width/nparts) + k % (width/nparts)
 * where dst[g] = out_peers[g] (direct NVLink stores into rank g's receive buffer, fused transpose) or,
 * when out_peers == NULL, out + g * part_stride (packed for ncclSend/Recv).
 * cols_pass: length-`height` FFTs down the columns of a (height x cols) array with row pitch `stride`,
 * optionally scattering rows [g*height/nparts, ...) to peer g at row pitch out_stride, column offset
 * col0 (the transpose back to row slabs). */
int genfft_cuda_plan_dist_rows(genfft_cuda_plan_t* plan, int precision, int64_t width, int64_t rows,
                               int nparts);
int genfft_cuda_exec_dist_rows_dev(genfft_cuda_plan_t plan, void* out, void* const* out_peers,
                                   int64_t part_stride, int64_t row0, const void* in, int64_t in_stride,
                                   int inverse, void* stream);
int genfft_cuda_plan_dist_cols(genfft_cuda_plan_t* plan, int precision, int64_t height, int64_t cols,
                               int nparts);
int genfft_cuda_exec_dist_cols_dev(genfft_cuda_plan_t plan, void* out, void* const* out_peers,
                                   int64_t out_stride, int64_t col0, void* data, int64_t stride,
                                   int inverse, void* stream);

/* strided block copy out[b*out_dist + r*out_stride + c] = in[b*in_dist + r*in_stride + c] (complex
 * elements); used to unpack ncclRecv buffers into row slabs on the NCCL path. */
int genfft_cuda_copy2d_dev(int precision, void* out, int64_t out_stride, int64_t out_dist, const void* in,
                           int64_t in_stride, int64_t in_dist, int64_t rows, int64_t cols, int64_t batch,
                           void* stream);

/* ---- distributed four-step 1D building blocks (genfft_b200/dist.py::DistFFT1D) ---------------------------------
 * A transform of N = H*W points, seen as an H x W row-major matrix whose row slabs live on different GPUs, is
 * column transforms (dist_cols above), the twiddle W_N^(kr*c), and row transforms (dist_rows above); nothing in the
 * reference corresponds to it (genFFT is single-threaded, one address space).
 * twiddle2d: data[r*stride + c] *= W_N^((row0 + r) * c) in place (conjugated for the inverse), N = n_total.
 * transpose: out[c*out_stride + r] = in[r*in_stride + c] (natural-order output of the four-step transform). */
/* dist_cols with the four-step twiddle W_n^(kr*c) fused into the peer stores (*fused = 1), or not (*fused = 0: the
 * caller runs twiddle2d on the receiving slab).  scatter_cols: the first global transpose (row slab -> the peers' column
 * blocks at rows [row0, row0 + rows)) in one launch. */
int genfft_cuda_exec_dist_cols_tw_dev(genfft_cuda_plan_t plan, void* const* out_peers, int64_t out_stride, int64_t col0,
                                      void* data, int64_t stride, int inverse, int64_t n_total, int* fused, void* stream);
int genfft_cuda_scatter_cols_dev(int precision, void* const* peers, int nparts, int64_t row0, const void* in,
                                 int64_t in_stride, int64_t rows, int64_t width, void* stream);
int genfft_cuda_twiddle2d_dev(int precision, void* data, int64_t stride, int64_t rows, int64_t cols, int64_t row0,
                              int64_t n_total, int inverse, void* stream);
int genfft_cuda_transpose_dev(int precision, void* out, int64_t out_stride, const void* in, int64_t in_stride,
                              int64_t rows, int64_t cols, void* stream);

/* Test hook: launches so far of the kernels compiled for addressing mode `mode` (tile_kernel.cuh enum Mode; a chained
 * launch counts for both of its passes) -- lets a test assert that a specialised mode, not the generic one, ran. */
uint64_t genfft_cuda_debug_mode_launch_count(int mode);

/* Test hook: x / d as the kernels compute it when they decode a tile index (magic-number multiply, exact for
 * x < 2^31); host code only. */
uint32_t genfft_cuda_debug_fast_div(uint32_t x, uint32_t d);

/* Measurement hook (bench.py, C1): `iters` times forward in -> mid then inverse mid -> out on `stream`, issued from a C
 * loop and followed by one stream synchronisation; *us_per_pair = host wall-clock per forward+inverse pair, i.e. what
 * a C/C++ caller of the device-pointer path pays per pair of small transforms (launch-rate bound). */
int genfft_cuda_debug_time_c2c_pairs(genfft_cuda_plan_t plan, void* out, void* mid, const void* in, int iters,
                                     void* stream, double* us_per_pair);

/* Stream-ordered barrier across the ranks of one node over IPC-mapped flag arrays: peer_flags[r] is rank r's array
 * of `world` uint32 epochs (zero-initialised with genfft_cuda_memset_dev, mapped with the IPC helpers below); `epoch`
 * must increase by one per barrier.  Replaces a collective-library call between the passes of the distributed 2D
 * transform (nothing in the reference corresponds to it: genFFT has no multi-device code). */
int genfft_cuda_peer_barrier_dev(void* const* peer_flags, int rank, int world, uint32_t epoch, void* stream);
int genfft_cuda_memset_dev(void* ptr, int value, size_t bytes);

/* CUDA IPC helpers so that ranks can map each other's receive buffers (cudaIpcMemHandle_t is 64 bytes). */
int genfft_cuda_malloc(void** ptr, size_t bytes);
int genfft_cuda_free(void* ptr);
int genfft_cuda_ipc_get_handle(void* ptr, unsigned char handle[64]);
int genfft_cuda_ipc_open_handle(void** ptr, const unsigned char handle[64]);
int genfft_cuda_ipc_close_handle(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* GENFFT_CUDA_H */
