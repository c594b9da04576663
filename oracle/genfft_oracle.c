/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Plain-C CPU restatement of the algorithm of mzient/genFFT (the reference) for the
 * transform hot path: bit-reversal scramble + in-place radix-2 decimation-in-time levels with
 * per-level fp64-computed twiddle tables, the real-FFT split ("DIT"), the vertical (column) FFT,
 * the 2D transform and the real-image 2D transforms (RealFFT2D forward / forward_2x).  Every
 * function cites the reference file:line it follows (see genfft_oracle_impl.inc).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (1) bit for bit against the reference's own generic scalar back-end compiled from
 *       /root/reference (oracle/_ref/libgenfft_ref.so, genfft_ref_generic_*), and against the
 *       committed outputs of that build under tests/golden/ (made by tests/golden/make_golden.py),
 *   (2) within the reference's own test tolerance FFT_Eps (test/test_util.h:62-72) against the
 *       reference's in-test comparand reference_impl::FFT_pow2 and, for n <= 512, the naive DFT
 *       (test/fft_ref_impl.h:90-135, test/test_reference.cpp:31-65).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (genfft_b200/, libgenfft_cuda.so) never does and has no CPU path.
 */
#define _USE_MATH_DEFINES
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define T float
#define FN(name) oracle_##name##_f32
#include "genfft_oracle_impl.inc"
#undef T
#undef FN

#define T double
#define FN(name) oracle_##name##_f64
#include "genfft_oracle_impl.inc"
#undef T
#undef FN
