// Register-resident radix-2/4/8/16 forward DFT butterflies (sign = -1).
//
// These replace the radix-2 butterfly sweeps of the reference
// (include/genFFT/generic/fft_impl_generic.h:41-60 and the AVX/SSE versions in
// include/genFFT/x86/fft_float_impl_x86.inl:42-64, fft_double_impl_x86.inl:43-65): the same DFT is
// computed with radix-16 register kernels, so a 4096-point transform needs 3 passes over the data
// instead of 12.  Inverse transforms are obtained by conjugating on load and on store, so only the
// forward butterflies exist.
#pragma once
#include <cuda_runtime.h>

namespace genfft_cuda {

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

template <typename T>
struct cpx {
  T x, y;
  __device__ __forceinline__ cpx() {}
  __device__ __forceinline__ cpx(T re, T im) : x(re), y(im) {}
};

template <typename T> __device__ __forceinline__ cpx<T> operator+(cpx<T> a, cpx<T> b) { return cpx<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ cpx<T> operator-(cpx<T> a, cpx<T> b) { return cpx<T>(a.x - b.x, a.y - b.y); }
// (a.x + i a.y) * (b.x + i b.y)
template <typename T> __device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return cpx<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T> __device__ __forceinline__ cpx<T> cmul_conj(cpx<T> a, cpx<T> b) {
  return cpx<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> __device__ __forceinline__ cpx<T> conj(cpx<T> a) { return cpx<T>(a.x, -a.y); }
// multiply by -i  (forward quarter turn)
template <typename T> __device__ __forceinline__ cpx<T> mul_mi(cpx<T> a) { return cpx<T>(a.y, -a.x); }
// multiply by +i
template <typename T> __device__ __forceinline__ cpx<T> mul_pi(cpx<T> a) { return cpx<T>(-a.y, a.x); }

// elementwise operations on the pair (x, y); one FMUL2 / FFMA2 each for T = float (overloads below)
template <typename T> __device__ __forceinline__ cpx<T> pair_mul(cpx<T> a, cpx<T> b) { return cpx<T>(a.x * b.x, a.y * b.y); }
template <typename T> __device__ __forceinline__ cpx<T> pair_fma(cpx<T> a, cpx<T> b, cpx<T> c) {
  return cpx<T>(a.x * b.x + c.x, a.y * b.y + c.y);
}

// ---- packed single precision ------------------------------------------------------------------------------------------
// sm_100 has two-wide FP32 instructions on register pairs (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).
// An interleaved complex value IS such a pair, so a complex add is one instruction instead of two, and more than half
// of a radix-16 butterfly's arithmetic is adds (the radix-4 steps are adds only): 864 -> 744 instructions in the
// twiddled 256-point column pass, 864 -> 688 in the 4096-point kernel, bit-identical results.  Measured on a B200
// (profiles/r02_ab_packed_f32.log): C5 10.03 -> 9.47 ms, C2 6.19 -> 6.28 TB/s, 2^21 x 256 3.85 -> 3.74 ms.  The
// two-instruction packed complex MULTIPLY (pair shuffles, spills under the 64-register cap) measured slower and is not
// here.  Non-template overloads: they win overload resolution against the generic templates above for T = float; the
// kernel-logic emulator of the CPU tests keeps the scalar forms.
#if !defined(GENFFT_EMU)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ cpx<float> unpack2(unsigned long long r) {
  cpx<float> c;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
  return c;
}
__device__ __forceinline__ cpx<float> operator+(cpx<float> a, cpx<float> b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
  return unpack2(r);
}
__device__ __forceinline__ cpx<float> operator-(cpx<float> a, cpx<float> b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
  return unpack2(r);
}
__device__ __forceinline__ cpx<float> pair_mul(cpx<float> a, cpx<float> b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
  return unpack2(r);
}
__device__ __forceinline__ cpx<float> pair_fma(cpx<float> a, cpx<float> b, cpx<float> c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)), "l"(pack2(c.x, c.y)));
  return unpack2(r);
}
#endif

template <typename T> struct consts {
  static constexpr T sqrt1_2 = T(0.70710678118654752440084436210484903928L);
  static constexpr T cos_pi_8 = T(0.92387953251128675612818318939678828682L);
  static constexpr T sin_pi_8 = T(0.38268343236508977172845998403039886676L);
};

// multiply by W_8^1 = (1 - i)/sqrt2 and W_8^3 = (-1 - i)/sqrt2
template <typename T> __device__ __forceinline__ cpx<T> mul_w8_1(cpx<T> a) {
  return cpx<T>((a.x + a.y) * consts<T>::sqrt1_2, (a.y - a.x) * consts<T>::sqrt1_2);
}
template <typename T> __device__ __forceinline__ cpx<T> mul_w8_3(cpx<T> a) {
  return cpx<T>((a.y - a.x) * consts<T>::sqrt1_2, -(a.x + a.y) * consts<T>::sqrt1_2);
}

// In-place DFTs on a strided set of registers x[0], x[S], x[2S], ...
// After the call, the register at slot s holds output bin out_bin<R>(s).

template <typename T, int S>
__device__ __forceinline__ void dft2(cpx<T>* x) {
  cpx<T> a = x[0], b = x[S];
  x[0] = a + b;
  x[S] = a - b;
}

// natural-order radix-4: slot s holds bin s
template <typename T, int S>
__device__ __forceinline__ void dft4(cpx<T>* x) {
  cpx<T> a0 = x[0] + x[2 * S];
  cpx<T> a1 = x[0] - x[2 * S];
  cpx<T> a2 = x[S] + x[3 * S];
  cpx<T> a3 = mul_mi(x[S] - x[3 * S]);
  x[0] = a0 + a2;
  x[S] = a1 + a3;
  x[2 * S] = a0 - a2;
  x[3 * S] = a1 - a3;
}

// Radix-R DFT over x[0..R) (unit register stride).  R = R1*R2 (Cooley-Tukey in registers):
//   step 1: R2 DFTs of length R1 over n1 (input index n = R2*n1 + n2)
//   step 2: twiddle by W_R^(n2*k1)
//   step 3: R1 DFTs of length R2 over n2, output bin k = k1 + R1*k2 lands in slot R2*k1 + k2.
template <int R> struct RegDFT;

template <> struct RegDFT<1> {
  template <typename T> static __device__ __forceinline__ void run(cpx<T>*) {}
  static __host__ __device__ constexpr int out_bin(int s) { return s; }
};

template <> struct RegDFT<2> {
  template <typename T> static __device__ __forceinline__ void run(cpx<T>* x) { dft2<T, 1>(x); }
  static __host__ __device__ constexpr int out_bin(int s) { return s; }
};

template <> struct RegDFT<4> {
  template <typename T> static __device__ __forceinline__ void run(cpx<T>* x) { dft4<T, 1>(x); }
  static __host__ __device__ constexpr int out_bin(int s) { return s; }
};

template <> struct RegDFT<8> {
  // R1 = 2, R2 = 4
  template <typename T> static __device__ __forceinline__ void run(cpx<T>* x) {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft2<T, 4>(x + n2);  // slots 4*k1 + n2
    x[5] = mul_w8_1(x[5]);
    x[6] = mul_mi(x[6]);
    x[7] = mul_w8_3(x[7]);
    dft4<T, 1>(x);
    dft4<T, 1>(x + 4);
  }
  // slot 4*k1 + k2 -> bin k1 + 2*k2
  static __host__ __device__ constexpr int out_bin(int s) { return (s >> 2) + 2 * (s & 3); }
};

template <> struct RegDFT<16> {
  // R1 = 4, R2 = 4
  template <typename T> static __device__ __forceinline__ void run(cpx<T>* x) {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4<T, 4>(x + n2);  // slots 4*k1 + n2
    const T c1 = consts<T>::cos_pi_8, s1 = consts<T>::sin_pi_8;
    // k1 = 1: W16^1, W16^2, W16^3
    x[5] = cmul(x[5], cpx<T>(c1, -s1));
    x[6] = mul_w8_1(x[6]);
    x[7] = cmul(x[7], cpx<T>(s1, -c1));
    // k1 = 2: W16^2, W16^4, W16^6
    x[9] = mul_w8_1(x[9]);
    x[10] = mul_mi(x[10]);
    x[11] = mul_w8_3(x[11]);
    // k1 = 3: W16^3, W16^6, W16^9
    x[13] = cmul(x[13], cpx<T>(s1, -c1));
    x[14] = mul_w8_3(x[14]);
    x[15] = cmul(x[15], cpx<T>(-c1, s1));
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4<T, 1>(x + 4 * k1);
  }
  // slot 4*k1 + k2 -> bin k1 + 4*k2
  static __host__ __device__ constexpr int out_bin(int s) { return (s >> 2) + 4 * (s & 3); }
};

}  // namespace genfft_cuda
