// GPU test of the C++ class mirror (include/genfft_cuda/fft.h) through HOST pointers -- what a CPU caller of
// genFFT would write.  The loops follow the structure of the reference's own tests:
//   TestFFT_Pow2      test/fft_test_impl.h:35-58   (forward, inverse, round trip)
//   TestDIT_Pow2      test/fft_test_impl.h:60-82   (DIT of the n/2 FFT == n FFT of the real signal)
//   TestRealFFT_Pow2  test/fft_test_impl.h:84-106  (half/full, no write past n/2+1)
//   TestFFTVert_Pow2  test/fft_test_impl.h:108-131 (vertical, odd column counts)
// with data from a default-seeded std::mt19937_64 U(-1,1) (test/test_util.h:36-57) and the reference's
// absolute tolerance FFT_Eps (test/test_util.h:62-72).  The comparand is a plain double-precision
// recursive FFT written here (the product never links the oracle).
#include <genfft_cuda/fft.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

static int g_fail = 0;
#define CHECK(cond, ...)                            \
  do {                                              \
    if (!(cond)) {                                  \
      if (g_fail < 20) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
      g_fail++;                                     \
    }                                               \
  } while (0)

typedef std::complex<double> cd;

static void fft_rec(std::vector<cd>& a, bool inv) {
  const size_t n = a.size();
  if (n == 1) return;
  std::vector<cd> e(n / 2), o(n / 2);
  for (size_t i = 0; i < n / 2; i++) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
  fft_rec(e, inv);
  fft_rec(o, inv);
  const double s = inv ? 1.0 : -1.0;
  for (size_t k = 0; k < n / 2; k++) {
    const double ang = s * 2.0 * M_PI * (double)k / (double)n;
    const cd t = cd(std::cos(ang), std::sin(ang)) * o[k];
    a[k] = e[k] + t;
    a[k + n / 2] = e[k] - t;
  }
}

template <class T> double fft_eps(int n);
template <> double fft_eps<float>(int n) { return 1e-5 + n * 1e-8; }
template <> double fft_eps<double>(int n) { return 1e-8 + n * 1e-12; }

template <class T>
void dummy(std::vector<std::complex<T>>& v) {
  std::mt19937_64 rng;
  std::uniform_real_distribution<T> dist(-1, 1);
  for (auto& c : v) { c.real(dist(rng)); c.imag(dist(rng)); }
}
template <class T>
void dummy(std::vector<T>& v) {
  std::mt19937_64 rng;
  std::uniform_real_distribution<T> dist(-1, 1);
  for (auto& x : v) x = dist(rng);
}

template <class T>
void test_fft_pow2(int n) {
  genfft::FFT<T> fft(n);
  CHECK(fft.size() == n && (bool)fft, "size/bool");
  std::vector<std::complex<T>> in(n), out(n), invout(n);
  dummy(in);
  fft.template transform<false>(out.data(), in.data());
  fft.template transform<true>(invout.data(), out.data());
  std::vector<cd> ref(in.begin(), in.end());
  fft_rec(ref, false);
  const double eps = fft_eps<T>(n);
  for (int i = 0; i < n; i++) {
    CHECK(std::abs(cd(out[i]) - ref[i]) <= eps * 1.5, "fwd n=%d i=%d", n, i);
    CHECK(std::abs(cd(invout[i]) / (double)n - cd(in[i])) <= eps, "roundtrip n=%d i=%d", n, i);
  }
  // README spelling
  std::vector<std::complex<T>> out2(n);
  fft.forward(out2.data(), in.data());
  for (int i = 0; i < n; i++) CHECK(out2[i] == out[i], "forward() != transform<false>()");
}

template <class T>
void test_real_fft_pow2(int n, bool half) {
  genfft::RealFFT<T> fft(n);
  const std::complex<T> fill(43, 21);
  std::vector<std::complex<T>> out(n, fill);
  std::vector<T> in(n);
  dummy(in);
  fft.forward(out.data(), in.data(), half);
  std::vector<cd> ref(n);
  for (int i = 0; i < n; i++) ref[i] = in[i];
  fft_rec(ref, false);
  const double eps = fft_eps<T>(n);
  const int limit = n == 1 ? 1 : (half ? n / 2 + 1 : n);
  for (int i = 0; i < limit; i++) CHECK(std::abs(cd(out[i]) - ref[i]) <= eps * 1.5, "r2c n=%d half=%d i=%d", n, half, i);
  for (int i = limit; i < n; i++) CHECK(out[i] == fill, "Corruption detected @ index %d (n=%d)", i, n);
}

template <class T>
void test_dit_pow2(int n, bool in_place) {
  std::vector<T> real_input(n);
  dummy(real_input);
  std::vector<std::complex<T>> out1(n), out3(n);
  genfft::FFT<T> half_fft(n == 1 ? 1 : n / 2);
  half_fft.template transform<false>(out1.data(), (std::complex<T>*)real_input.data());
  genfft::DIT<T> dit(n);
  std::complex<T>* out_ptr = in_place ? out1.data() : out3.data();
  dit.apply(out_ptr, out1.data(), false);
  std::vector<cd> ref(n);
  for (int i = 0; i < n; i++) ref[i] = real_input[i];
  fft_rec(ref, false);
  const double eps = fft_eps<T>(n);
  for (int i = 0; i < n; i++) CHECK(std::abs(cd(out_ptr[i]) - ref[i]) <= eps * 1.5, "dit n=%d i=%d", n, i);
}

template <class T>
void test_vert_pow2(int n, int cols) {
  genfft::FFTVert<T> fft(n);
  std::vector<std::complex<T>> in((size_t)n * cols), out((size_t)n * cols), invout((size_t)n * cols);
  dummy(in);
  fft.template transform<false>(out.data(), cols, in.data(), cols, cols);
  fft.template transform<true>(invout.data(), cols, out.data(), cols, cols);
  const double eps = fft_eps<T>(n);
  for (int c = 0; c < cols; c += (cols > 8 ? cols / 5 : 1)) {
    std::vector<cd> ref(n);
    for (int r = 0; r < n; r++) ref[r] = in[(size_t)r * cols + c];
    fft_rec(ref, false);
    for (int r = 0; r < n; r++) {
      CHECK(std::abs(cd(out[(size_t)r * cols + c]) - ref[r]) <= eps * 1.5, "vert n=%d cols=%d r=%d c=%d", n, cols, r, c);
      CHECK(std::abs(cd(invout[(size_t)r * cols + c]) / (double)n - cd(in[(size_t)r * cols + c])) <= eps, "vert roundtrip");
    }
  }
}

template <class T>
void test_fft2d(int w, int h) {
  genfft::FFT2D<T> fft(w, h);
  CHECK(fft.cols() == w && fft.rows() == h, "cols/rows");
  std::vector<std::complex<T>> in((size_t)w * h), out((size_t)w * h), back((size_t)w * h);
  dummy(in);
  fft.template transform<false>(out.data(), w, in.data(), w);
  fft.template transform<true>(back.data(), w, out.data(), w);
  // separable reference
  std::vector<cd> ref(in.begin(), in.end());
  for (int r = 0; r < h; r++) {
    std::vector<cd> row(ref.begin() + (size_t)r * w, ref.begin() + (size_t)(r + 1) * w);
    fft_rec(row, false);
    std::copy(row.begin(), row.end(), ref.begin() + (size_t)r * w);
  }
  for (int c = 0; c < w; c++) {
    std::vector<cd> col(h);
    for (int r = 0; r < h; r++) col[r] = ref[(size_t)r * w + c];
    fft_rec(col, false);
    for (int r = 0; r < h; r++) ref[(size_t)r * w + c] = col[r];
  }
  const double eps = fft_eps<T>(w * h);
  for (size_t i = 0; i < in.size(); i++) {
    CHECK(std::abs(cd(out[i]) - ref[i]) <= eps * 1.5, "2d %dx%d i=%zu", w, h, i);
    CHECK(std::abs(cd(back[i]) / (double)(w * h) - cd(in[i])) <= eps, "2d roundtrip");
  }
}

template <class T>
void run_all() {
  for (int n = 1; n <= (1 << 18); n += n) test_fft_pow2<T>(n);
  for (int n = 1; n <= (1 << 18); n += n)
    for (int half = 0; half < 2; half++) test_real_fft_pow2<T>(n, half != 0);
  for (int n = 2; n <= (1 << 16); n += n)
    for (int ip = 0; ip < 2; ip++) test_dit_pow2<T>(n, ip != 0);
  const int vert[][2] = {{1, 1}, {2, 3}, {4, 7}, {8, 33}, {16, 47}, {32, 63}, {128, 767}, {256, 999}, {512, 1023}, {4096, 31}};
  for (auto& v : vert) test_vert_pow2<T>(v[0], v[1]);
  test_fft2d<T>(8, 4);
  test_fft2d<T>(64, 128);
  test_fft2d<T>(1024, 32);
  // two-for-one real helper (FFTReal.h:35-66) on top of a complex transform
  const int n = 256;
  std::vector<T> a(n), b(n);
  dummy(a);
  for (int i = 0; i < n; i++) b[i] = a[(i * 7 + 3) % n];
  std::vector<std::complex<T>> z(n), fz(n), fa(n), fb(n);
  genfft::FFT<T> fft(n);
  fft.transform_interleave(fz.data(), a.data(), b.data());  // fft.h:100-105
  for (int i = 0; i < n; i++) z[i] = std::complex<T>(a[i], b[i]);
  std::vector<std::complex<T>> fz2(n);
  fft.template transform<false>(fz2.data(), z.data());
  for (int i = 0; i < n; i++) CHECK(std::abs(cd(fz[i]) - cd(fz2[i])) <= fft_eps<T>(n), "transform_interleave");
  genfft::separate_2x_real_FFT(fa.data(), fb.data(), fz.data(), n);
  std::vector<cd> ra(a.begin(), a.end()), rb(b.begin(), b.end());
  fft_rec(ra, false);
  fft_rec(rb, false);
  for (int i = 0; i < n; i++) {
    CHECK(std::abs(cd(fa[i]) - ra[i]) <= fft_eps<T>(n) * 2, "separate a");
    CHECK(std::abs(cd(fb[i]) - rb[i]) <= fft_eps<T>(n) * 2, "separate b");
  }
}

template <class T>
void test_real_fft2d(int w, int h) {
  genfft::RealFFT2D<T> fft(w, h);
  std::vector<T> in((size_t)w * h);
  dummy(in);
  std::vector<std::complex<T>> out((size_t)w * h);
  fft.forward(out.data(), w, in.data(), w);
  std::vector<cd> ref(in.begin(), in.end());
  for (int r = 0; r < h; r++) {
    std::vector<cd> row(ref.begin() + (size_t)r * w, ref.begin() + (size_t)(r + 1) * w);
    fft_rec(row, false);
    std::copy(row.begin(), row.end(), ref.begin() + (size_t)r * w);
  }
  for (int c = 0; c < w; c++) {
    std::vector<cd> col(h);
    for (int r = 0; r < h; r++) col[r] = ref[(size_t)r * w + c];
    fft_rec(col, false);
    for (int r = 0; r < h; r++) ref[(size_t)r * w + c] = col[r];
  }
  const double eps = fft_eps<T>(w * h);
  for (size_t i = 0; i < in.size(); i++) CHECK(std::abs(cd(out[i]) - ref[i]) <= eps * 1.5, "real 2d %dx%d i=%zu", w, h, i);
}

// forward_2x (FFTReal.h:106-118): spectrum of in1 + i*in2, with different strides of the two images
template <class T>
void test_real_fft2d_2x(int w, int h) {
  genfft::RealFFT2D<T> fft(w, h);
  const int s1 = w + 2, s2 = w + 4;
  std::vector<T> in1((size_t)s1 * h), in2((size_t)s2 * h);
  dummy(in1);
  dummy(in2);
  std::vector<std::complex<T>> out((size_t)w * h);
  fft.forward_2x(out.data(), w, in1.data(), s1, in2.data(), s2);
  std::vector<cd> ref((size_t)w * h);
  for (int r = 0; r < h; r++) {
    std::vector<cd> row(w);
    for (int c = 0; c < w; c++) row[c] = cd(in1[(size_t)r * s1 + c], in2[(size_t)r * s2 + c]);
    fft_rec(row, false);
    std::copy(row.begin(), row.end(), ref.begin() + (size_t)r * w);
  }
  for (int c = 0; c < w; c++) {
    std::vector<cd> col(h);
    for (int r = 0; r < h; r++) col[r] = ref[(size_t)r * w + c];
    fft_rec(col, false);
    for (int r = 0; r < h; r++) ref[(size_t)r * w + c] = col[r];
  }
  const double eps = fft_eps<T>(w * h);
  for (size_t i = 0; i < out.size(); i++) CHECK(std::abs(cd(out[i]) - ref[i]) <= eps * 2, "real 2d 2x %dx%d i=%zu", w, h, i);
}

int main() {
  test_real_fft2d<float>(64, 32);
  test_real_fft2d_2x<float>(64, 32);
  test_real_fft2d_2x<double>(16, 128);
  test_real_fft2d<double>(16, 128);
  genfft::FFT<float> empty;
  CHECK(!(bool)empty && empty.size() == 0, "default-constructed plan must be empty");
  bool threw = false;
  try { genfft::FFT<float> bad(12); } catch (const genfft::cuda_error& e) { threw = e.code == GENFFT_CUDA_ERR_SIZE; }
  CHECK(threw, "non power of two must be rejected");
  run_all<float>();
  run_all<double>();
  std::printf("%s (%d failures)\n", g_fail ? "FAILED" : "PASSED", g_fail);
  return g_fail ? 1 : 0;
}
