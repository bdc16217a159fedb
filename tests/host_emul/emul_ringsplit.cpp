// Host emulation of the radix-4 split chirp-z ring transform and of the whole-ring power-of-two transform of
// commander_b200/csrc/ringfft.cu (ring_split_kernel, ring_pow2_kernel): the same header code (ring_split.cuh,
// blue_fft.cuh) sequenced as the kernels sequence it, the CTA's threads run in a loop.  TEST ONLY, not a fallback.
#define BLUE_FFT_HOST
#include <cmath>
#include <vector>
#include "../../commander_b200/csrc/ring_split.cuh"
using namespace cmdr;

static double2 cmul(double2 a, double2 b) { return bf_mul(a, b); }

static std::vector<double2> twiddles2(int M) {
  std::vector<double2> tw(bf2_tw_total(M));
  for (int k = 0; k < bf2_num_strided(M); ++k) {
    const int h = bf2_half(M, k), st = h >> (bf2_stages(M, k) - 1);
    double2 *T = tw.data() + bf2_tw_offset(M, k);
    for (int j = 0; j < st; ++j) { T[j].x = std::cos(-M_PI * j / h); T[j].y = std::sin(-M_PI * j / h); }
  }
  tw[bf2_tw_total(M) - 1].x = 1.0; tw[bf2_tw_total(M) - 1].y = 0.0;
  return tw;
}
template <int DIR>
static void strided(double2 *x, int M, const double2 *T) {
  const int ns = bf2_num_strided(M);
  for (int kk = 0; kk < ns; ++kk) {
    const int k = DIR == 0 ? kk : ns - 1 - kk;
    const int S = bf2_stages(M, k), h = bf2_half(M, k);
    const double2 *Tk = T + bf2_tw_offset(M, k);
    for (int q = 0; q < (M >> S); ++q) {
      if (S == 3) { if (DIR == 0) dif_itemS<3, true>(x, h, Tk, q); else dit_itemS<3, true>(x, h, Tk, q); }
      else        { if (DIR == 0) dif_itemS<4, true>(x, h, Tk, q); else dit_itemS<4, true>(x, h, Tk, q); }
    }
  }
}
static void fft_dif(double2 *x, int M, const double2 *T) {
  strided<0>(x, M, T);
  for (int q = 0; q < (M >> 4); ++q) dif_itemS<4, true>(x, 8, T + bf2_tw_total(M) - 1, q);
}
static void fft_dit(double2 *x, int M, const double2 *T) {
  for (int q = 0; q < (M >> 4); ++q) dit_itemS<4, true>(x, 8, T + bf2_tw_total(M) - 1, q);
  strided<1>(x, M, T);
}
static void convolve(double2 *x, int M, const double2 *T, const double2 *v) {
  strided<0>(x, M, T);
  for (int q = 0; q < (M >> 4); ++q) conv_mid16<true>(x, v, q);
  strided<1>(x, M, T);
}
static std::vector<double2> sub_filter(int i, int M, const double2 *T) {     // ring_split_filter_kernel
  std::vector<double2> w(bf_padded(M));
  for (int k = 0; k < M; ++k) {
    const long long j = k < i ? k : (k > M - i ? M - k : -1);
    double2 val; val.x = val.y = 0.0;
    if (j >= 0) { const double2 e = rs_expipi(j * j, i); val.x = e.x; val.y = -e.y; }
    w[bf_pidx<true>(k)] = val;
  }
  fft_dif(w.data(), M, T);
  std::vector<double2> out(M);
  for (int k = 0; k < M; ++k) out[k] = w[bf_pidx<true>(k)];
  return out;
}

// dir 0: x_j = sum_k Z_k e^{+2 pi i j k / n} (synthesis, `in` = folded spectrum without shift, shifted = 0 here);
// dir 1: Z_k = sum_j z_j e^{-2 pi i j k / n} (analysis).  n = 4 i, M >= 2 i - 1 a power of two >= 1024.
extern "C" void emul_ring_split(const double *inv, double *outv, int n, int M, int dir) {
  const double2 *in = reinterpret_cast<const double2 *>(inv);
  double2 *out = reinterpret_cast<double2 *>(outv);
  const int i = n >> 2;
  std::vector<double2> tw = twiddles2(M), v = sub_filter(i, M, tw.data());
  std::vector<double2> zbuf(n), work(bf_padded(M));
  if (dir == 0) {
    for (int k = 0; k < n; ++k) zbuf[rs_slot(k, i)] = cmul(in[k], rs_expipi(rs_fold_angle(k, 0), n));
  } else {
    for (int t = 0; t < i; ++t) {
      double2 y[4], o[4];
      for (int q = 0; q < 4; ++q) { y[q].x = in[t + q * i].x; y[q].y = -in[t + q * i].y; }
      rs_butterfly4(y, o);
      for (int r = 0; r < 4; ++r) zbuf[r * i + t] = cmul(o[r], rs_expipi(rs_twiddle_angle(t, r), n));
    }
  }
  for (int r = 0; r < 4; ++r) {
    for (int t = 0; t < M; ++t) { double2 z; z.x = z.y = 0.0; work[bf_pidx<true>(t)] = t < i ? zbuf[r * i + t] : z; }
    convolve(work.data(), M, tw.data(), v.data());
    for (int t = 0; t < i; ++t) zbuf[r * i + t] = work[bf_pidx<true>(t)];
  }
  const double invM = 1.0 / M;
  if (dir == 0) {
    for (int t = 0; t < i; ++t) {
      double2 cr[4], o[4];
      for (int r = 0; r < 4; ++r) cr[r] = cmul(zbuf[r * i + t], rs_expipi(rs_twiddle_angle(t, r), n));
      rs_butterfly4(cr, o);
      for (int q = 0; q < 4; ++q) { out[t + q * i].x = invM * o[q].x; out[t + q * i].y = invM * o[q].y; }
    }
  } else {
    for (int k = 0; k < n; ++k) {
      const long long a = k >> 2;
      const double2 g = cmul(zbuf[rs_slot(k, i)], rs_expipi(4 * a * a, n));
      out[k].x = g.x * invM; out[k].y = -g.y * invM;
    }
  }
}

// whole-ring power-of-two transform: dir 0 inverse (spectrum written to bit-reversed slots, DIT), dir 1 forward (DIF,
// bins read from bit-reversed slots)
extern "C" void emul_ring_pow2(const double *inv, double *outv, int n, int dir) {
  const double2 *in = reinterpret_cast<const double2 *>(inv);
  double2 *out = reinterpret_cast<double2 *>(outv);
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  std::vector<double2> tw = twiddles2(n), work(bf_padded(n));
  if (dir == 0) {
    for (int k = 0; k < n; ++k) work[bf_pidx<true>((int)bf_bitrev((unsigned)k, bits))] = in[k];
    fft_dit(work.data(), n, tw.data());
    for (int j = 0; j < n; ++j) out[j] = work[bf_pidx<true>(j)];
  } else {
    for (int j = 0; j < n; ++j) work[bf_pidx<true>(j)] = in[j];
    fft_dif(work.data(), n, tw.data());
    for (int k = 0; k < n; ++k) out[k] = work[bf_pidx<true>((int)bf_bitrev((unsigned)k, bits))];
  }
}

// one ring of length n = 2 h through one complex transform of length h (ring_half_kernel): dir 0 takes the Hermitian
// spectrum X (n complex) and returns the n real samples, dir 1 takes n real samples and returns X_b, b < n
extern "C" void emul_ring_half(const double *inv, double *outv, int n, int dir) {
  const int h = n / 2;
  int bits = 0;
  while ((1 << bits) < h) ++bits;
  std::vector<double2> tw = twiddles2(h), work(bf_padded(h));
  if (dir == 0) {
    const double2 *X = reinterpret_cast<const double2 *>(inv);
    for (int k = 0; k < h; ++k)
      work[bf_pidx<true>((int)bf_bitrev((unsigned)k, bits))] = rh_pack(X[k], X[k + h], rs_expipi(k, h));
    fft_dit(work.data(), h, tw.data());
    for (int p = 0; p < h; ++p) { outv[2 * p] = work[bf_pidx<true>(p)].x; outv[2 * p + 1] = work[bf_pidx<true>(p)].y; }
  } else {
    double2 *X = reinterpret_cast<double2 *>(outv);
    for (int p = 0; p < h; ++p) { work[bf_pidx<true>(p)].x = inv[2 * p]; work[bf_pidx<true>(p)].y = inv[2 * p + 1]; }
    fft_dif(work.data(), h, tw.data());
    for (int b = 0; b < n; ++b) {
      const int k = b & (h - 1), k2 = (h - k) & (h - 1);
      double2 wb = rs_expipi(2LL * b, n); wb.y = -wb.y;
      X[b] = rh_unpack(work[bf_pidx<true>((int)bf_bitrev((unsigned)k, bits))], work[bf_pidx<true>((int)bf_bitrev((unsigned)k2, bits))], wb);
    }
  }
}
