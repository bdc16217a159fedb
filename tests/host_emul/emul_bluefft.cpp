// Host emulation of the in-place FFT pair of the fused Bluestein ring kernels (commander_b200/csrc/blue_fft.cuh),
// compiled with g++ for the CPU test-suite: the device code's work items are run in a loop over "threads", pass by
// pass (a pass boundary is a __syncthreads() on the device).  TEST ONLY, not a fallback.
#define BLUE_FFT_HOST
#include <cmath>
#include <vector>
#include "../../commander_b200/csrc/blue_fft.cuh"
using namespace cmdr;

static std::vector<double2> twiddles(int M) {             // pass-major, as ringfft.cu's twiddle_kernel builds them
  std::vector<double2> tw(bf_tw_total(M) + 1);
  const int np = bf_num_passes(M);
  for (int k = 0; k < np; ++k) {
    int h, fused;
    bf_pass(M, k, &h, &fused);
    if (!fused) continue;
    double2 *T = tw.data() + bf_tw_offset(M, k);
    for (int j = 0; j < h / 2; ++j) { T[j].x = std::cos(-M_PI * j / h); T[j].y = std::sin(-M_PI * j / h); }
  }
  return tw;
}

// x: M interleaved complex doubles, in place.  dir 0: DIF forward (bit-reversed output); 1: DIT inverse (bit-reversed
// input, unnormalised).  nthreads only changes the order in which the items of a pass are visited.
extern "C" void emul_fft(double *xv, int M, int dir, int nthreads) {
  double2 *x = reinterpret_cast<double2 *>(xv);
  std::vector<double2> tw = twiddles(M);
  const int np = bf_num_passes(M);
  for (int kk = 0; kk < np; ++kk) {
    const int k = dir == 0 ? kk : np - 1 - kk;
    int h, fused;
    bf_pass(M, k, &h, &fused);
    const double2 *T = tw.data() + bf_tw_offset(M, k);
    const int items = fused ? M / 4 : M / 2;
    for (int tid = 0; tid < nthreads; ++tid)
      for (int q = tid; q < items; q += nthreads) {
        if (dir == 0) { if (fused) dif_item4(x, h, T, q); else dif_item2(x, q); }
        else          { if (fused) dit_item4(x, h, T, q); else dit_item2(x, q); }
      }
  }
}

// the register-blocked, padded variant (bf2_* schedule): x has bf_padded(M) slots
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
static void pass2(double2 *x, int M, int S, int h, const double2 *T, int nthreads) {
  const int items = M >> S;
  for (int tid = 0; tid < nthreads; ++tid)
    for (int q = tid; q < items; q += nthreads) {
      if (S == 3) { if (DIR == 0) dif_itemS<3, true>(x, h, T, q); else dit_itemS<3, true>(x, h, T, q); }
      else        { if (DIR == 0) dif_itemS<4, true>(x, h, T, q); else dit_itemS<4, true>(x, h, T, q); }
    }
}

extern "C" void emul_fft2(double *xv, int M, int dir, int nthreads) {
  double2 *x = reinterpret_cast<double2 *>(xv);
  std::vector<double2> tw = twiddles2(M);
  const int ns = bf2_num_strided(M);
  if (dir == 0) {
    for (int k = 0; k < ns; ++k) pass2<0>(x, M, bf2_stages(M, k), bf2_half(M, k), tw.data() + bf2_tw_offset(M, k), nthreads);
    pass2<0>(x, M, 4, 8, tw.data() + bf2_tw_total(M) - 1, nthreads);
  } else {
    pass2<1>(x, M, 4, 8, tw.data() + bf2_tw_total(M) - 1, nthreads);
    for (int k = ns - 1; k >= 0; --k) pass2<1>(x, M, bf2_stages(M, k), bf2_half(M, k), tw.data() + bf2_tw_offset(M, k), nthreads);
  }
}

// u -> IDFT(DFT(u) .* V) * M with V given in bit-reversed order, exactly as blue_fused2_kernel sequences it:
// strided DIF passes, conv_mid16, strided DIT passes
extern "C" void emul_conv2(double *xv, const double *vbr, int M, int nthreads) {
  double2 *x = reinterpret_cast<double2 *>(xv);
  const double2 *v = reinterpret_cast<const double2 *>(vbr);
  std::vector<double2> tw = twiddles2(M);
  const int ns = bf2_num_strided(M);
  for (int k = 0; k < ns; ++k) pass2<0>(x, M, bf2_stages(M, k), bf2_half(M, k), tw.data() + bf2_tw_offset(M, k), nthreads);
  for (int tid = 0; tid < nthreads; ++tid)
    for (int q = tid; q < M / 16; q += nthreads) conv_mid16<true>(x, v, q);
  for (int k = ns - 1; k >= 0; --k) pass2<1>(x, M, bf2_stages(M, k), bf2_half(M, k), tw.data() + bf2_tw_offset(M, k), nthreads);
}

extern "C" unsigned emul_bitrev(unsigned v, int bits) { return bf_bitrev(v, bits); }
