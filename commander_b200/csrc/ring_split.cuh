// ring_split.cuh -- index / phase arithmetic of the two ring transforms that keep a whole ring pair in shared
// memory (ringfft.cu: ring_split_kernel, ring_pow2_kernel).  Shared with the host emulation test
// (tests/host_emul/emul_ringsplit.cpp), hence free of CUDA-only constructs.
//
// (1) Long polar-cap rings.  A HEALPix cap ring has n = 4 i points (i = ring number), an awkward length for every
//     i that is not smooth; chirp-z (Bluestein) on the whole ring needs a power-of-two work length >= 2 n - 1, i.e.
//     16384 complex = 256 KB for n > 4096 -- more than an SM's shared memory, which is why round 1 sent these
//     rings (75 % of the cap pixels at nside 2048) through cuFFT and HBM 13 times.  But n is always a multiple
//     of 4, so one radix-4 Cooley-Tukey step turns the length-n transform into FOUR length-i transforms plus a
//     4-point butterfly per output, and a length-i chirp-z needs only M >= 2 i - 1 <= 4096 (64 KB):
//         synthesis  x_j = sum_k Z_k e^{+2 pi i j k / n},  k = 4 a + r,  j = t + i q:
//                    x_{t + i q} = sum_r  i^{q r}  e^{2 pi i t r / n}  Y_r[t],    Y_r = DFT+_i ( Z_{4 a + r} )_a
//         analysis   Z_{4 a + r} = conj DFT+_i ( u_r )_a,   u_r[t] = e^{2 pi i t r / n} sum_q i^{q r} conj(z_{t + i q})
//     and DFT+_i by chirp-z:  Y[t] = c_t / M * IFFT_M( FFT_M(c . y) .* V_i )[t],  c_t = e^{i pi t^2 / i}.
//     The four sub-spectra live in one n-element buffer ("zbuf", r-major) that first holds the folded input,
//     then (in place, one r after the other through the M-element work area) the raw convolutions, so HBM sees
//     the phases and the map exactly once.
// (2) Belt rings of power-of-two length: one in-place FFT of the whole ring (fold straight into bit-reversed
//     order -> DIT inverse -> natural order, or DIF forward -> unfold from bit-reversed order).
#pragma once
#include "blue_fft.cuh"

namespace cmdr {

// exp(i pi num / den) with exact integer range reduction (num >= 0)
#ifdef BLUE_FFT_HOST
inline double2 rs_expipi(long long num, long long den) {
  const long long r = num % (2 * den);
  double2 v; v.x = std::cos(M_PI * (double)r / (double)den); v.y = std::sin(M_PI * (double)r / (double)den);
  return v;
}
#else
__device__ __forceinline__ double2 rs_expipi(long long num, long long den) {
  const long long r = num % (2 * den);
  double s, c;
  sincospi((double)r / (double)den, &s, &c);
  double2 v; v.x = c; v.y = s;
  return v;
}
#endif

// position of bin k = 4 a + r in the r-major buffer of four sub-spectra of length i
BF_HD int rs_slot(int k, int i) { return (k & 3) * i + (k >> 2); }

// factor applied to folded bin k on the way into zbuf: phi0 shift e^{i pi k / n} (shifted rings) times the input
// chirp e^{i pi a^2 / i} of its sub-transform -- one sincospi of the combined angle (k + 4 a^2) / n
BF_HD long long rs_fold_angle(int k, int shifted) {
  const long long a = k >> 2;
  return (shifted ? (long long)k : 0LL) + 4 * a * a;
}

// factor of sub-spectrum r at position t in the radix-4 pass: twiddle e^{2 pi i t r / n} times chirp e^{i pi t^2 / i}
// = exp(i pi (4 t^2 + 2 t r) / n)
BF_HD long long rs_twiddle_angle(int t, int r) { return 4LL * t * t + 2LL * t * r; }

// out_q = sum_r i^{q r} c_r   (the 4-point DFT with e^{+2 pi i q r / 4})
BF_HD void rs_butterfly4(const double2 (&c)[4], double2 (&o)[4]) {
  const double2 s02 = bf_add(c[0], c[2]), d02 = bf_sub(c[0], c[2]);
  const double2 s13 = bf_add(c[1], c[3]), d13 = bf_mul_pi(bf_sub(c[1], c[3]));   // i (c1 - c3)
  o[0] = bf_add(s02, s13);
  o[1] = bf_add(d02, d13);
  o[2] = bf_sub(s02, s13);
  o[3] = bf_sub(d02, d13);
}

// (3) One ring of even length n = 2 h through ONE complex transform of length h (ring_half_kernel).  The samples of
//     a ring are real, so its spectrum X_b = sum_j x_j e^{-2 pi i j b / n} is Hermitian and half of a length-n complex
//     transform is redundant; the whole-ring kernels use that redundancy to pack the north ring and its southern
//     mirror into one complex transform (z = x_N + i x_S), at the price of a footprint of n complex numbers per
//     CTA (1 CTA per SM for the belt at nside 2048).  Packing even and odd samples of ONE ring instead,
//         y_p = x_{2p} + i x_{2p+1},   p < h,
//     costs the same arithmetic per ring pair and halves the footprint, so that 3 CTAs share an SM:
//         synthesis  Y_k = (X_k + X_{k+h}) + i w^k (X_k - X_{k+h}),  w = e^{2 pi i / n};   y = DFT+_h(Y)
//         analysis   Y = DFT-_h(y);  X_b = E_k + e^{-2 pi i b / n} O_k,  k = b mod h,
//                    E_k = (Y_k + conj Y_{h-k}) / 2,  O_k = (Y_k - conj Y_{h-k}) / (2 i)
BF_HD double2 rh_pack(double2 Xk, double2 Xkh, double2 wk) {           // wk = e^{i pi k / h}
  const double2 a = bf_add(Xk, Xkh), b = bf_mul_pi(bf_mul(bf_sub(Xk, Xkh), wk));
  return bf_add(a, b);
}
BF_HD double2 rh_unpack(double2 Ya, double2 Yb, double2 wb) {          // Ya = Y_k, Yb = Y_{(h-k) mod h}, wb = e^{-2 pi i b / n}
  double2 E, O;
  E.x = 0.5 * (Ya.x + Yb.x); E.y = 0.5 * (Ya.y - Yb.y);               // (Ya + conj Yb) / 2
  O.x = 0.5 * (Ya.y + Yb.y); O.y = -0.5 * (Ya.x - Yb.x);              // (Ya - conj Yb) / (2 i)
  return bf_add(E, bf_mul(O, wb));
}

}  // namespace cmdr
