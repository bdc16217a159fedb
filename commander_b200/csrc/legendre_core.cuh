// legendre_core.cuh -- per-ring-pair Legendre / Wigner-d recurrence state shared by the
// synthesis and analysis kernels (and by the host emulation test, which compiles this
// header with a plain C++ compiler: CMDR_HD expands to nothing there).
//
// Replaces the Legendre stage inside libsharp2's sharp_execute, reached from
// commander3/src/sharp.f90:226-240.  The only Legendre
// arithmetic physically in the reference is commander3/src/math_tools.f90:926-1028.
//
// Formulation (differs from both libsharp2 and the oracle on purpose -- it is the
// cheapest one in FP64 issue slots):
//   lambda_l = g_l * mu_l, where g is chosen so the three-term recurrence reads
//       mu_{l+1} = (A'_l x [+- C'_l]) mu_l - mu_{l-1}          (2 DFMA-pipe ops / l / function)
//   g_l is folded into the a_lm on load (synthesis) or on store (analysis).
//   Tiny values are carried as mu * 2^(SCALE_BITS*k), k<=0; a ring joins the
//   accumulation when k reaches 0, i.e. when |mu| >= 2^-THRESH_BITS.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CMDR_HD __host__ __device__ __forceinline__
#else
#define CMDR_HD inline
#endif

namespace cmdr {

constexpr int SCALE_BITS = 512;    // one scale step
constexpr int THRESH_BITS = 70;    // accumulate once |mu| >= 2^-70 (libsharp2 starts at 2^-60, sharp_ftol; g_l = O(1..100) leaves margin)
// rescale when the biased exponent reaches (SCALE_BITS-THRESH_BITS)+1023
constexpr int RESCALE_EXP = SCALE_BITS - THRESH_BITS + 1023;

CMDR_HD int hi_word(double v) {
#ifdef __CUDA_ARCH__
  return __double2hiint(v);
#else
  union { double d; uint64_t u; } c; c.d = v; return (int)(c.u >> 32);
#endif
}
CMDR_HD double pow2i(int e) {   // 2^e for -1022 <= e <= 1023
#ifdef __CUDA_ARCH__
  return __hiloint2double((e + 1023) << 20, 0);
#else
  union { double d; uint64_t u; } c; c.u = (uint64_t)(e + 1023) << 52; return c.d;
#endif
}
CMDR_HD bool needs_rescale(double v) { return ((hi_word(v) >> 20) & 0x7ff) >= RESCALE_EXP; }

// base^n as mant * 2^ex with mant in [0.5,1);  0 < base <= 1, n >= 0.
CMDR_HD void pow_scaled(double base, int n, double &mant, int &ex) {
  int be;
  double bm = frexp(base, &be);
  double rm = 0.5; int re = 1;            // 1.0 = 0.5 * 2^1
  while (n) {
    if (n & 1) { rm *= bm; re += be; if (rm < 0.5) { rm *= 2.0; re -= 1; } }
    bm *= bm; be *= 2; if (bm < 0.5) { bm *= 2.0; be -= 1; }
    n >>= 1;
  }
  mant = rm; ex = re;
}

// Splits mant*2^ex (mant in [0.5,1)) into stored = mant*2^(ex-SCALE_BITS*k), k<=0.
CMDR_HD double split_scale(double mant, int ex, int &k) {
  int t = ex + THRESH_BITS;
  // floor division by SCALE_BITS (power of two) for negative t
  int kk = t >= 0 ? 0 : -((-t + SCALE_BITS - 1) / SCALE_BITS);
  k = kk;
  int e = ex - SCALE_BITS * kk;            // in [-THRESH_BITS, SCALE_BITS-THRESH_BITS) or >= -THRESH_BITS when kk==0
  return mant * pow2i(e);
}

// Per-ring geometry the kernels need.
struct RingTrig { double cth, sth, sh, ch; };   // cos/sin(theta), sin/cos(theta/2) of the NORTH ring

// ---- spin 0 --------------------------------------------------------------
// start: lambda_mm = K0[m] * sth^m      (K0 carries (-1)^m sqrt((2m+1)/4pi (2m-1)!!/(2m)!!))
CMDR_HD void start_spin0(int m, double K0m, const RingTrig &g, double &mu, int &k) {
  double mant; int ex;
  pow_scaled(g.sth, m, mant, ex);
  mu = K0m * split_scale(mant, ex, k);
}
// one step: returns mu_{l+1}
CMDR_HD double step0(double A, double x, double cur, double prev) {
  return fma(A * x, cur, -prev);
}

// one step of the two-l-per-step form (coef.cpp, fill_spin0_x2): returns nu_{j+1}; x2 = x^2
CMDR_HD double step0x2(double A, double B, double x2, double cur, double prev) {
  return fma(fma(A, x2, B), cur, -prev);
}

// ---- spin 2 --------------------------------------------------------------
// P = (+2)lambda_{l0,m}, M = (-2)lambda_{l0,m} at l0 = max(m,2).
//   m>=2: P = K2[m] sth^(m-2) sh^4 , M = K2[m] sth^(m-2) ch^4
//   m==0: P = M = c20 * sth^2 ;  m==1: P = -c21 ch sh^3, M = +c21 ch^3 sh
CMDR_HD void start_spin2(int m, double K2m, const RingTrig &g, double &P, double &M, int &k) {
  if (m >= 2) {
    double mant; int ex;
    pow_scaled(g.sth, m - 2, mant, ex);
    double base = K2m * split_scale(mant, ex, k);
    double s2 = g.sh * g.sh, c2 = g.ch * g.ch;
    P = base * (s2 * s2);
    M = base * (c2 * c2);
  } else if (m == 0) {
    k = 0;
    P = M = K2m * g.sth * g.sth;
  } else {
    k = 0;
    P = -K2m * g.ch * (g.sh * g.sh * g.sh);
    M = K2m * (g.ch * g.ch * g.ch) * g.sh;
  }
}

// ---- arbitrary spin s >= 1 (conviqt: commander3/src/comm_conviqt_mod.f90:234-239) ---------------
CMDR_HD double ipow(double b, int n) {      // b^n, n >= 0, by repeated squaring
  double r = 1.0;
  while (n) { if (n & 1) r *= b; b *= b; n >>= 1; }
  return r;
}
// P = (+s)lambda_{l0,m}, M = (-s)lambda_{l0,m} at l0 = max(m,s); Ksm from build_start_norms_spin.
// (+-s)lambda_lm(theta) = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,+-s}(theta)   (Goldberg et al. 1967)
CMDR_HD void start_spin_s(int m, int s, double Ksm, const RingTrig &g, double &P, double &M, int &k) {
  if (m >= s) {
    double mant; int ex;
    pow_scaled(g.sth, m - s, mant, ex);
    double base = Ksm * split_scale(mant, ex, k);
    P = base * ipow(g.sh, 2 * s);
    M = base * ipow(g.ch, 2 * s);
  } else {
    k = 0;
    P = Ksm * ipow(g.ch, s - m) * ipow(g.sh, s + m);
    M = (((s - m) & 1) ? -Ksm : Ksm) * ipow(g.ch, s + m) * ipow(g.sh, s - m);
  }
}

// libsharp-style per-ring m cut-off (contributions above it are far below FP64
// resolution).  Same form as the oracle's get_mlim.
inline int mlim_for_ring(int lmax, int spin, double sth, double cth) {
  double ofs = lmax * 0.01; if (ofs < 100.) ofs = 100.;
  double b = -2 * spin * fabs(cth);
  double t1 = lmax * sth + ofs;
  double c = (double)spin * spin - t1 * t1;
  double discr = b * b - 4 * c;
  if (discr <= 0) return lmax;
  double res = (-b + sqrt(discr)) / 2.;
  if (res > lmax) res = lmax;
  return (int)(res + 0.5);
}

}  // namespace cmdr
