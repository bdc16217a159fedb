// Host emulation of the per-ring Legendre recurrence used by the CUDA kernels
// (commander_b200/csrc/legendre_core.cuh + coef.cpp), compiled with g++ for the CPU
// test-suite.  TEST ONLY: it exercises the product's math headers on the host so that
// sign / normalisation / scaling bugs are caught without a GPU.  It is not a fallback.
#include <algorithm>
#include <vector>
#include <cmath>
#include "../../commander_b200/csrc/legendre_core.cuh"
#include "../../commander_b200/csrc/sht_internal.h"
using namespace cmdr;
// spin 0 as the transform kernels run it: two l per recurrence step (step0x2 + build_coef_table_x2), lambda of both
// parities rebuilt from the nu_j with the mix rows exactly as prep_s0_kernel / the flush of anal0_kernel use them
static int emul_lambda_x2(int lmax, int m, const RingTrig &g, double *out) {
  std::vector<int> mval{m};
  std::vector<double> rec, mix; std::vector<long long> ofs;
  build_coef_table_x2(lmax, mval, rec, mix, ofs);
  std::vector<double> K0, K2;
  build_start_norms(m, K0, K2);
  const double SD = ldexp(1.0, -SCALE_BITS);
  for (int l = 0; l <= lmax; ++l) out[l] = 0;
  if (m > lmax) return 0;
  const int J = (lmax - m) / 2 + 1;
  int k; double cur, prev = 0, nu_prev_on = 0;
  start_spin0(m, K0[m], g, cur, k);
  const double x2 = g.cth * g.cth;
  int cnt = 0;
  for (int j = 0; j < J; ++j) {
    const double *mx = &mix[4 * j];
    const double nu = k == 0 ? cur : 0.0;
    const int l = m + 2 * j;
    out[l] = mx[0] * nu + mx[3] * nu_prev_on;
    if (l + 1 <= lmax) out[l + 1] = g.cth * mx[2] * nu;
    nu_prev_on = nu;
    double nxt = step0x2(rec[2 * j], rec[2 * j + 1], x2, cur, prev); prev = cur; cur = nxt;
    if (++cnt == 4) { cnt = 0; if (k < 0 && needs_rescale(cur)) { cur *= SD; prev *= SD; ++k; } }
  }
  return 0;
}

// spin 2 with the scalar front phase of the kernels: scalar two-l-per-step recurrence until the ring is within
// FRONT_MARGIN_BITS of the threshold (checked every 4 l), conversion at the last multiple of 8 rows (the kernels'
// FRONT_TR) before that point -- or at row `force_row` >= 0: in a warp the first ring to get there switches all of
// them --, then the spin-2 recurrences as in emul_lambda
extern "C" int emul_lambda_front(int lmax, int m, int nside, int north, int force_row, double *outP, double *outM) {
  std::vector<int> mval{m};
  std::vector<double> tab, rec, mix; std::vector<long long> ofs, ofs0;
  build_coef_table(lmax, 2, mval, tab, ofs);
  build_coef_table_x2(lmax, mval, rec, mix, ofs0);
  std::vector<double> K0, K2;
  build_start_norms(m, K0, K2);
  long double omc, ns = nside;
  if (north < nside) omc = (long double)north * north / (3.0L * ns * ns);
  else omc = 1.0L - (2.0L * ns - north) * 2.0L / (3.0L * ns);
  RingTrig g{(double)(1.0L - omc), (double)sqrtl(omc * (2.0L - omc)), (double)sqrtl(0.5L * omc), (double)sqrtl(1.0L - 0.5L * omc)};
  const double SD = ldexp(1.0, -SCALE_BITS);
  for (int l = 0; l <= lmax; ++l) { outP[l] = 0; outM[l] = 0; }
  if (m < 2 || m > lmax) return -1;
  int k; double cur, prev = 0;
  start_spin0(m, K0[m], g, cur, k);
  if (k >= 0) return -1;                         // not eligible: the kernels use the plain path
  const double x2 = g.cth * g.cth;
  const int J = (lmax - m) / 2 + 1;
  // first pass: the row at which this ring asks for the switch (checked every two rows = 4 l)
  int jt = 0;
  {
    double c = cur, pr = prev; int kk = k;
    for (;; jt += 2) {
      if (jt + 2 > J - 1 || front_must_switch(c, kk, front_margin_bits(g.sth))) break;
      for (int q = 0; q < 2; ++q) { double nxt = step0x2(rec[2 * (jt + q)], rec[2 * (jt + q) + 1], x2, c, pr); pr = c; c = nxt; }
      if (kk < 0 && needs_rescale(c)) { c *= SD; pr *= SD; ++kk; }
    }
  }
  const int jb = force_row >= 0 ? std::min(force_row & ~1, (J - 1) & ~1) : jt / 8 * 8;
  for (int j = 0; j < jb; j += 2) {
    for (int q = 0; q < 2; ++q) { double nxt = step0x2(rec[2 * (j + q)], rec[2 * (j + q) + 1], x2, cur, prev); prev = cur; cur = nxt; }
    if (k < 0 && needs_rescale(cur)) { cur *= SD; prev *= SD; ++k; }
  }
  const int lb = m + 2 * jb;
  double P, Pp, M, Mp;
  spin2_front_convert(lb, m, lmax, g.cth, g.sth, cur, prev, &mix[4 * jb], jb > 0 ? mix[4 * (jb - 1) + 2] : 0.0,
                      &tab[4 * (lb - m)], &tab[4 * (lb - m) + 4], P, Pp, M, Mp);
  int cnt = 0;
  for (int l = lb; l <= lmax; ++l) {
    const double *c = &tab[4 * (l - m)];
    outP[l] = k == 0 ? c[2] * P : 0.0; outM[l] = k == 0 ? c[2] * M : 0.0;
    double up = fma(c[0], g.cth, c[1]), um = fma(c[0], g.cth, -c[1]);
    double np_ = fma(up, P, -Pp), nm_ = fma(um, M, -Mp);
    Pp = P; P = np_; Mp = M; M = nm_;
    if (++cnt == 4) { cnt = 0; if (k < 0 && (needs_rescale(P) || needs_rescale(M))) { P *= SD; Pp *= SD; M *= SD; Mp *= SD; ++k; } }
  }
  return lb;
}

extern "C" int emul_lambda(int spin, int lmax, int m, int nside, int north, double *outP, double *outM) {
  std::vector<int> mval{m};
  std::vector<double> tab; std::vector<long long> ofs;
  build_coef_table(lmax, spin, mval, tab, ofs);
  std::vector<double> K0, K2, Ks;
  build_start_norms(m, K0, K2);
  if (spin != 0 && spin != 2) build_start_norms_spin(m, spin, Ks);
  long double omc, ns = nside;
  if (north < nside) omc = (long double)north * north / (3.0L * ns * ns);
  else omc = 1.0L - (2.0L * ns - north) * 2.0L / (3.0L * ns);
  RingTrig g{(double)(1.0L - omc), (double)sqrtl(omc * (2.0L - omc)), (double)sqrtl(0.5L * omc), (double)sqrtl(1.0L - 0.5L * omc)};
  if (spin == -1) return emul_lambda_x2(lmax, m, g, outP);
  const double SD = ldexp(1.0, -SCALE_BITS);
  int l0 = spin == 0 ? m : (m > spin ? m : spin);
  for (int l = 0; l <= lmax; ++l) { outP[l] = 0; if (outM) outM[l] = 0; }
  if (l0 > lmax) return 0;
  int k;
  if (spin == 0) {
    double cur, prev = 0;
    start_spin0(m, K0[m], g, cur, k);
    int cnt = 0;
    for (int l = m; l <= lmax; ++l) {
      double A = tab[2 * (l - m)], gg = tab[2 * (l - m) + 1];
      outP[l] = k == 0 ? gg * cur : 0.0;
      double nxt = step0(A, g.cth, cur, prev); prev = cur; cur = nxt;
      if (++cnt == 8) { cnt = 0; if (k < 0 && needs_rescale(cur)) { cur *= SD; prev *= SD; ++k; } }
    }
  } else {
    double P, M, Pp = 0, Mp = 0;
    if (spin == 2) start_spin2(m, K2[m], g, P, M, k); else start_spin_s(m, spin, Ks[m], g, P, M, k);
    int cnt = 0;
    for (int l = l0; l <= lmax; ++l) {
      const double *c = &tab[4 * (l - l0)];
      outP[l] = k == 0 ? c[2] * P : 0.0; outM[l] = k == 0 ? c[2] * M : 0.0;
      double up = fma(c[0], g.cth, c[1]), um = fma(c[0], g.cth, -c[1]);
      double np_ = fma(up, P, -Pp), nm_ = fma(um, M, -Mp);
      Pp = P; P = np_; Mp = M; M = nm_;
      if (++cnt == 8) { cnt = 0; if (k < 0 && (needs_rescale(P) || needs_rescale(M))) { P *= SD; Pp *= SD; M *= SD; Mp *= SD; ++k; } }
    }
  }
  return 0;
}
