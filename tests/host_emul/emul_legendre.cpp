// Host emulation of the per-ring Legendre recurrence used by the CUDA kernels
// (commander_b200/csrc/legendre_core.cuh + coef.cpp), compiled with g++ for the CPU
// test-suite.  TEST ONLY: it exercises the product's math headers on the host so that
// sign / normalisation / scaling bugs are caught without a GPU.  It is not a fallback.
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
