// coef.cpp -- host-side (long double) construction of the recurrence coefficient
// tables and start-value normalisations used by the Legendre kernels.
//
// libsharp2 builds the equivalent tables in its Ylmgen setup, reached from
// sharp_execute (commander3/src/sharp.f90:234).  Here they are computed once per
// (alm_info, spin) in 80-bit arithmetic, rounded to FP64 and uploaded.
//
// spin 0:  x lam_l = b_{l+1} lam_{l+1} + b_l lam_{l-1},  b_l = sqrt((l^2-m^2)/(4l^2-1))
//          (the recurrence of commander3/src/math_tools.f90:1003-1023)
//          lam_l = g_l mu_l,  g_m = g_{m+1} = 1,  g_{l+1} = g_{l-1} b_l / b_{l+1}
//          mu_{l+1} = A'_l x mu_l - mu_{l-1},      A'_l = g_l / (g_{l+1} b_{l+1})
// spin s:  L_{l+1} = al_l (x +- C_l) L_l - be_l L_{l-1}   (Wigner d^l_{-m,+-s} times sqrt((2l+1)/4pi))
//          al_l = sqrt((2l+3)/(2l+1)) (2l+1)(l+1)/D_{l+1},  D_l = sqrt((l^2-m^2)(l^2-s^2))
//          be_l = sqrt((2l+3)/(2l-1)) (l+1) D_l / (l D_{l+1}),   C_l = m s / (l (l+1))
//          g_{l0} = g_{l0+1} = 1, g_{l+1} = be_l g_{l-1};  A'_l = al_l g_l / g_{l+1},  C'_l = A'_l C_l
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "sht_internal.h"

namespace cmdr {

static void fill_spin0(int lmax, int m, double *out /* {A', g} per l = m..lmax */) {
  auto beta = [m](int l) -> long double {
    return sqrtl(((long double)l * l - (long double)m * m) / (4.0L * l * l - 1.0L));
  };
  long double g_prev = 1.0L, g_cur = 1.0L;   // g_{l-1}, g_l
  for (int l = m; l <= lmax; ++l) {
    long double b1 = beta(l + 1);
    long double g_next = (l == m) ? 1.0L : g_prev * beta(l) / b1;
    long double A = g_cur / (g_next * b1);
    out[2 * (l - m)] = (double)A;
    out[2 * (l - m) + 1] = (double)g_cur;
    g_prev = g_cur; g_cur = g_next;
  }
}

// spin 0, two l per step (the scheme libsharp2 itself uses for scalar transforms): with l_j = m + 2j and
// q_j = lam_{l_j+1} / x, the three-term recurrence above applied twice gives a recurrence in x^2,
//     al_j q_{j+1} = (x^2 - be_j) q_j - al_{j-1} q_{j-1},  al_j = b_{l_j+2} b_{l_j+3},  be_j = b_{l_j+1}^2 + b_{l_j+2}^2,
// and both parities of lam come from the q_j alone:
//     lam_{l_j} = b_{l_j+1} q_j + b_{l_j} q_{j-1},      lam_{l_j+1} = x q_j.
// Normalised as q_j = h_j nu_j with nu_{j+1} = (a_j x^2 + b'_j) nu_j - nu_{j-1}  (2 DFMA-pipe ops per TWO l):
//     h_0 = h_1 = sqrt(2m+3) (so that nu_0 = lam_mm, the start value of the one-step form),
//     h_{j+1} = al_{j-1} h_{j-1} / al_j,   a_j = h_j / (al_j h_{j+1}),   b'_j = -be_j a_j.
// rec row j = {a_j, b'_j};  mix row j = {u_j, v_j, h_j, v_{j-1}} with u_j = h_j b_{l_j+1}, v_j = h_j b_{l_j+2}:
//     sum_l a_l lam_l = sum_j nu_j (u_j a_{l_j} + v_j a_{l_j+2}) + x sum_j nu_j h_j a_{l_j+1}.
static void fill_spin0_x2(int lmax, int m, double *rec, double *mix) {
  auto beta = [m](int l) -> long double {
    return sqrtl(((long double)l * l - (long double)m * m) / (4.0L * l * l - 1.0L));
  };
  const int J = (lmax - m) / 2 + 1;
  long double h_prev = sqrtl(2.0L * m + 3.0L), h_cur = h_prev;   // h_{j-1}, h_j
  long double al_prev = 0.0L, v_prev = 0.0L;                      // al_{j-1}, v_{j-1}
  for (int j = 0; j < J; ++j) {
    const int l = m + 2 * j;
    const long double b1 = beta(l + 1), b2 = beta(l + 2), b3 = beta(l + 3);
    const long double al = b2 * b3, be = b1 * b1 + b2 * b2;
    const long double h_next = j == 0 ? h_cur : al_prev * h_prev / al;
    const long double a = h_cur / (al * h_next);
    rec[2 * j] = (double)a; rec[2 * j + 1] = (double)(-be * a);
    mix[4 * j] = (double)(h_cur * b1); mix[4 * j + 1] = (double)(h_cur * b2); mix[4 * j + 2] = (double)h_cur; mix[4 * j + 3] = (double)v_prev;
    v_prev = h_cur * b2;
    h_prev = h_cur; h_cur = h_next; al_prev = al;
  }
}

void build_coef_table_x2(int lmax, const std::vector<int> &mval, std::vector<double> &rec, std::vector<double> &mix,
                         std::vector<long long> &ofs) {
  const int nm = (int)mval.size();
  ofs.assign(nm + 1, 0);
  for (int i = 0; i < nm; ++i) {
    long long n = lmax >= mval[i] ? (lmax - mval[i]) / 2 + 1 : 0;
    ofs[i + 1] = ofs[i] + n * 2;
  }
  // zero padding: whole tiles of up to 256 rows are staged with cp.async past the last row of the last m
  rec.assign((size_t)ofs[nm] + 2 * 256 + 4, 0.0);
  mix.assign((size_t)2 * ofs[nm] + 4 * 256 + 4, 0.0);
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if ((int)nt > nm) nt = nm > 0 ? nm : 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t]() {
      for (int i = (int)t; i < nm; i += (int)nt)
        if (lmax >= mval[i]) fill_spin0_x2(lmax, mval[i], rec.data() + ofs[i], mix.data() + 2 * ofs[i]);
    });
  }
  for (auto &t : th) t.join();
}

static void fill_spins(int lmax, int m, int s, double *out /* {A', C', g, pad} per l = l0..lmax */) {
  int l0 = m > s ? m : s;
  auto D = [m, s](int l) -> long double {
    return sqrtl(((long double)l * l - (long double)m * m) * ((long double)l * l - (long double)s * s));
  };
  long double g_prev = 1.0L, g_cur = 1.0L;
  for (int l = l0; l <= lmax; ++l) {
    long double l1 = l + 1.0L, D1 = D(l + 1);
    long double al = sqrtl((2.0L * l + 3.0L) / (2.0L * l + 1.0L)) * (2.0L * l + 1.0L) * l1 / D1;
    long double g_next;
    if (l == l0) g_next = 1.0L;
    else {
      long double be = sqrtl((2.0L * l + 3.0L) / (2.0L * l - 1.0L)) * l1 * D(l) / (l * D1);
      g_next = be * g_prev;
    }
    long double A = al * g_cur / g_next;
    long double C = (long double)m * s / ((long double)l * l1);
    double *o = out + 4 * (size_t)(l - l0);
    o[0] = (double)A; o[1] = (double)(A * C); o[2] = (double)g_cur; o[3] = 0.0;
    g_prev = g_cur; g_cur = g_next;
  }
}

void build_coef_table(int lmax, int spin, const std::vector<int> &mval, std::vector<double> &tab,
                      std::vector<long long> &ofs) {
  const int nm = (int)mval.size();
  const int per = spin == 0 ? 2 : 4;
  ofs.assign(nm + 1, 0);
  for (int i = 0; i < nm; ++i) {
    int l0 = mval[i] > spin ? mval[i] : spin;
    long long n = lmax >= l0 ? (lmax - l0 + 1) : 0;
    ofs[i + 1] = ofs[i] + n * per;
  }
  // zero padding: the kernels stage whole tiles of up to 256 rows with cp.async and may run past
  // the last row of the last m
  tab.assign((size_t)ofs[nm] + 4 * 256 + 4, 0.0);
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if ((int)nt > nm) nt = nm > 0 ? nm : 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t]() {
      for (int i = (int)t; i < nm; i += (int)nt) {
        int m = mval[i];
        int l0 = m > spin ? m : spin;
        if (lmax < l0) continue;
        if (spin == 0) fill_spin0(lmax, m, tab.data() + ofs[i]);
        else fill_spins(lmax, m, spin, tab.data() + ofs[i]);
      }
    });
  }
  for (auto &t : th) t.join();
}

// K0[m] = (-1)^m sqrt((2m+1)/(4pi) prod_{k<=m} (2k-1)/(2k))
// K2[m] : m>=2: (-1)^m sqrt((2m+1)/4pi) sqrt(prod_{k<=m}(2k-1)/(2k)) sqrt((m-1)m/((m+1)(m+2))) * 4
//         m==0: sqrt(5/4pi) sqrt(6)/4 ; m==1: sqrt(5/4pi) * 2     (see legendre_core.cuh start_spin2)
void build_start_norms(int mmax, std::vector<double> &K0, std::vector<double> &K2) {
  K0.assign(mmax + 1, 0.0);
  K2.assign(mmax + 1, 0.0);
  const long double fourpi = 4.0L * acosl(-1.0L);
  long double prod = 1.0L;   // prod_{k<=m} (2k-1)/(2k)
  for (int m = 0; m <= mmax; ++m) {
    if (m > 0) prod *= (2.0L * m - 1.0L) / (2.0L * m);
    long double sg = (m & 1) ? -1.0L : 1.0L;
    K0[m] = (double)(sg * sqrtl((2.0L * m + 1.0L) / fourpi * prod));
    if (m >= 2) {
      long double r = ((long double)(m - 1) * m) / ((long double)(m + 1) * (m + 2));
      K2[m] = (double)(sg * sqrtl((2.0L * m + 1.0L) / fourpi * prod * r) * 4.0L);
    } else if (m == 0) {
      K2[m] = (double)(sqrtl(5.0L / fourpi) * sqrtl(6.0L) / 4.0L);
    } else {
      K2[m] = (double)(sqrtl(5.0L / fourpi) * 2.0L);
    }
  }
}

// Start-value normalisations for arbitrary spin s >= 1 (see legendre_core.cuh start_spin_s):
//   m >= s: (+-s)lambda_{mm} = Ks[m] sth^(m-s) {sh^(2s), ch^(2s)}
//           Ks[m] = (-1)^m sqrt((2m+1)/4pi) sqrt(prod_{k<=m}(2k-1)/(2k)) 2^s sqrt(prod_{i=1..s}(m-s+i)/(m+i))
//   m <  s: (+s)lambda_{sm} = Ks[m] ch^(s-m) sh^(s+m),  (-s)lambda_{sm} = (-1)^(s-m) Ks[m] ch^(s+m) sh^(s-m)
//           Ks[m] = (-1)^m sqrt((2s+1)/4pi) sqrt((2s)!/((s+m)!(s-m)!))
void build_start_norms_spin(int mmax, int spin, std::vector<double> &Ks) {
  Ks.assign(mmax + 1, 0.0);
  const long double fourpi = 4.0L * acosl(-1.0L);
  long double prod = 1.0L;   // prod_{k<=m} (2k-1)/(2k)
  for (int m = 0; m <= mmax; ++m) {
    if (m > 0) prod *= (2.0L * m - 1.0L) / (2.0L * m);
    long double sg = (m & 1) ? -1.0L : 1.0L;
    if (m >= spin) {
      long double r = 1.0L;
      for (int i = 1; i <= spin; ++i) r *= (long double)(m - spin + i) / (long double)(m + i);
      Ks[m] = (double)(sg * sqrtl((2.0L * m + 1.0L) / fourpi * prod * r) * powl(2.0L, spin));
    } else {
      long double b = 1.0L;    // (2s)! / ((s+m)! (s-m)!) = binomial(2s, s+m)
      for (int i = 1; i <= spin - m; ++i) b *= (long double)(spin + m + i) / (long double)i;
      Ks[m] = (double)(sg * sqrtl((2.0L * spin + 1.0L) / fourpi * b));
    }
  }
}

}  // namespace cmdr
