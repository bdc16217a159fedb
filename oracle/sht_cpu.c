/* oracle/sht_cpu.c -- CPU restatement of the SHT hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a checker and a timed CPU baseline.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it.  The product (commander_b200/) never links it.
 *
 * PARITY STATUS: "parity unpinned".  The reference's arithmetic for this path is
 * libsharp2 (HEALPix 3.70 bundle: cmake/project_instructions.cmake:102,106;
 * cmake/healpix.cmake:27), which is neither vendored under /root/reference nor
 * fetchable, and the reference has no tests/golden vectors for it.  This file
 * restates the published algorithm libsharp2 implements:
 *   - HEALPix ring geometry         (call site commander3/src/sharp.f90:145-166)
 *   - real-packed m-major alm       (commander3/src/sharp.f90:115-134,
 *                                    commander3/src/comm_map_mod.f90:228-261)
 *   - scaled three-term Legendre recurrence for spin 0
 *                                   (same recurrence as the reference's own
 *                                    commander3/src/math_tools.f90:989-1023)
 *   - Wigner-d three-term recurrence for spin 2 (HEALPix "COSMO" convention,
 *                                    commander3/src/comm_map_mod.f90:1002)
 *   - per-ring FFT with phi0 shift and aliasing fold
 *   - job types YtW/Y/Yt/WY         (commander3/src/sharp.f90:8-14)
 * and is pinned by tests/test_oracle.py against oracle/sht_def.py (dense
 * definition, itself pinned to scipy.special.sph_harm_y and the closed forms the
 * reference fixes) and against tests/golden/.
 *
 * Deliberately a different construction from the CUDA product: plain (unscaled-
 * coefficient) three-term recurrences, starting values through long-double
 * logarithms, libsharp-style 2^+-800 dynamic rescaling, own radix-2/Bluestein FFT.
 *
 * Build: see oracle/Makefile.  OpenMP over m (Legendre) and over rings (FFT).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SCL 800           /* rescaling granularity in bits (libsharp: 2^+-800) */
#define TBITS 100         /* values below 2^-TBITS are not accumulated */

enum { JOB_YTW = 0, JOB_Y = 1, JOB_YT = 2, JOB_WY = 3 };

/* ------------------------------------------------------------------ geometry */
typedef struct {
  int npairs;
  int *north;          /* north ring number of the pair (1..2nside) */
  double *cth, *sth;   /* of the northern ring */
  long double *l2sh, *l2ch; /* log2 sin(theta/2), log2 cos(theta/2) */
  int *nph;
  double *phi0, *wgt;
  int64_t *ofsN, *ofsS; /* local pixel offset of north/south ring, -1 if absent */
} pairs_t;

static void ring_geom(int nside, int ring, double *cth, double *sth, int *nph,
                      double *phi0, int *north_out) {
  int64_t npix = 12 * (int64_t)nside * nside;
  int north = ring > 2 * nside ? 4 * nside - ring : ring;
  if (north < nside) {
    *cth = 1.0 - (double)north * north / (3.0 * (double)nside * nside);
    *sth = sin(2.0 * asin(north / (sqrt(6.0) * nside)));
    *nph = 4 * north;
    *phi0 = M_PI / *nph;
  } else {
    *cth = (2.0 * nside - north) * 8.0 * nside / (double)npix;
    *sth = sqrt((1.0 - *cth) * (1.0 + *cth));
    *nph = 4 * nside;
    *phi0 = ((north - nside) & 1) ? 0.0 : M_PI / *nph;
  }
  *north_out = north;
}

static pairs_t *make_pairs(int nside, int nrings, const int *rings, const double *weight) {
  /* slot per possible north ring */
  int nn = 2 * nside;
  int64_t *oN = malloc(sizeof(int64_t) * (nn + 1)), *oS = malloc(sizeof(int64_t) * (nn + 1));
  for (int i = 0; i <= nn; ++i) oN[i] = oS[i] = -1;
  int64_t ofs = 0;
  for (int i = 0; i < nrings; ++i) {
    int ring = rings ? rings[i] : i + 1;
    double c, s, p0; int nph, north;
    ring_geom(nside, ring, &c, &s, &nph, &p0, &north);
    if (ring == north) oN[north] = ofs; else oS[north] = ofs;
    ofs += nph;
  }
  int np = 0;
  for (int i = 1; i <= nn; ++i) if (oN[i] >= 0 || oS[i] >= 0) ++np;
  pairs_t *P = calloc(1, sizeof(pairs_t));
  P->npairs = np;
  P->north = malloc(sizeof(int) * (np + 1));
  P->cth = malloc(sizeof(double) * (np + 1)); P->sth = malloc(sizeof(double) * (np + 1));
  P->l2sh = malloc(sizeof(long double) * (np + 1)); P->l2ch = malloc(sizeof(long double) * (np + 1));
  P->nph = malloc(sizeof(int) * (np + 1));
  P->phi0 = malloc(sizeof(double) * (np + 1)); P->wgt = malloc(sizeof(double) * (np + 1));
  P->ofsN = malloc(sizeof(int64_t) * (np + 1)); P->ofsS = malloc(sizeof(int64_t) * (np + 1));
  int k = 0;
  for (int i = 1; i <= nn; ++i) {
    if (oN[i] < 0 && oS[i] < 0) continue;
    int north;
    ring_geom(nside, i, &P->cth[k], &P->sth[k], &P->nph[k], &P->phi0[k], &north);
    long double th = (i < nside) ? 2.0L * asinl((long double)i / (sqrtl(6.0L) * nside))
                                 : acosl((long double)(2 * nside - i) * 8.0L * nside / (12.0L * nside * nside));
    P->l2sh[k] = log2l(sinl(0.5L * th));
    P->l2ch[k] = log2l(cosl(0.5L * th));
    P->north[k] = i;
    P->wgt[k] = 4.0 * M_PI / (12.0 * (double)nside * nside) * (weight ? weight[i - 1] : 1.0);
    P->ofsN[k] = oN[i]; P->ofsS[k] = oS[i];
    ++k;
  }
  free(oN); free(oS);
  return P;
}

static void free_pairs(pairs_t *P) {
  free(P->north); free(P->cth); free(P->sth); free(P->l2sh); free(P->l2ch); free(P->nph);
  free(P->phi0); free(P->wgt); free(P->ofsN); free(P->ofsS); free(P);
}

/* libsharp-style per-ring m cut-off: contributions with m above it are below
 * double precision and skipped (only when `mlim_skip` is requested). */
static int get_mlim(int lmax, int spin, double sth, double cth) {
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

/* ------------------------------------------------------------------ FFT */
typedef struct fftplan {
  int n, m;               /* n = length; m = pow2 work length (n if n is pow2) */
  double *tw;             /* twiddles for length m: cos, sin interleaved, m/2 entries */
  double *chirp;          /* n entries (re,im): exp(+i pi k^2 / n)  (Bluestein only) */
  double *fchirp;         /* m entries: FFT_m of conj-chirp filter for backward(+) sign */
  struct fftplan *next;
} fftplan;

static fftplan *g_plans = NULL;

static void fft_pow2(const fftplan *p, double *x, int sign) {
  /* in-place radix-2 DIT, length p->m, x interleaved re/im, sign=+1 => e^{+i..} */
  int n = p->m;
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double tr = x[2*i], ti = x[2*i+1]; x[2*i] = x[2*j]; x[2*i+1] = x[2*j+1]; x[2*j] = tr; x[2*j+1] = ti; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len) {
      for (int k = 0; k < half; ++k) {
        double wr = p->tw[2*(k*step)], wi = sign * p->tw[2*(k*step)+1];
        double *a = x + 2*(i+k), *b = x + 2*(i+k+half);
        double tr = b[0]*wr - b[1]*wi, ti = b[0]*wi + b[1]*wr;
        b[0] = a[0] - tr; b[1] = a[1] - ti; a[0] += tr; a[1] += ti;
      }
    }
  }
}

static fftplan *get_plan(int n) {
  fftplan *p;
  for (p = g_plans; p; p = p->next) if (p->n == n) return p;
  p = calloc(1, sizeof(fftplan));
  p->n = n;
  int pow2 = (n & (n - 1)) == 0;
  int m = n;
  if (!pow2) { m = 1; while (m < 2 * n - 1) m <<= 1; }
  p->m = m;
  p->tw = malloc(sizeof(double) * (m > 1 ? m : 2));
  for (int k = 0; k < m / 2; ++k) {
    long double a = 2.0L * M_PIl * k / m;
    p->tw[2*k] = (double)cosl(a); p->tw[2*k+1] = (double)sinl(a);
  }
  if (!pow2) {
    p->chirp = malloc(sizeof(double) * 2 * n);
    for (int k = 0; k < n; ++k) {
      int64_t k2 = ((int64_t)k * k) % (2 * (int64_t)n);
      long double a = M_PIl * k2 / n;
      p->chirp[2*k] = (double)cosl(a); p->chirp[2*k+1] = (double)sinl(a);
    }
    /* backward transform z_j = sum_k Z_k e^{+2 pi i jk/n} = c_j sum_k (Z_k c_k) conj(c_{j-k}) */
    p->fchirp = calloc(2 * m, sizeof(double));
    for (int k = 0; k < n; ++k) {
      p->fchirp[2*k] = p->chirp[2*k]; p->fchirp[2*k+1] = -p->chirp[2*k+1];
      if (k) { p->fchirp[2*(m-k)] = p->chirp[2*k]; p->fchirp[2*(m-k)+1] = -p->chirp[2*k+1]; }
    }
    fft_pow2(p, p->fchirp, -1);
  }
  p->next = g_plans; g_plans = p;
  return p;
}

/* complex DFT of length n, in place, x interleaved; sign=+1: sum x_k e^{+2pi i jk/n}.
 * work: at least 2*m doubles */
static void cfft(const fftplan *p, double *x, int sign, double *work) {
  int n = p->n, m = p->m;
  if (n == m) { fft_pow2(p, x, sign); return; }
  /* Bluestein.  For sign=-1 use conj(x) trick: DFT^-(x) = conj(DFT^+(conj x)) */
  if (sign < 0) for (int k = 0; k < n; ++k) x[2*k+1] = -x[2*k+1];
  for (int k = 0; k < n; ++k) {
    double cr = p->chirp[2*k], ci = p->chirp[2*k+1];
    work[2*k] = x[2*k]*cr - x[2*k+1]*ci; work[2*k+1] = x[2*k]*ci + x[2*k+1]*cr;
  }
  memset(work + 2*n, 0, sizeof(double) * 2 * (m - n));
  fft_pow2(p, work, -1);
  for (int k = 0; k < m; ++k) {
    double ar = work[2*k], ai = work[2*k+1], br = p->fchirp[2*k], bi = p->fchirp[2*k+1];
    work[2*k] = ar*br - ai*bi; work[2*k+1] = ar*bi + ai*br;
  }
  fft_pow2(p, work, +1);
  double inv = 1.0 / m;
  for (int k = 0; k < n; ++k) {
    double cr = p->chirp[2*k], ci = p->chirp[2*k+1];
    double ar = work[2*k]*inv, ai = work[2*k+1]*inv;
    x[2*k] = ar*cr - ai*ci; x[2*k+1] = ar*ci + ai*cr;
  }
  if (sign < 0) for (int k = 0; k < n; ++k) x[2*k+1] = -x[2*k+1];
}

/* ------------------------------------------------------------------ Legendre
 * All kernels handle one m and a block of up to NV ring pairs.  Values are kept
 * as lam * 2^(SCL*scale), scale<=0; accumulation starts when scale reaches 0. */

static const double BIG = 0x1p+700, SMALL = 0x1p-800; /* rescale when |lam|>=2^(SCL-TBITS) */

/* log2 of sqrt( prod_{k=1..m} (2k-1)/(2k) ) tables */
static long double *g_l2half = NULL; static int g_l2half_n = 0;
static void ensure_l2half(int mmax) {
  if (g_l2half_n > mmax) return;
  free(g_l2half);
  g_l2half = malloc(sizeof(long double) * (mmax + 2));
  long double acc = 0;
  g_l2half[0] = 0;
  for (int k = 1; k <= mmax + 1; ++k) { acc += log2l(1.0L - 1.0L / (2.0L * k)); g_l2half[k] = 0.5L * acc; }
  g_l2half_n = mmax + 1;
}

static void split_scaled(long double lg2, int sign, double *val, int *scale) {
  /* value = sign * 2^lg2  ->  val * 2^(SCL*scale) with val in [2^-TBITS, 2^(SCL-TBITS)) or scale=0 */
  long double q = floorl((lg2 + TBITS) / SCL);
  int sc = q >= 0 ? 0 : (int)q;
  *scale = sc;
  *val = sign * (double)exp2l(lg2 - (long double)SCL * sc);
}

/* spin-0 start: lambda_mm(theta) = (-1)^m sqrt((2m+1)/4pi * (2m-1)!!/(2m)!!) sin^m(theta)
 * sin(theta) = 2 sin(theta/2) cos(theta/2) */
static void start0(int m, long double l2sh, long double l2ch, double *val, int *scale) {
  long double lg = 0.5L * log2l((2.0L * m + 1.0L) / (4.0L * M_PIl)) + g_l2half[m]
                 + m * (1.0L + l2sh + l2ch);
  split_scaled(lg, (m & 1) ? -1 : 1, val, scale);
}

/* spin-s start at l0=max(m,s) for P = +s lambda_{l0,m}, M = -s lambda_{l0,m}
 *   m>=s: d^m_{-m,s'} = sqrt((2m)!/((m+s')!(m-s')!)) cos^{m-s'} sin^{m+s'}   (half angles)
 *   m< s: d^s_{-m,+s} = sqrt((2s)!/((s+m)!(s-m)!)) cos^{s-m} sin^{s+m}
 *         d^s_{-m,-s} = (-1)^{s-m} (same) cos^{s+m} sin^{s-m}
 * times (-1)^m sqrt((2 l0+1)/4pi). */
static long double l2binom_sqrt(int n2, int k) {
  /* 0.5*log2( n2! / (k! (n2-k)!) ) */
  return 0.5L * (lgammal(n2 + 1.0L) - lgammal(k + 1.0L) - lgammal(n2 - k + 1.0L)) / logl(2.0L);
}
static void starts(int m, int s, long double l2sh, long double l2ch,
                   double *vp, double *vm, int *scale) {
  int l0 = m > s ? m : s;
  long double nrm = 0.5L * log2l((2.0L * l0 + 1.0L) / (4.0L * M_PIl));
  long double lgp, lgm; int sgp = (m & 1) ? -1 : 1, sgm = sgp;
  if (m >= s) {
    /* sqrt((2m)!/((m+s)!(m-s)!)) = 2^m sqrt(prod (2k-1)/(2k)) sqrt(prod_{i=1..s} (m-s+i)/(m+i)) */
    long double c = m + g_l2half[m];
    long double r = 0;
    for (int i = 1; i <= s; ++i) r += log2l((long double)(m - s + i) / (long double)(m + i));
    c += 0.5L * r;
    lgp = nrm + c + (m - s) * l2ch + (m + s) * l2sh;
    lgm = nrm + c + (m + s) * l2ch + (m - s) * l2sh;
  } else {
    long double c = l2binom_sqrt(2 * s, s + m);
    lgp = nrm + c + (s - m) * l2ch + (s + m) * l2sh;
    lgm = nrm + c + (s + m) * l2ch + (s - m) * l2sh;
    if ((s - m) & 1) sgm = -sgm;
  }
  /* common scale from the larger of the two */
  long double lgmax = lgp > lgm ? lgp : lgm;
  long double q = floorl((lgmax + TBITS) / SCL);
  int sc = q >= 0 ? 0 : (int)q;
  *scale = sc;
  *vp = sgp * (double)exp2l(lgp - (long double)SCL * sc);
  *vm = sgm * (double)exp2l(lgm - (long double)SCL * sc);
}

/* spin-0 coefficient tables for one m: lam_{l+1} = x A[l] lam_l - B[l] lam_{l-1} */
static void coef0(int lmax, int m, double *A, double *B) {
  double prev = 0;
  for (int l = m; l <= lmax; ++l) {
    double fl2 = (double)(l + 1) * (l + 1);
    double a = sqrt((4.0 * fl2 - 1.0) / (fl2 - (double)m * m));
    A[l] = a; B[l] = (l == m) ? 0.0 : a / prev; prev = a;
  }
}
/* spin-s: L_{l+1} = A[l] (x -+ C[l]) L_l - B[l] L_{l-1};  P uses x + C, M uses x - C */
static void coefs(int lmax, int m, int s, double *A, double *B, double *C) {
  int l0 = m > s ? m : s;
  for (int l = l0; l <= lmax; ++l) {
    double l1 = l + 1.0;
    double D1 = sqrt((l1 * l1 - (double)m * m) * (l1 * l1 - (double)s * s));
    double D0 = sqrt(((double)l * l - (double)m * m) * ((double)l * l - (double)s * s));
    A[l] = sqrt((2.0 * l + 3.0) / (2.0 * l + 1.0)) * (2.0 * l + 1.0) * l1 / D1;
    B[l] = (l == l0) ? 0.0 : sqrt((2.0 * l + 3.0) / (2.0 * l - 1.0)) * l1 * D0 / (l * D1);
    C[l] = (double)m * s / ((double)l * l1);
  }
}

/* Vector blocks: NV0 ring pairs per block for spin 0, NV2 for spin 2 (several SIMD
 * registers per array so that the dependent recurrence chains overlap). */
#define NV0 32
#define NV2 16
#define NVMAX 32

/* --- spin 0 synthesis: outN = sum_l a_l lam_l ; outS = sum_l (-1)^(l+m) a_l lam_l */
static void leg_synth0(int lmax, int m, const double *A, const double *B,
                       const double *ar, const double *ai, int nb,
                       const double *cth, const double *lam0, const int *scale0,
                       double *oNr, double *oNi, double *oSr, double *oSi) {
  double x[NV0], l1[NV0], l2[NV0], per[NV0], pei[NV0], por[NV0], poi[NV0], cf[NV0];
  int sc[NV0];
  for (int v = 0; v < NV0; ++v) {
    int u = v < nb ? v : nb - 1;
    x[v] = cth[u]; l2[v] = lam0[u]; l1[v] = 0; sc[v] = scale0[u];
    per[v] = pei[v] = por[v] = poi[v] = 0;
  }
  int l = m, nact = 0;
  for (int v = 0; v < NV0; ++v) nact += (sc[v] == 0);
  /* careful phase: some rings still carry a scale factor */
  while (l <= lmax && nact < NV0) {
    double a_r = ar[l], a_i = ai[l], Al = A[l], Bl = B[l], mx = 0;
    if (nact > 0) {
      for (int v = 0; v < NV0; ++v) cf[v] = sc[v] == 0 ? l2[v] : 0.0;
      if ((l - m) & 1) {
#pragma omp simd
        for (int v = 0; v < NV0; ++v) { por[v] += cf[v] * a_r; poi[v] += cf[v] * a_i; }
      } else {
#pragma omp simd
        for (int v = 0; v < NV0; ++v) { per[v] += cf[v] * a_r; pei[v] += cf[v] * a_i; }
      }
    }
#pragma omp simd reduction(max:mx)
    for (int v = 0; v < NV0; ++v) {
      double t = x[v] * Al * l2[v] - Bl * l1[v];
      l1[v] = l2[v]; l2[v] = t;
      mx = fmax(mx, fabs(t));
    }
    if (mx >= BIG)
      for (int v = 0; v < NV0; ++v)
        if (sc[v] < 0 && fabs(l2[v]) >= BIG) { l1[v] *= SMALL; l2[v] *= SMALL; if (++sc[v] == 0) ++nact; }
    ++l;
  }
  /* fast phase, all active */
  for (; l <= lmax; ++l) {
    double a_r = ar[l], a_i = ai[l], Al = A[l], Bl = B[l];
    if ((l - m) & 1) {
#pragma omp simd
      for (int v = 0; v < NV0; ++v) { por[v] += l2[v] * a_r; poi[v] += l2[v] * a_i; }
    } else {
#pragma omp simd
      for (int v = 0; v < NV0; ++v) { per[v] += l2[v] * a_r; pei[v] += l2[v] * a_i; }
    }
#pragma omp simd
    for (int v = 0; v < NV0; ++v) {
      double t = x[v] * Al * l2[v] - Bl * l1[v];
      l1[v] = l2[v]; l2[v] = t;
    }
  }
  for (int v = 0; v < nb; ++v) {
    oNr[v] = per[v] + por[v]; oNi[v] = pei[v] + poi[v];
    oSr[v] = per[v] - por[v]; oSi[v] = pei[v] - poi[v];
  }
}

/* --- spin 0 analysis: a_l += sum_rings lam_l * (l+m even ? sN+sS : sN-sS) */
static void leg_anal0(int lmax, int m, const double *A, const double *B,
                      double *ar, double *ai, int nb,
                      const double *cth, const double *lam0, const int *scale0,
                      const double *qNr, const double *qNi, const double *qSr, const double *qSi) {
  double x[NV0], l1[NV0], l2[NV0], er[NV0], ei[NV0], odr[NV0], odi[NV0], cf[NV0];
  int sc[NV0];
  for (int v = 0; v < NV0; ++v) {
    int u = v < nb ? v : nb - 1;
    double z = v < nb ? 1.0 : 0.0;
    x[v] = cth[u]; l2[v] = lam0[u]; l1[v] = 0; sc[v] = scale0[u];
    er[v] = z * (qNr[u] + qSr[u]); ei[v] = z * (qNi[u] + qSi[u]);
    odr[v] = z * (qNr[u] - qSr[u]); odi[v] = z * (qNi[u] - qSi[u]);
  }
  int l = m, nact = 0;
  for (int v = 0; v < NV0; ++v) nact += (sc[v] == 0);
  while (l <= lmax && nact < NV0) {
    double Al = A[l], Bl = B[l], mx = 0;
    if (nact > 0) {
      double sr = 0, si = 0;
      const double *pr = ((l - m) & 1) ? odr : er, *pi = ((l - m) & 1) ? odi : ei;
      for (int v = 0; v < NV0; ++v) cf[v] = sc[v] == 0 ? l2[v] : 0.0;
#pragma omp simd reduction(+:sr,si)
      for (int v = 0; v < NV0; ++v) { sr += cf[v] * pr[v]; si += cf[v] * pi[v]; }
      ar[l] += sr; ai[l] += si;
    }
#pragma omp simd reduction(max:mx)
    for (int v = 0; v < NV0; ++v) {
      double t = x[v] * Al * l2[v] - Bl * l1[v];
      l1[v] = l2[v]; l2[v] = t;
      mx = fmax(mx, fabs(t));
    }
    if (mx >= BIG)
      for (int v = 0; v < NV0; ++v)
        if (sc[v] < 0 && fabs(l2[v]) >= BIG) { l1[v] *= SMALL; l2[v] *= SMALL; if (++sc[v] == 0) ++nact; }
    ++l;
  }
  for (; l <= lmax; ++l) {
    const double *pr = ((l - m) & 1) ? odr : er, *pi = ((l - m) & 1) ? odi : ei;
    double sr = 0, si = 0, Al = A[l], Bl = B[l];
#pragma omp simd reduction(+:sr,si)
    for (int v = 0; v < NV0; ++v) { sr += l2[v] * pr[v]; si += l2[v] * pi[v]; }
    ar[l] += sr; ai[l] += si;
#pragma omp simd
    for (int v = 0; v < NV0; ++v) {
      double t = x[v] * Al * l2[v] - Bl * l1[v];
      l1[v] = l2[v]; l2[v] = t;
    }
  }
}

/* --- spin s synthesis.  cp_l = -(E+iB)_l , cm_l = -(E-iB)_l  (complex, given re/im)
 *  A1 = sum cp P, A2 = sum cm M (north);  A3 = sum sg cp M, A4 = sum sg cm P (south),
 *  sg_l = (-1)^(l+m).  Output Q = (A1+A2)/2, U = -i (A1-A2)/2. */
static void leg_synths(int lmax, int m, int s, const double *A, const double *B, const double *C,
                       const double *cpr, const double *cpi, const double *cmr, const double *cmi,
                       int nb, const double *cth, const double *P0, const double *M0, const int *scale0,
                       double *QNr, double *QNi, double *UNr, double *UNi,
                       double *QSr, double *QSi, double *USr, double *USi) {
  double x[NV2], p1[NV2], p2[NV2], m1[NV2], m2[NV2], cf[NV2];
  double a1r[NV2], a1i[NV2], a2r[NV2], a2i[NV2], a3r[NV2], a3i[NV2], a4r[NV2], a4i[NV2];
  int sc[NV2];
  int l0 = m > s ? m : s;
  for (int v = 0; v < NV2; ++v) {
    int u = v < nb ? v : nb - 1;
    x[v] = cth[u]; p2[v] = P0[u]; m2[v] = M0[u]; p1[v] = m1[v] = 0; sc[v] = scale0[u];
    a1r[v] = a1i[v] = a2r[v] = a2i[v] = a3r[v] = a3i[v] = a4r[v] = a4i[v] = 0;
  }
  int nact = 0;
  for (int v = 0; v < NV2; ++v) nact += (sc[v] == 0);
  for (int l = l0; l <= lmax; ++l) {
    const double sg = ((l + m) & 1) ? -1.0 : 1.0;   /* (+s)lambda(pi-theta) = (-1)^(l+m) (-s)lambda(theta), any s */
    const double Al = A[l], Bl = B[l], Cl = C[l];
    const double c_pr = cpr[l], c_pi = cpi[l], c_mr = cmr[l], c_mi = cmi[l];
    if (nact == NV2) {
#pragma omp simd
      for (int v = 0; v < NV2; ++v) {
        double P = p2[v], M = m2[v], sP = sg * P, sM = sg * M;
        a1r[v] += c_pr * P;  a1i[v] += c_pi * P;
        a2r[v] += c_mr * M;  a2i[v] += c_mi * M;
        a3r[v] += c_pr * sM; a3i[v] += c_pi * sM;
        a4r[v] += c_mr * sP; a4i[v] += c_mi * sP;
        double tp = Al * (x[v] + Cl) * P - Bl * p1[v];
        double tm = Al * (x[v] - Cl) * M - Bl * m1[v];
        p1[v] = P; p2[v] = tp; m1[v] = M; m2[v] = tm;
      }
    } else {
      double mx = 0;
      if (nact > 0) {
        for (int v = 0; v < NV2; ++v) cf[v] = sc[v] == 0 ? 1.0 : 0.0;
#pragma omp simd
        for (int v = 0; v < NV2; ++v) {
          double P = cf[v] * p2[v], M = cf[v] * m2[v], sP = sg * P, sM = sg * M;
          a1r[v] += c_pr * P;  a1i[v] += c_pi * P;
          a2r[v] += c_mr * M;  a2i[v] += c_mi * M;
          a3r[v] += c_pr * sM; a3i[v] += c_pi * sM;
          a4r[v] += c_mr * sP; a4i[v] += c_mi * sP;
        }
      }
#pragma omp simd reduction(max:mx)
      for (int v = 0; v < NV2; ++v) {
        double P = p2[v], M = m2[v];
        double tp = Al * (x[v] + Cl) * P - Bl * p1[v];
        double tm = Al * (x[v] - Cl) * M - Bl * m1[v];
        p1[v] = P; p2[v] = tp; m1[v] = M; m2[v] = tm;
        mx = fmax(mx, fmax(fabs(tp), fabs(tm)));
      }
      if (mx >= BIG)
        for (int v = 0; v < NV2; ++v)
          if (sc[v] < 0 && (fabs(p2[v]) >= BIG || fabs(m2[v]) >= BIG)) {
            p1[v] *= SMALL; p2[v] *= SMALL; m1[v] *= SMALL; m2[v] *= SMALL;
            if (++sc[v] == 0) ++nact;
          }
    }
  }
  for (int v = 0; v < nb; ++v) {
    QNr[v] = 0.5 * (a1r[v] + a2r[v]); QNi[v] = 0.5 * (a1i[v] + a2i[v]);
    UNr[v] = 0.5 * (a1i[v] - a2i[v]); UNi[v] = -0.5 * (a1r[v] - a2r[v]);
    QSr[v] = 0.5 * (a3r[v] + a4r[v]); QSi[v] = 0.5 * (a3i[v] + a4i[v]);
    USr[v] = 0.5 * (a3i[v] - a4i[v]); USi[v] = -0.5 * (a3r[v] - a4r[v]);
  }
}

/* --- spin s analysis.  zp = qQ + i qU, zm = qQ - i qU per ring.
 *  S1_l = sum_r P zpN + sg M zpS ; S2_l = sum_r M zmN + sg P zmS
 *  even spin: E_l = -(S1+S2)/2 ; B_l = (i/2)(S1-S2) */
static void leg_anals(int lmax, int m, int s, const double *A, const double *B, const double *C,
                      double *Er, double *Ei, double *Br, double *Bi,
                      int nb, const double *cth, const double *P0, const double *M0, const int *scale0,
                      const double *QNr, const double *QNi, const double *UNr, const double *UNi,
                      const double *QSr, const double *QSi, const double *USr, const double *USi) {
  double x[NV2], p1[NV2], p2[NV2], m1[NV2], m2[NV2], cf[NV2];
  double zpNr[NV2], zpNi[NV2], zmNr[NV2], zmNi[NV2], zpSr[NV2], zpSi[NV2], zmSr[NV2], zmSi[NV2];
  int sc[NV2];
  int l0 = m > s ? m : s;
  for (int v = 0; v < NV2; ++v) {
    int u = v < nb ? v : nb - 1;
    double z = v < nb ? 1.0 : 0.0;
    x[v] = cth[u]; p2[v] = P0[u]; m2[v] = M0[u]; p1[v] = m1[v] = 0; sc[v] = scale0[u];
    zpNr[v] = z * (QNr[u] - UNi[u]); zpNi[v] = z * (QNi[u] + UNr[u]);
    zmNr[v] = z * (QNr[u] + UNi[u]); zmNi[v] = z * (QNi[u] - UNr[u]);
    zpSr[v] = z * (QSr[u] - USi[u]); zpSi[v] = z * (QSi[u] + USr[u]);
    zmSr[v] = z * (QSr[u] + USi[u]); zmSi[v] = z * (QSi[u] - USr[u]);
  }
  int nact = 0;
  for (int v = 0; v < NV2; ++v) nact += (sc[v] == 0);
  for (int l = l0; l <= lmax; ++l) {
    const double sg = ((l + m) & 1) ? -1.0 : 1.0;   /* (+s)lambda(pi-theta) = (-1)^(l+m) (-s)lambda(theta), any s */
    const double Al = A[l], Bl = B[l], Cl = C[l];
    double s1r = 0, s1i = 0, s2r = 0, s2i = 0;
    if (nact == NV2) {
#pragma omp simd reduction(+:s1r,s1i,s2r,s2i)
      for (int v = 0; v < NV2; ++v) {
        double P = p2[v], M = m2[v], sP = sg * P, sM = sg * M;
        s1r += P * zpNr[v] + sM * zpSr[v]; s1i += P * zpNi[v] + sM * zpSi[v];
        s2r += M * zmNr[v] + sP * zmSr[v]; s2i += M * zmNi[v] + sP * zmSi[v];
        double tp = Al * (x[v] + Cl) * P - Bl * p1[v];
        double tm = Al * (x[v] - Cl) * M - Bl * m1[v];
        p1[v] = P; p2[v] = tp; m1[v] = M; m2[v] = tm;
      }
    } else {
      double mx = 0;
      if (nact > 0) {
        for (int v = 0; v < NV2; ++v) cf[v] = sc[v] == 0 ? 1.0 : 0.0;
#pragma omp simd reduction(+:s1r,s1i,s2r,s2i)
        for (int v = 0; v < NV2; ++v) {
          double P = cf[v] * p2[v], M = cf[v] * m2[v], sP = sg * P, sM = sg * M;
          s1r += P * zpNr[v] + sM * zpSr[v]; s1i += P * zpNi[v] + sM * zpSi[v];
          s2r += M * zmNr[v] + sP * zmSr[v]; s2i += M * zmNi[v] + sP * zmSi[v];
        }
      }
#pragma omp simd reduction(max:mx)
      for (int v = 0; v < NV2; ++v) {
        double P = p2[v], M = m2[v];
        double tp = Al * (x[v] + Cl) * P - Bl * p1[v];
        double tm = Al * (x[v] - Cl) * M - Bl * m1[v];
        p1[v] = P; p2[v] = tp; m1[v] = M; m2[v] = tm;
        mx = fmax(mx, fmax(fabs(tp), fabs(tm)));
      }
      if (mx >= BIG)
        for (int v = 0; v < NV2; ++v)
          if (sc[v] < 0 && (fabs(p2[v]) >= BIG || fabs(m2[v]) >= BIG)) {
            p1[v] *= SMALL; p2[v] *= SMALL; m1[v] *= SMALL; m2[v] *= SMALL;
            if (++sc[v] == 0) ++nact;
          }
    }
    /* adjoint of (+s)a = sp (E + iB), (-s)a = -(E - iB):  E = (sp S1 - S2)/2,  B = (i/2)(-S2 - sp S1) */
    const double sp = (s & 1) ? 1.0 : -1.0;
    Er[l] += 0.5 * (sp * s1r - s2r); Ei[l] += 0.5 * (sp * s1i - s2i);
    Br[l] += 0.5 * (sp * s1i + s2i); Bi[l] += 0.5 * (-s2r - sp * s1r);
  }
}

/* ------------------------------------------------------------------ driver */
static int g_mlim_skip = 0;
static double g_t_leg = 0, g_t_fft = 0;   /* wall seconds of the last osht_execute: Legendre, FFT */
void osht_last_times(double *t2) { t2[0] = g_t_leg; t2[1] = g_t_fft; }
static double wall(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}
void osht_set_mlim_skip(int on) { g_mlim_skip = on; }
int osht_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int64_t osht_map_size(int nside, int nrings, const int *rings) {
  int64_t n = 0;
  for (int i = 0; i < nrings; ++i) {
    int ring = rings ? rings[i] : i + 1;
    int north = ring > 2 * nside ? 4 * nside - ring : ring;
    n += north < nside ? 4 * north : 4 * nside;
  }
  return n;
}
int64_t osht_alm_count(int lmax, int nm, const int *ms) {
  int64_t n = 0;
  for (int i = 0; i < nm; ++i) { int m = ms ? ms[i] : i; n += (int64_t)(lmax + 1 - m) * (m == 0 ? 1 : 2); }
  return n;
}

/* phase buffer index: ph[((c*npairs + p)*2 + h)*nm + im] complex (2 doubles) */
#define PH(c, p, h, im) (2 * ((((int64_t)(c) * npairs + (p)) * 2 + (h)) * nm + (im)))

int osht_execute(int type, int spin, int nside, int lmax,
                 int nrings, const int *rings, const double *weight,
                 int nm, const int *ms_in, double **alm, double **map, int add, int nthreads) {
  if (spin < 0 || spin > 32) return -1;
  /* overall sign of the spin-s combination: -1 for even s (HEALPix COSMO convention at s = 2), +1 for odd s
   * (libsharp2's sharp_Ylmgen_get_norm as remembered; unpinned for s != 2) */
  const double ssg = (spin & 1) ? 1.0 : -1.0;
  if (type < 0 || type > 3) return -2;
  const int ncomp = spin == 0 ? 1 : 2;
  const int synth = (type == JOB_Y || type == JOB_WY);
  pairs_t *P = make_pairs(nside, nrings, rings, weight);
  const int npairs = P->npairs;
  int *ms = malloc(sizeof(int) * (nm + 1));
  int64_t *mstart = malloc(sizeof(int64_t) * (nm + 1));
  int64_t idx = 0; int mmax = 0;
  for (int i = 0; i < nm; ++i) {
    ms[i] = ms_in ? ms_in[i] : i; mstart[i] = idx;
    idx += (int64_t)(lmax + 1 - ms[i]) * (ms[i] == 0 ? 1 : 2);
    if (ms[i] > mmax) mmax = ms[i];
  }
  ensure_l2half(mmax > spin ? mmax : spin);
  int64_t nph_tot = (int64_t)ncomp * npairs * 2 * nm;
  double *ph = calloc(2 * (nph_tot > 0 ? nph_tot : 1), sizeof(double));
  /* FFT plans (serial creation, then read-only) */
  int maxm = 4;
  for (int p = 0; p < npairs; ++p) { fftplan *pl = get_plan(P->nph[p]); if (pl->m > maxm) maxm = pl->m; }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const double rt2 = sqrt(2.0), irt2 = 1.0 / rt2;

  double t_a = wall();
  /* ---------------- analysis: rings -> phases */
  if (!synth) {
#pragma omp parallel
    {
      double *z = malloc(sizeof(double) * 2 * maxm), *work = malloc(sizeof(double) * 2 * maxm);
#pragma omp for schedule(dynamic, 4)
      for (int p = 0; p < npairs; ++p) {
        int n = P->nph[p]; fftplan *pl = get_plan(n);
        double w = (type == JOB_YTW) ? P->wgt[p] : 1.0;
        for (int c = 0; c < ncomp; ++c) {
          const double *mp = map[c];
          for (int j = 0; j < n; ++j) {
            z[2*j]   = P->ofsN[p] >= 0 ? w * mp[P->ofsN[p] + j] : 0.0;
            z[2*j+1] = P->ofsS[p] >= 0 ? w * mp[P->ofsS[p] + j] : 0.0;
          }
          cfft(pl, z, -1, work);   /* Z_k = sum_j (xN+i xS)_j e^{-2 pi i jk/n} */
          for (int im = 0; im < nm; ++im) {
            int m = ms[im]; int k = m % n, k2 = (n - k) % n;
            /* XN_k = (Z_k + conj Z_{-k})/2 ; XS_k = (Z_k - conj Z_{-k})/(2i) */
            double zr = z[2*k], zi = z[2*k+1], yr = z[2*k2], yi = -z[2*k2+1];
            double xnr = 0.5 * (zr + yr), xni = 0.5 * (zi + yi);
            double xsr = 0.5 * (zi - yi), xsi = -0.5 * (zr - yr);
            double cm = m == 0 ? 1.0 : 2.0;
            double ang = fmod((double)m * P->phi0[p], 2.0 * M_PI);
            double cr = cm * cos(ang), ci = -cm * sin(ang);   /* c_m e^{-i m phi0} */
            ph[PH(c,p,0,im)]   = xnr * cr - xni * ci; ph[PH(c,p,0,im)+1] = xnr * ci + xni * cr;
            ph[PH(c,p,1,im)]   = xsr * cr - xsi * ci; ph[PH(c,p,1,im)+1] = xsr * ci + xsi * cr;
          }
        }
      }
      free(z); free(work);
    }
  }

  double t_b = wall();
  /* ---------------- Legendre over m */
#pragma omp parallel
  {
    double *A = malloc(sizeof(double) * (lmax + 2)), *B = malloc(sizeof(double) * (lmax + 2)),
           *C = malloc(sizeof(double) * (lmax + 2));
    double *c0 = calloc(4 * (size_t)(lmax + 2), sizeof(double));
    double *c1 = c0 + (lmax + 2), *c2 = c1 + (lmax + 2), *c3 = c2 + (lmax + 2);
    double *lam0 = malloc(sizeof(double) * (npairs + 1)), *lamM = malloc(sizeof(double) * (npairs + 1));
    int *sc0 = malloc(sizeof(int) * (npairs + 1));
    double bufs[16][NVMAX];
#pragma omp for schedule(dynamic, 1)
    for (int im = 0; im < nm; ++im) {
      int m = ms[im];
      int l0 = m > spin ? m : spin;
      double nrm = m == 0 ? 1.0 : irt2;   /* real-packed -> complex */
      /* first pair that survives the m cut-off (pairs are sorted by theta) */
      int pfirst = 0;
      if (g_mlim_skip) while (pfirst < npairs && m > get_mlim(lmax, spin, P->sth[pfirst], P->cth[pfirst])) ++pfirst;
      if (spin == 0) {
        coef0(lmax, m, A, B);
        for (int p = pfirst; p < npairs; ++p) start0(m, P->l2sh[p], P->l2ch[p], &lam0[p], &sc0[p]);
        if (synth) {
          const double *a = alm[0] + mstart[im];
          for (int l = m; l <= lmax; ++l) {
            if (m == 0) { c0[l] = a[l]; c1[l] = 0; }
            else { c0[l] = nrm * a[2*(l-m)]; c1[l] = nrm * a[2*(l-m)+1]; }
          }
          for (int p = pfirst; p < npairs; p += NV0) {
            int nb = npairs - p < NV0 ? npairs - p : NV0;
            leg_synth0(lmax, m, A, B, c0, c1, nb, P->cth + p, lam0 + p, sc0 + p,
                       bufs[0], bufs[1], bufs[2], bufs[3]);
            for (int v = 0; v < nb; ++v) {
              ph[PH(0,p+v,0,im)] = bufs[0][v]; ph[PH(0,p+v,0,im)+1] = bufs[1][v];
              ph[PH(0,p+v,1,im)] = bufs[2][v]; ph[PH(0,p+v,1,im)+1] = bufs[3][v];
            }
          }
        } else {
          for (int l = m; l <= lmax; ++l) c0[l] = c1[l] = 0;
          for (int p = pfirst; p < npairs; p += NV0) {
            int nb = npairs - p < NV0 ? npairs - p : NV0;
            for (int v = 0; v < nb; ++v) {
              bufs[0][v] = ph[PH(0,p+v,0,im)]; bufs[1][v] = ph[PH(0,p+v,0,im)+1];
              bufs[2][v] = ph[PH(0,p+v,1,im)]; bufs[3][v] = ph[PH(0,p+v,1,im)+1];
            }
            leg_anal0(lmax, m, A, B, c0, c1, nb, P->cth + p, lam0 + p, sc0 + p,
                      bufs[0], bufs[1], bufs[2], bufs[3]);
          }
          double *a = alm[0] + mstart[im];
          for (int l = m; l <= lmax; ++l) {
            if (m == 0) a[l] = (add ? a[l] : 0.0) + c0[l];
            else {
              a[2*(l-m)]   = (add ? a[2*(l-m)]   : 0.0) + nrm * c0[l];
              a[2*(l-m)+1] = (add ? a[2*(l-m)+1] : 0.0) + nrm * c1[l];
            }
          }
        }
      } else {
        if (l0 <= lmax) coefs(lmax, m, spin, A, B, C);
        for (int p = pfirst; p < npairs; ++p) starts(m, spin, P->l2sh[p], P->l2ch[p], &lam0[p], &lamM[p], &sc0[p]);
        if (synth) {
          const double *aE = alm[0] + mstart[im], *aB = alm[1] + mstart[im];
          for (int l = l0; l <= lmax; ++l) {
            double er, ei, br, bi;
            if (m == 0) { er = aE[l]; ei = 0; br = aB[l]; bi = 0; }
            else { er = nrm * aE[2*(l-m)]; ei = nrm * aE[2*(l-m)+1]; br = nrm * aB[2*(l-m)]; bi = nrm * aB[2*(l-m)+1]; }
            c0[l] = ssg * (er - bi); c1[l] = ssg * (ei + br);   /* (+s)a = ssg (E + iB), ssg = -1 for even spin */
            c2[l] = -(er + bi); c3[l] = -(ei - br);             /* (-s)a = ssg (-1)^s (E - iB) = -(E - iB) */
          }
          for (int p = 0; p < pfirst; ++p)
            for (int c = 0; c < 2; ++c) for (int h = 0; h < 2; ++h) ph[PH(c,p,h,im)] = ph[PH(c,p,h,im)+1] = 0;
          for (int p = pfirst; p < npairs; p += NV2) {
            int nb = npairs - p < NV2 ? npairs - p : NV2;
            if (l0 > lmax) { for (int k = 0; k < 8; ++k) for (int v = 0; v < NVMAX; ++v) bufs[k][v] = 0; }
            else leg_synths(lmax, m, spin, A, B, C, c0, c1, c2, c3, nb, P->cth + p, lam0 + p, lamM + p, sc0 + p,
                            bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5], bufs[6], bufs[7]);
            for (int v = 0; v < nb; ++v) {
              ph[PH(0,p+v,0,im)] = bufs[0][v]; ph[PH(0,p+v,0,im)+1] = bufs[1][v];
              ph[PH(1,p+v,0,im)] = bufs[2][v]; ph[PH(1,p+v,0,im)+1] = bufs[3][v];
              ph[PH(0,p+v,1,im)] = bufs[4][v]; ph[PH(0,p+v,1,im)+1] = bufs[5][v];
              ph[PH(1,p+v,1,im)] = bufs[6][v]; ph[PH(1,p+v,1,im)+1] = bufs[7][v];
            }
          }
        } else {
          for (int l = 0; l <= lmax; ++l) c0[l] = c1[l] = c2[l] = c3[l] = 0;
          if (l0 <= lmax)
          for (int p = pfirst; p < npairs; p += NV2) {
            int nb = npairs - p < NV2 ? npairs - p : NV2;
            for (int v = 0; v < nb; ++v) {
              bufs[0][v] = ph[PH(0,p+v,0,im)]; bufs[1][v] = ph[PH(0,p+v,0,im)+1];
              bufs[2][v] = ph[PH(1,p+v,0,im)]; bufs[3][v] = ph[PH(1,p+v,0,im)+1];
              bufs[4][v] = ph[PH(0,p+v,1,im)]; bufs[5][v] = ph[PH(0,p+v,1,im)+1];
              bufs[6][v] = ph[PH(1,p+v,1,im)]; bufs[7][v] = ph[PH(1,p+v,1,im)+1];
            }
            leg_anals(lmax, m, spin, A, B, C, c0, c1, c2, c3, nb, P->cth + p, lam0 + p, lamM + p, sc0 + p,
                      bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5], bufs[6], bufs[7]);
          }
          double *aE = alm[0] + mstart[im], *aB = alm[1] + mstart[im];
          for (int l = m; l <= lmax; ++l) {
            if (m == 0) {
              aE[l] = (add ? aE[l] : 0.0) + c0[l]; aB[l] = (add ? aB[l] : 0.0) + c2[l];
            } else {
              aE[2*(l-m)]   = (add ? aE[2*(l-m)]   : 0.0) + nrm * c0[l];
              aE[2*(l-m)+1] = (add ? aE[2*(l-m)+1] : 0.0) + nrm * c1[l];
              aB[2*(l-m)]   = (add ? aB[2*(l-m)]   : 0.0) + nrm * c2[l];
              aB[2*(l-m)+1] = (add ? aB[2*(l-m)+1] : 0.0) + nrm * c3[l];
            }
          }
        }
      }
    }
    free(A); free(B); free(C); free(c0); free(lam0); free(lamM); free(sc0);
  }

  double t_c = wall();
  /* ---------------- synthesis: phases -> rings */
  if (synth) {
#pragma omp parallel
    {
      double *z = malloc(sizeof(double) * 2 * maxm), *work = malloc(sizeof(double) * 2 * maxm);
#pragma omp for schedule(dynamic, 4)
      for (int p = 0; p < npairs; ++p) {
        int n = P->nph[p]; fftplan *pl = get_plan(n);
        double w = (type == JOB_WY) ? P->wgt[p] : 1.0;
        for (int c = 0; c < ncomp; ++c) {
          memset(z, 0, sizeof(double) * 2 * n);
          for (int im = 0; im < nm; ++im) {
            int m = ms[im]; int k = m % n, k2 = (n - k) % n;
            double ang = fmod((double)m * P->phi0[p], 2.0 * M_PI);
            double cr = cos(ang), ci = sin(ang);
            double nr = ph[PH(c,p,0,im)], ni = ph[PH(c,p,0,im)+1];
            double sr = ph[PH(c,p,1,im)], si = ph[PH(c,p,1,im)+1];
            double pnr = nr * cr - ni * ci, pni = nr * ci + ni * cr;   /* p_m = ph_m e^{i m phi0} */
            double psr = sr * cr - si * ci, psi = sr * ci + si * cr;
            /* Z = XN + i XS ; X_{m} += p_m ; X_{-m} += conj p_m (m>0) ; m=0: X_0 += Re p_0 */
            if (m == 0) { z[0] += pnr; z[1] += psr; }
            else {
              z[2*k]    += pnr - psi; z[2*k+1]  += pni + psr;
              z[2*k2]   += pnr + psi; z[2*k2+1] += -pni + psr;
            }
          }
          cfft(pl, z, +1, work);
          double *mp = map[c];
          if (P->ofsN[p] >= 0) { double *d = mp + P->ofsN[p]; if (add) for (int j = 0; j < n; ++j) d[j] += w * z[2*j]; else for (int j = 0; j < n; ++j) d[j] = w * z[2*j]; }
          if (P->ofsS[p] >= 0) { double *d = mp + P->ofsS[p]; if (add) for (int j = 0; j < n; ++j) d[j] += w * z[2*j+1]; else for (int j = 0; j < n; ++j) d[j] = w * z[2*j+1]; }
        }
      }
      free(z); free(work);
    }
  }
  double t_d = wall();
  g_t_leg = t_c - t_b; g_t_fft = (t_b - t_a) + (t_d - t_c);
  free(ph); free(ms); free(mstart); free_pairs(P);
  return 0;
}
