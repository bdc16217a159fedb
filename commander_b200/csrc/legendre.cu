// legendre.cu -- hand-written sm_100a FP64 Legendre / Wigner-d kernels.
//
// Replaces the Legendre stage of libsharp2's sharp_execute (reached from
// commander3/src/sharp.f90:226-240) for the job types of commander3/src/sharp.f90:8-14:
//   synthesis (SHARP_Y, SHARP_WY):   ph[m][ring] = sum_l a_lm lambda_lm(theta_ring)
//   analysis  (SHARP_Yt, SHARP_YtW): a_lm       = sum_ring lambda_lm(theta_ring) ph[m][ring]
// for spin 0 (T) and spin 2 (Q,U <-> E,B; HEALPix "COSMO" convention,
// commander3/src/comm_map_mod.f90:1002).
//
// Work decomposition: grid = (ring-pair chunks, local m).  A thread owns R ring pairs
// (north ring + its southern mirror share one recurrence through the l+m parity), runs
// the recurrence in registers over l, and reads per-l data (recurrence coefficients and,
// for synthesis, the pre-scaled a_lm) as warp-uniform broadcasts from a shared-memory
// tile of TL consecutive l.  FP64-pipe cost per (l, m, ring pair):
//   spin 0: 2 (recurrence) + 2 (accumulate) ; spin 2: 4 + 8.
// Analysis reduces over rings with a register butterfly (reduce-scatter over the l of a
// group, then all-reduce) and one shared-memory hop across the warps of the CTA.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "legendre_core.cuh"

namespace cmdr {

constexpr int TL = 128;    // l per shared-memory tile, analysis kernels (one entry per thread)
constexpr int TLS = 256;   // l per tile, synthesis kernels (TLS / NT entries per thread)
constexpr int NT = 128;    // threads per CTA
constexpr unsigned FULL = 0xffffffffu;
constexpr double SCALE_DOWN = 7.458340731200207e-155;   // 2^-512

static_assert(SCALE_BITS == 512, "SCALE_DOWN must equal 2^-SCALE_BITS");

struct KParams {
  int lmax, nm, real_packed;
  int nslots, NPL, NML, ncomp_tot, comp0;
  int slot_begin;   // first slot handled by this launch (nslots = one past the last)
  const int *mval;
  const long long *mvstart;
  const double *coef;
  const long long *cofs;
  const double *Kstart;
  const double *trig;
  const int *mlim;
  double *alm0, *alm1;
  double4 *ph;
};

struct __align__(16) TileS0 { double A, ar, ai, pad; };
struct __align__(16) TileS2 { double A, C, cpr, cpi, cmr, cmi; };
struct __align__(16) TileA0 { double A, g; };
struct __align__(16) TileA2 { double A, C, g, pad; };
static_assert(TL == NT, "tile staging assumes one entry per thread");

__device__ __forceinline__ size_t ph_index(const KParams &p, int comp, int im, int slot) {
  int owner = slot / p.NPL, local = slot - owner * p.NPL;
  return ((size_t)(owner * p.ncomp_tot + p.comp0 + comp) * p.NML + im) * p.NPL + local;
}

// ------------------------------------------------------------------------------------
// spin-0 synthesis
// ------------------------------------------------------------------------------------
template <int MODE, int R>
__device__ __forceinline__ void synth0_group(const TileS0 *t, const double (&x)[R], double (&cur)[R],
                                             double (&prev)[R], double (&per)[R], double (&pei)[R],
                                             double (&por)[R], double (&poi)[R], int (&k)[R]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double A = t[j].A, ar = t[j].ar, ai = t[j].ai;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (MODE >= 1) {
        double v = (MODE == 2 || k[r] == 0) ? cur[r] : 0.0;
        if (j & 1) { por[r] = fma(v, ar, por[r]); poi[r] = fma(v, ai, poi[r]); }
        else       { per[r] = fma(v, ar, per[r]); pei[r] = fma(v, ai, pei[r]); }
      }
      double nxt = step0(A, x[r], cur[r], prev[r]);
      prev[r] = cur[r]; cur[r] = nxt;
    }
  }
  if (MODE < 2) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (k[r] < 0 && needs_rescale(cur[r])) { cur[r] *= SCALE_DOWN; prev[r] *= SCALE_DOWN; ++k[r]; }
  }
}

template <int R>
__global__ void __launch_bounds__(NT) synth0_kernel(KParams p) {
  __shared__ TileS0 tile[2][TLS];
  const int im = blockIdx.y, m = p.mval[im];
  const int tid = threadIdx.x;
  const int chunk0 = p.slot_begin + blockIdx.x * (NT * R);
  double x[R], cur[R], prev[R], per[R], pei[R], por[R], poi[R];
  int k[R], slot[R];
  bool any = false;
  const double K = p.Kstart[m];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    slot[r] = chunk0 + r * NT + tid;
    bool valid = slot[r] < p.nslots && m <= p.mlim[min(slot[r], p.nslots - 1)];
    per[r] = pei[r] = por[r] = poi[r] = 0.0;
    prev[r] = 0.0; cur[r] = 0.0; k[r] = 0; x[r] = 0.0;
    if (valid) {
      const double4 tg = reinterpret_cast<const double4 *>(p.trig)[slot[r]];
      RingTrig g{tg.x, tg.y, tg.z, tg.w};
      x[r] = g.cth;
      start_spin0(m, K, g, cur[r], k[r]);
      any = true;
    }
  }
  if (!__syncthreads_or(any)) {   // whole chunk beyond the m cut-off: phases are zero
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (slot[r] < p.nslots) p.ph[ph_index(p, 0, im, slot[r])] = make_double4(0, 0, 0, 0);
    return;
  }
  const double *coef = p.coef + p.cofs[im];
  const double *a = p.alm0;
  const long long mvs = p.mvstart[im];
  const double nrm = (p.real_packed && m > 0) ? 0.70710678118654752440 : 1.0;
  // Tiles are double buffered: the next tile's global loads are issued before the current
  // tile is consumed and land in shared memory afterwards -> one barrier per tile and the
  // global latency is hidden behind the FP64 work.
  auto load_entry = [&](int l) {
    TileS0 e{0.0, 0.0, 0.0, 0.0};
    if (l <= p.lmax) {
      double2 c = reinterpret_cast<const double2 *>(coef)[l - m];   // {A', g}
      double gs = c.y * nrm;
      e.A = c.x;
      if (p.real_packed) {
        if (m == 0) { e.ar = gs * a[mvs + l]; }
        else { e.ar = gs * a[mvs + 2 * (long long)l]; e.ai = gs * a[mvs + 2 * (long long)l + 1]; }
      } else {
        e.ar = gs * a[2 * (mvs + l)];
        e.ai = m == 0 ? 0.0 : gs * a[2 * (mvs + l) + 1];
      }
    }
    return e;
  };
#pragma unroll
  for (int q = 0; q < TLS / NT; ++q) tile[0][tid + q * NT] = load_entry(m + tid + q * NT);
  __syncthreads();
  int buf = 0;
  for (int lt = m; lt <= p.lmax; lt += TLS, buf ^= 1) {
    const bool more = lt + TLS <= p.lmax;
    TileS0 nxt[TLS / NT];
    if (more) {
#pragma unroll
      for (int q = 0; q < TLS / NT; ++q) nxt[q] = load_entry(lt + TLS + tid + q * NT);
    }
    const int ngroups = min(TLS, p.lmax - lt + 8) / 8;
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
      bool all_on = true, none_on = true;
#pragma unroll
      for (int r = 0; r < R; ++r) { all_on &= (k[r] == 0); none_on &= (k[r] < 0); }
      if (__all_sync(FULL, all_on)) synth0_group<2, R>(tile[buf] + 8 * g, x, cur, prev, per, pei, por, poi, k);
      else if (__all_sync(FULL, none_on)) synth0_group<0, R>(tile[buf] + 8 * g, x, cur, prev, per, pei, por, poi, k);
      else synth0_group<1, R>(tile[buf] + 8 * g, x, cur, prev, per, pei, por, poi, k);
    }
    if (more) {
#pragma unroll
      for (int q = 0; q < TLS / NT; ++q) tile[buf ^ 1][tid + q * NT] = nxt[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (slot[r] < p.nslots)
      p.ph[ph_index(p, 0, im, slot[r])] =
          make_double4(per[r] + por[r], pei[r] + poi[r], per[r] - por[r], pei[r] - poi[r]);
}

// ------------------------------------------------------------------------------------
// spin-2 synthesis.  cp = -(E+iB) g, cm = -(E-iB) g;  P,M = mu of (+2)lambda, (-2)lambda.
//   a1 = sum cp P, a2 = sum cm M (north);  a3 = sum sg cp M, a4 = sum sg cm P (south)
//   Q = (a1+a2)/2, U = -i (a1-a2)/2
// ------------------------------------------------------------------------------------
template <int MODE, int R>
__device__ __forceinline__ void synth2_group(const TileS2 *t, const double (&x)[R], double (&P)[R],
                                             double (&Pp)[R], double (&M)[R], double (&Mp)[R],
                                             double (&a)[R][8], int (&k)[R]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double A = t[j].A, C = t[j].C;
    const double cpr = t[j].cpr, cpi = t[j].cpi, cmr = t[j].cmr, cmi = t[j].cmi;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (MODE >= 1) {
        bool on = (MODE == 2 || k[r] == 0);
        double vp = on ? P[r] : 0.0, vm = on ? M[r] : 0.0;
        a[r][0] = fma(vp, cpr, a[r][0]); a[r][1] = fma(vp, cpi, a[r][1]);
        a[r][2] = fma(vm, cmr, a[r][2]); a[r][3] = fma(vm, cmi, a[r][3]);
        if (j & 1) {
          a[r][4] = fma(-vm, cpr, a[r][4]); a[r][5] = fma(-vm, cpi, a[r][5]);
          a[r][6] = fma(-vp, cmr, a[r][6]); a[r][7] = fma(-vp, cmi, a[r][7]);
        } else {
          a[r][4] = fma(vm, cpr, a[r][4]); a[r][5] = fma(vm, cpi, a[r][5]);
          a[r][6] = fma(vp, cmr, a[r][6]); a[r][7] = fma(vp, cmi, a[r][7]);
        }
      }
      double up = fma(A, x[r], C), um = fma(A, x[r], -C);
      double np_ = fma(up, P[r], -Pp[r]), nm_ = fma(um, M[r], -Mp[r]);
      Pp[r] = P[r]; P[r] = np_; Mp[r] = M[r]; M[r] = nm_;
    }
  }
  if (MODE < 2) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (k[r] < 0 && (needs_rescale(P[r]) || needs_rescale(M[r]))) {
        P[r] *= SCALE_DOWN; Pp[r] *= SCALE_DOWN; M[r] *= SCALE_DOWN; Mp[r] *= SCALE_DOWN; ++k[r];
      }
  }
}

template <int R>
__global__ void __launch_bounds__(NT) synth2_kernel(KParams p) {
  __shared__ TileS2 tile[2][TLS];
  const int im = blockIdx.y, m = p.mval[im];
  const int tid = threadIdx.x;
  const int chunk0 = p.slot_begin + blockIdx.x * (NT * R);
  const int l0 = max(m, 2);
  double x[R], P[R], Pp[R], M[R], Mp[R], a[R][8];
  int k[R], slot[R];
  bool any = false;
  const double K = p.Kstart[m];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    slot[r] = chunk0 + r * NT + tid;
    bool valid = slot[r] < p.nslots && m <= p.mlim[min(slot[r], p.nslots - 1)] && l0 <= p.lmax;
#pragma unroll
    for (int q = 0; q < 8; ++q) a[r][q] = 0.0;
    P[r] = M[r] = Pp[r] = Mp[r] = 0.0; k[r] = 0; x[r] = 0.0;
    if (valid) {
      const double4 tg = reinterpret_cast<const double4 *>(p.trig)[slot[r]];
      RingTrig g{tg.x, tg.y, tg.z, tg.w};
      x[r] = g.cth;
      start_spin2(m, K, g, P[r], M[r], k[r]);
      any = true;
    }
  }
  if (!__syncthreads_or(any)) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (slot[r] < p.nslots) {
        p.ph[ph_index(p, 0, im, slot[r])] = make_double4(0, 0, 0, 0);
        p.ph[ph_index(p, 1, im, slot[r])] = make_double4(0, 0, 0, 0);
      }
    return;
  }
  const double *coef = p.coef + p.cofs[im];
  const double *aE = p.alm0, *aB = p.alm1;
  const long long mvs = p.mvstart[im];
  const double nrm = (p.real_packed && m > 0) ? 0.70710678118654752440 : 1.0;
  auto load_entry = [&](int l) {
    TileS2 e{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (l <= p.lmax) {
      double4 c = reinterpret_cast<const double4 *>(coef)[l - l0];   // {A', C', g, 0}
      double gs = c.z * nrm;
      e.A = c.x; e.C = c.y;
      double er, ei = 0.0, br, bi = 0.0;
      if (p.real_packed) {
        if (m == 0) { er = aE[mvs + l]; br = aB[mvs + l]; }
        else {
          er = aE[mvs + 2 * (long long)l]; ei = aE[mvs + 2 * (long long)l + 1];
          br = aB[mvs + 2 * (long long)l]; bi = aB[mvs + 2 * (long long)l + 1];
        }
      } else {
        er = aE[2 * (mvs + l)]; br = aB[2 * (mvs + l)];
        if (m > 0) { ei = aE[2 * (mvs + l) + 1]; bi = aB[2 * (mvs + l) + 1]; }
      }
      e.cpr = -gs * (er - bi); e.cpi = -gs * (ei + br);
      e.cmr = -gs * (er + bi); e.cmi = -gs * (ei - br);
    }
    return e;
  };
#pragma unroll
  for (int q = 0; q < TLS / NT; ++q) tile[0][tid + q * NT] = load_entry(l0 + tid + q * NT);
  __syncthreads();
  int buf = 0;
  for (int lt = l0; lt <= p.lmax; lt += TLS, buf ^= 1) {
    const bool more = lt + TLS <= p.lmax;
    TileS2 nxt[TLS / NT];
    if (more) {
#pragma unroll
      for (int q = 0; q < TLS / NT; ++q) nxt[q] = load_entry(lt + TLS + tid + q * NT);
    }
    const int ngroups = min(TLS, p.lmax - lt + 8) / 8;
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
      bool all_on = true, none_on = true;
#pragma unroll
      for (int r = 0; r < R; ++r) { all_on &= (k[r] == 0); none_on &= (k[r] < 0); }
      if (__all_sync(FULL, all_on)) synth2_group<2, R>(tile[buf] + 8 * g, x, P, Pp, M, Mp, a, k);
      else if (__all_sync(FULL, none_on)) synth2_group<0, R>(tile[buf] + 8 * g, x, P, Pp, M, Mp, a, k);
      else synth2_group<1, R>(tile[buf] + 8 * g, x, P, Pp, M, Mp, a, k);
    }
    if (more) {
#pragma unroll
      for (int q = 0; q < TLS / NT; ++q) tile[buf ^ 1][tid + q * NT] = nxt[q];
    }
    __syncthreads();
  }
  // sg_l = (-1)^(l+m+2) = sg0 * (-1)^(l-l0)
  const double sg0 = ((l0 + m) & 1) ? -0.5 : 0.5;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (slot[r] < p.nslots) {
      double4 q, u;
      q.x = 0.5 * (a[r][0] + a[r][2]);  q.y = 0.5 * (a[r][1] + a[r][3]);
      u.x = 0.5 * (a[r][1] - a[r][3]);  u.y = -0.5 * (a[r][0] - a[r][2]);
      q.z = sg0 * (a[r][4] + a[r][6]);  q.w = sg0 * (a[r][5] + a[r][7]);
      u.z = sg0 * (a[r][5] - a[r][7]);  u.w = -sg0 * (a[r][4] - a[r][6]);
      p.ph[ph_index(p, 0, im, slot[r])] = q;
      p.ph[ph_index(p, 1, im, slot[r])] = u;
    }
}

// ------------------------------------------------------------------------------------
// Ring reduction helpers for analysis.  Each lane holds v[G] partial sums (one per l of
// the group); on return lane L holds the warp total of entry (L >> SH) in v[0].
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(FULL, v, mask); }

__device__ __forceinline__ void warp_reduce_scatter8(double (&v)[8], int lane) {
  {  // 8 -> 4 over lane bit 4
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
      v[i] = keep + shfl_xor_d(send, 16);
    }
  }
  {  // 4 -> 2 over lane bit 3
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
      v[i] = keep + shfl_xor_d(send, 8);
    }
  }
  {  // 2 -> 1 over lane bit 2
    const bool hi = lane & 4;
    double send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + shfl_xor_d(send, 4);
  }
  v[0] += shfl_xor_d(v[0], 2);
  v[0] += shfl_xor_d(v[0], 1);
}   // entry index held by lane: (lane >> 2) & 7

__device__ __forceinline__ void warp_reduce_scatter4(double (&v)[4], int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
      v[i] = keep + shfl_xor_d(send, 16);
    }
  }
  {
    const bool hi = lane & 8;
    double send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + shfl_xor_d(send, 8);
  }
  v[0] += shfl_xor_d(v[0], 4);
  v[0] += shfl_xor_d(v[0], 2);
  v[0] += shfl_xor_d(v[0], 1);
}   // entry index held by lane: (lane >> 3) & 3

// ------------------------------------------------------------------------------------
// spin-0 analysis:  a_l = sum_rings mu_l * (l-m even ? qN+qS : qN-qS)
// ------------------------------------------------------------------------------------
template <int MODE, int R>
__device__ __forceinline__ void anal0_group(const TileA0 *tA, const double (&x)[R], double (&cur)[R],
                                            double (&prev)[R], const double (&sr)[R], const double (&si)[R],
                                            const double (&dr)[R], const double (&di)[R], int (&k)[R],
                                            double (&accr)[8], double (&acci)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double A = tA[j].A;
    double ar = 0.0, ai = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (MODE >= 1) {
        double v = (MODE == 2 || k[r] == 0) ? cur[r] : 0.0;
        if (j & 1) { ar = fma(v, dr[r], ar); ai = fma(v, di[r], ai); }
        else       { ar = fma(v, sr[r], ar); ai = fma(v, si[r], ai); }
      }
      double nxt = step0(A, x[r], cur[r], prev[r]);
      prev[r] = cur[r]; cur[r] = nxt;
    }
    accr[j] = ar; acci[j] = ai;
  }
  if (MODE < 2) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (k[r] < 0 && needs_rescale(cur[r])) { cur[r] *= SCALE_DOWN; prev[r] *= SCALE_DOWN; ++k[r]; }
  }
}

template <int R>
__global__ void __launch_bounds__(NT) anal0_kernel(KParams p) {
  __shared__ TileA0 tile[2][TL];
  __shared__ double red[2][NT / 32][TL][2];
  const int im = blockIdx.y, m = p.mval[im];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk0 = p.slot_begin + blockIdx.x * (NT * R);
  double x[R], cur[R], prev[R], sr[R], si[R], dr[R], di[R];
  int k[R];
  bool any = false;
  const double K = p.Kstart[m];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int slot = chunk0 + r * NT + tid;
    bool valid = slot < p.nslots && m <= p.mlim[min(slot, p.nslots - 1)];
    prev[r] = cur[r] = x[r] = 0.0; k[r] = 0;
    sr[r] = si[r] = dr[r] = di[r] = 0.0;
    if (valid) {
      const double4 tg = reinterpret_cast<const double4 *>(p.trig)[slot];
      RingTrig g{tg.x, tg.y, tg.z, tg.w};
      x[r] = g.cth;
      start_spin0(m, K, g, cur[r], k[r]);
      double4 q = p.ph[ph_index(p, 0, im, slot)];
      sr[r] = q.x + q.z; si[r] = q.y + q.w; dr[r] = q.x - q.z; di[r] = q.y - q.w;
      any = true;
    }
  }
  if (!__syncthreads_or(any)) return;
  const double *coef = p.coef + p.cofs[im];
  const long long mvs = p.mvstart[im];
  // real-packed: orthonormal real basis (sqrt2 both ways).  complex a_lm: the phases carry the
  // factor 2 of the m>0 terms, so the adjoint needs 1/2 to return sum conj(Y) x as libsharp2 does
  const double nrm = m > 0 ? (p.real_packed ? 0.70710678118654752440 : 0.5) : 1.0;
  auto load_entry = [&](int l) {
    TileA0 e{0.0, 0.0};
    if (l <= p.lmax) { double2 c = reinterpret_cast<const double2 *>(coef)[l - m]; e.A = c.x; e.g = c.y; }
    return e;
  };
  tile[0][tid] = load_entry(m + tid);
  __syncthreads();
  int buf = 0;
  for (int lt = m; lt <= p.lmax; lt += TL, buf ^= 1) {
    const bool more = lt + TL <= p.lmax;
    TileA0 nxt;
    if (more) nxt = load_entry(lt + TL + tid);
    const int ngroups = min(TL, p.lmax - lt + 8) / 8;
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
      bool all_on = true, none_on = true;
#pragma unroll
      for (int r = 0; r < R; ++r) { all_on &= (k[r] == 0); none_on &= (k[r] < 0); }
      double accr[8], acci[8];
      const bool w_none = __all_sync(FULL, none_on);
      if (__all_sync(FULL, all_on)) anal0_group<2, R>(tile[buf] + 8 * g, x, cur, prev, sr, si, dr, di, k, accr, acci);
      else if (w_none) anal0_group<0, R>(tile[buf] + 8 * g, x, cur, prev, sr, si, dr, di, k, accr, acci);
      else anal0_group<1, R>(tile[buf] + 8 * g, x, cur, prev, sr, si, dr, di, k, accr, acci);
      if (!w_none) { warp_reduce_scatter8(accr, lane); warp_reduce_scatter8(acci, lane); }
      if ((lane & 3) == 0) {
        int j = (lane >> 2) & 7;
        red[buf][warp][8 * g + j][0] = w_none ? 0.0 : accr[0];
        red[buf][warp][8 * g + j][1] = w_none ? 0.0 : acci[0];
      }
    }
    if (more) tile[buf ^ 1][tid] = nxt;
    __syncthreads();
    // cross-warp sum of this tile (one l per thread); red[buf]/tile[buf] are rewritten only
    // after the next barrier
    {
      int l = lt + tid;
      if (l <= p.lmax) {
        double re = 0.0, im_ = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { re += red[buf][w][tid][0]; im_ += red[buf][w][tid][1]; }
        double gs = tile[buf][tid].g * nrm;
        double *a = p.alm0;
        if (p.real_packed) {
          if (m == 0) atomicAdd(&a[mvs + l], gs * re);
          else { atomicAdd(&a[mvs + 2 * (long long)l], gs * re); atomicAdd(&a[mvs + 2 * (long long)l + 1], gs * im_); }
        } else {
          atomicAdd(&a[2 * (mvs + l)], gs * re);
          if (m > 0) atomicAdd(&a[2 * (mvs + l) + 1], gs * im_);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// spin-2 analysis.  zp = qQ + i qU, zm = qQ - i qU per ring (north, south*sg0);
//   S1_l = sum P zpN + sg M zpS ; S2_l = sum M zmN + sg P zmS
//   E_l = -(S1+S2)/2 ; B_l = (i/2)(S1-S2)
// ------------------------------------------------------------------------------------
template <int MODE, int R>
__device__ __forceinline__ void anal2_group(const TileA2 *t, const double (&x)[R], double (&P)[R],
                                            double (&Pp)[R], double (&M)[R], double (&Mp)[R],
                                            const double (&z)[R][8], int (&k)[R], double (&acc)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double A = t[j].A, C = t[j].C;
    double s1r = 0.0, s1i = 0.0, s2r = 0.0, s2i = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (MODE >= 1) {
        bool on = (MODE == 2 || k[r] == 0);
        double vp = on ? P[r] : 0.0, vm = on ? M[r] : 0.0;
        // z: 0,1 zpN ; 2,3 zmN ; 4,5 zpS ; 6,7 zmS
        s1r = fma(vp, z[r][0], s1r); s1i = fma(vp, z[r][1], s1i);
        s2r = fma(vm, z[r][2], s2r); s2i = fma(vm, z[r][3], s2i);
        if (j & 1) {
          s1r = fma(-vm, z[r][4], s1r); s1i = fma(-vm, z[r][5], s1i);
          s2r = fma(-vp, z[r][6], s2r); s2i = fma(-vp, z[r][7], s2i);
        } else {
          s1r = fma(vm, z[r][4], s1r); s1i = fma(vm, z[r][5], s1i);
          s2r = fma(vp, z[r][6], s2r); s2i = fma(vp, z[r][7], s2i);
        }
      }
      double up = fma(A, x[r], C), um = fma(A, x[r], -C);
      double np_ = fma(up, P[r], -Pp[r]), nm_ = fma(um, M[r], -Mp[r]);
      Pp[r] = P[r]; P[r] = np_; Mp[r] = M[r]; M[r] = nm_;
    }
    acc[0][j] = s1r; acc[1][j] = s1i; acc[2][j] = s2r; acc[3][j] = s2i;
  }
  if (MODE < 2) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (k[r] < 0 && (needs_rescale(P[r]) || needs_rescale(M[r]))) {
        P[r] *= SCALE_DOWN; Pp[r] *= SCALE_DOWN; M[r] *= SCALE_DOWN; Mp[r] *= SCALE_DOWN; ++k[r];
      }
  }
}

template <int R>
__global__ void __launch_bounds__(NT) anal2_kernel(KParams p) {
  __shared__ TileA2 tile[2][TL];
  __shared__ double red[2][NT / 32][TL][4];
  const int im = blockIdx.y, m = p.mval[im];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk0 = p.slot_begin + blockIdx.x * (NT * R);
  const int l0 = max(m, 2);
  if (l0 > p.lmax) return;
  double x[R], P[R], Pp[R], M[R], Mp[R], z[R][8];
  int k[R];
  bool any = false;
  const double K = p.Kstart[m];
  const double sg0 = ((l0 + m) & 1) ? -1.0 : 1.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int slot = chunk0 + r * NT + tid;
    bool valid = slot < p.nslots && m <= p.mlim[min(slot, p.nslots - 1)];
    P[r] = M[r] = Pp[r] = Mp[r] = x[r] = 0.0; k[r] = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) z[r][q] = 0.0;
    if (valid) {
      const double4 tg = reinterpret_cast<const double4 *>(p.trig)[slot];
      RingTrig g{tg.x, tg.y, tg.z, tg.w};
      x[r] = g.cth;
      start_spin2(m, K, g, P[r], M[r], k[r]);
      double4 q = p.ph[ph_index(p, 0, im, slot)], u = p.ph[ph_index(p, 1, im, slot)];
      z[r][0] = q.x - u.y; z[r][1] = q.y + u.x;            // zpN = qQ + i qU
      z[r][2] = q.x + u.y; z[r][3] = q.y - u.x;            // zmN = qQ - i qU
      z[r][4] = sg0 * (q.z - u.w); z[r][5] = sg0 * (q.w + u.z);
      z[r][6] = sg0 * (q.z + u.w); z[r][7] = sg0 * (q.w - u.z);
      any = true;
    }
  }
  if (!__syncthreads_or(any)) return;
  const double *coef = p.coef + p.cofs[im];
  const long long mvs = p.mvstart[im];
  // real-packed: orthonormal real basis (sqrt2 both ways).  complex a_lm: the phases carry the
  // factor 2 of the m>0 terms, so the adjoint needs 1/2 to return sum conj(Y) x as libsharp2 does
  const double nrm = m > 0 ? (p.real_packed ? 0.70710678118654752440 : 0.5) : 1.0;
  auto load_entry = [&](int l) {
    TileA2 e{0.0, 0.0, 0.0, 0.0};
    if (l <= p.lmax) { double4 c = reinterpret_cast<const double4 *>(coef)[l - l0]; e.A = c.x; e.C = c.y; e.g = c.z; }
    return e;
  };
  tile[0][tid] = load_entry(l0 + tid);
  __syncthreads();
  int buf = 0;
  for (int lt = l0; lt <= p.lmax; lt += TL, buf ^= 1) {
    const bool more = lt + TL <= p.lmax;
    TileA2 nxt;
    if (more) nxt = load_entry(lt + TL + tid);
    const int ngroups = min(TL, p.lmax - lt + 4) / 4;
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
      bool all_on = true, none_on = true;
#pragma unroll
      for (int r = 0; r < R; ++r) { all_on &= (k[r] == 0); none_on &= (k[r] < 0); }
      double acc[4][4];
      const bool w_none = __all_sync(FULL, none_on);
      if (__all_sync(FULL, all_on)) anal2_group<2, R>(tile[buf] + 4 * g, x, P, Pp, M, Mp, z, k, acc);
      else if (w_none) anal2_group<0, R>(tile[buf] + 4 * g, x, P, Pp, M, Mp, z, k, acc);
      else anal2_group<1, R>(tile[buf] + 4 * g, x, P, Pp, M, Mp, z, k, acc);
      if (!w_none) {
#pragma unroll
        for (int v = 0; v < 4; ++v) warp_reduce_scatter4(acc[v], lane);
      }
      if ((lane & 7) == 0) {
        int j = (lane >> 3) & 3;
#pragma unroll
        for (int v = 0; v < 4; ++v) red[buf][warp][4 * g + j][v] = w_none ? 0.0 : acc[v][0];
      }
    }
    if (more) tile[buf ^ 1][tid] = nxt;
    __syncthreads();
    {
      int l = lt + tid;
      if (l <= p.lmax) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int w = 0; w < NT / 32; ++w)
#pragma unroll
          for (int v = 0; v < 4; ++v) s[v] += red[buf][w][tid][v];
        double gs = tile[buf][tid].g * nrm;
        double Er = -0.5 * gs * (s[0] + s[2]), Ei = -0.5 * gs * (s[1] + s[3]);
        double Br = -0.5 * gs * (s[1] - s[3]), Bi = 0.5 * gs * (s[0] - s[2]);
        double *aE = p.alm0, *aB = p.alm1;
        if (p.real_packed) {
          if (m == 0) { atomicAdd(&aE[mvs + l], Er); atomicAdd(&aB[mvs + l], Br); }
          else {
            atomicAdd(&aE[mvs + 2 * (long long)l], Er); atomicAdd(&aE[mvs + 2 * (long long)l + 1], Ei);
            atomicAdd(&aB[mvs + 2 * (long long)l], Br); atomicAdd(&aB[mvs + 2 * (long long)l + 1], Bi);
          }
        } else {
          atomicAdd(&aE[2 * (mvs + l)], Er); atomicAdd(&aB[2 * (mvs + l)], Br);
          if (m > 0) { atomicAdd(&aE[2 * (mvs + l) + 1], Ei); atomicAdd(&aB[2 * (mvs + l) + 1], Bi); }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
// Ring pairs per thread (register blocking).  Defaults were picked from the sweep recorded in
// profiles/; CMDR_SHT_R_S0 / _S2 / _A0 / _A2 override them for tuning runs.
static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

static KParams make_params(const LegGeom &g, const LegAlm &a, double *alm0, double *alm1, double4 *ph) {
  KParams p;
  p.lmax = a.lmax; p.nm = a.nm; p.real_packed = a.real_packed;
  p.slot_begin = g.slot_begin; p.nslots = g.slot_end >= 0 ? g.slot_end : g.nslots; p.NPL = g.NPL; p.NML = g.NML; p.ncomp_tot = g.ncomp_tot; p.comp0 = g.comp0;
  p.mval = a.mval; p.mvstart = a.mvstart; p.coef = a.coef; p.cofs = a.cofs; p.Kstart = a.Kstart;
  p.trig = g.trig; p.mlim = g.mlim;
  p.alm0 = alm0; p.alm1 = alm1; p.ph = ph;
  return p;
}

template <int R, typename K>
static void launch_r(K kernel, const KParams &p, int /*nslots*/, int nm, cudaStream_t st) {
  const int n = p.nslots - p.slot_begin;
  if (n <= 0) return;
  dim3 grid((n + NT * R - 1) / (NT * R), nm);
  kernel<<<grid, NT, 0, st>>>(p);
}

void launch_legendre_synth(int spin, const LegGeom &g, const LegAlm &a, const double *const *alm,
                           double4 *ph, cudaStream_t st) {
  if (a.nm == 0 || g.nslots == 0) return;
  KParams p = make_params(g, a, const_cast<double *>(alm[0]), spin ? const_cast<double *>(alm[1]) : nullptr, ph);
  static const int r0 = env_int("CMDR_SHT_R_S0", 4), r2 = env_int("CMDR_SHT_R_S2", 2);
  if (spin == 0) {
    switch (r0) {
      case 2: launch_r<2>(synth0_kernel<2>, p, g.nslots, a.nm, st); break;
      case 6: launch_r<6>(synth0_kernel<6>, p, g.nslots, a.nm, st); break;
      case 8: launch_r<8>(synth0_kernel<8>, p, g.nslots, a.nm, st); break;
      default: launch_r<4>(synth0_kernel<4>, p, g.nslots, a.nm, st); break;
    }
  } else {
    switch (r2) {
      case 1: launch_r<1>(synth2_kernel<1>, p, g.nslots, a.nm, st); break;
      case 3: launch_r<3>(synth2_kernel<3>, p, g.nslots, a.nm, st); break;
      case 4: launch_r<4>(synth2_kernel<4>, p, g.nslots, a.nm, st); break;
      default: launch_r<2>(synth2_kernel<2>, p, g.nslots, a.nm, st); break;
    }
  }
  count_launch();
  CMDR_CUDA_CHECK(cudaGetLastError());
}

void launch_legendre_anal(int spin, const LegGeom &g, const LegAlm &a, double *const *alm,
                          const double4 *ph, cudaStream_t st) {
  if (a.nm == 0 || g.nslots == 0) return;
  KParams p = make_params(g, a, alm[0], spin ? alm[1] : nullptr, const_cast<double4 *>(ph));
  static const int r0 = env_int("CMDR_SHT_R_A0", 4), r2 = env_int("CMDR_SHT_R_A2", 4);
  if (spin == 0) {
    switch (r0) {
      case 2: launch_r<2>(anal0_kernel<2>, p, g.nslots, a.nm, st); break;
      case 6: launch_r<6>(anal0_kernel<6>, p, g.nslots, a.nm, st); break;
      case 8: launch_r<8>(anal0_kernel<8>, p, g.nslots, a.nm, st); break;
      default: launch_r<4>(anal0_kernel<4>, p, g.nslots, a.nm, st); break;
    }
  } else {
    switch (r2) {
      case 3: launch_r<3>(anal2_kernel<3>, p, g.nslots, a.nm, st); break;
      case 2: launch_r<2>(anal2_kernel<2>, p, g.nslots, a.nm, st); break;
      default: launch_r<4>(anal2_kernel<4>, p, g.nslots, a.nm, st); break;
    }
  }
  count_launch();
  CMDR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace cmdr
