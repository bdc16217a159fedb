// ringfft.cu -- per-ring Fourier stage of the SHT on the rings this GPU owns.
//
// Replaces libsharp2's per-ring FFT stage (phi0 shift, aliasing fold, c2r/r2c) inside
// sharp_execute (commander3/src/sharp.f90:226-240); ring geometry as set up through
// sharp_make_subset_healpix_geom_info (commander3/src/sharp.f90:145-166).
//
// Design: the north ring and its southern mirror have the same length and phi0, so
// one complex transform of z = x_north + i x_south serves both.  Belt rings (length
// 4 nside) go through one batched cuFFT Z2Z plan.  Polar-cap rings have 4i points for
// every i < nside (thousands of distinct, mostly non-smooth lengths); they are turned
// into power-of-two circular convolutions (chirp-z / Bluestein) so that ~log2(8 nside)
// batched power-of-two cuFFT plans cover all of them with a handful of launches.
#include <cufft.h>

#include <cstdio>
#include <cstdlib>
#include <map>

#include "blue_fft.cuh"
#include "kernels.h"
#include "ring_split.cuh"

namespace cmdr {

#define CMDR_CUFFT_CHECK(x)                                                              \
  do {                                                                                   \
    cufftResult r_ = (x);                                                                \
    if (r_ != CUFFT_SUCCESS) {                                                           \
      fprintf(stderr, "cmdr_sht: cuFFT error %d at %s:%d\n", (int)r_, __FILE__, __LINE__); \
      abort();                                                                           \
    }                                                                                    \
  } while (0)

struct FParams {
  int npairs, ncomp;
  const int *nph, *shifted, *zidx, *znp, *zlen, *zblue;
  const long long *zbase, *ofsN, *ofsS;
  const double *wgt;
  double2 *buf;          // FFT work buffer
  const double2 *vtab;   // Bluestein filter spectra
  PhaseLayout L;
  double4 *ph;
  double *map0, *map1, *map2;
  const double *ps0, *ps1, *ps2;   // optional factor per local pixel on the synthesised map (nullptr: 1)
  int weighted, add;
  int ring_major;                  // phase layout inside a block (kernels.h: ph_index)
};

__device__ __forceinline__ size_t zoff(const FParams &p, int pair, int c) {
  return (size_t)p.ncomp * p.zbase[pair] + ((size_t)c * p.znp[pair] + p.zidx[pair]) * p.zlen[pair];
}
// exp(i pi r / n) with exact integer range reduction
__device__ __forceinline__ double2 expipi(long long num, int n) {
  long long r = num % (2LL * n);
  double s, c;
  sincospi((double)r / (double)n, &s, &c);
  return make_double2(c, s);
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// phase element of (m owner `src`, its local m index `im`) for component c and local ring pair `pair`
__device__ __forceinline__ double4 *ph_at(const FParams &p, const PhaseLayout &L, int src, int c, int im, int pair) {
  return p.ph + ph_index(src, L.ncomp_tot, L.comp0 + c, L.NML, L.NPL, im, L.pair0 + pair, p.ring_major);
}
__device__ __forceinline__ double *map_ptr(const FParams &p, int c) {
  return c == 0 ? p.map0 : (c == 1 ? p.map1 : p.map2);
}
__device__ __forceinline__ const double *ps_ptr(const FParams &p, int c) {
  return c == 0 ? p.ps0 : (c == 1 ? p.ps1 : p.ps2);
}

// ---- synthesis, before the FFT: fold phases into the spectrum Z = X_north + i X_south
// Generic gather form (any ring length, aliasing allowed): bin k sums the m == +-k (mod n).
// Short rings have few bins and long m-chains, so the chain of each bin is split over
// `parts` threads and combined through shared memory.
__device__ __forceinline__ double2 fold_one(const FParams &p, const PhaseLayout &L, int c, int pair, int n,
                                            bool shifted, int m, bool conj_term) {
  int src = L.m2src[m];
  if (src < 0) return make_double2(0.0, 0.0);
  double4 q = *ph_at(p, L, src, c, L.m2im[m], pair);
  double2 e = shifted ? expipi(m, n) : make_double2(1.0, 0.0);
  double2 pn = cmul(make_double2(q.x, q.y), e), ps = cmul(make_double2(q.z, q.w), e);
  if (conj_term) return make_double2(pn.x + ps.y, -pn.y + ps.x);       // conj p_m
  if (m == 0) return make_double2(pn.x, ps.x);
  return make_double2(pn.x - ps.y, pn.y + ps.x);
}

template <int NT, bool PAD = false>
__device__ __forceinline__ void fold_into(const FParams &p, int pair, int c, double2 *out, int len, double2 *part_sum) {
  const int n = p.nph[pair];
  const bool blue = p.zblue[pair], shifted = p.shifted[pair];
  const PhaseLayout &L = p.L;
  int parts = 1;
  while (parts * 2 * n <= NT) parts *= 2;           // power of two, parts * n <= NT
  if (parts > 1) {
    const int k = threadIdx.x % n, part = threadIdx.x / n;   // threads >= parts*n idle
    double2 acc = make_double2(0.0, 0.0);
    if (part < parts) {
      for (int m = k + part * n; m <= L.mmax; m += parts * n) {
        double2 t = fold_one(p, L, c, pair, n, shifted, m, false);
        acc.x += t.x; acc.y += t.y;
      }
      int start = (n - k) % n;
      if (start == 0) start = n;
      for (int m = start + part * n; m <= L.mmax; m += parts * n) {
        double2 t = fold_one(p, L, c, pair, n, shifted, m, true);
        acc.x += t.x; acc.y += t.y;
      }
    }
    part_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int kk = threadIdx.x; kk < len; kk += NT) {
      double2 tot = make_double2(0.0, 0.0);
      if (kk < n) {
        for (int q = 0; q < parts; ++q) { tot.x += part_sum[kk + q * n].x; tot.y += part_sum[kk + q * n].y; }
        if (blue) tot = cmul(tot, expipi((long long)kk * kk, n));
      }
      out[bf_pidx<PAD>(kk)] = tot;
    }
    return;
  }
  for (int k = threadIdx.x; k < len; k += NT) {
    double2 acc = make_double2(0.0, 0.0);
    if (k < n) {
      for (int m = k; m <= L.mmax; m += n) {          // m == k (mod n): X_k += p_m
        double2 t = fold_one(p, L, c, pair, n, shifted, m, false);
        acc.x += t.x; acc.y += t.y;
      }
      int start = (n - k) % n;
      if (start == 0) start = n;
      for (int m = start; m <= L.mmax; m += n) {      // m == -k (mod n): X_k += conj p_m
        double2 t = fold_one(p, L, c, pair, n, shifted, m, true);
        acc.x += t.x; acc.y += t.y;
      }
      if (blue) acc = cmul(acc, expipi((long long)k * k, n));
    }
    out[bf_pidx<PAD>(k)] = acc;
  }
}

__global__ void __launch_bounds__(256) fold_kernel(FParams p, int first_pair) {
  __shared__ double2 part_sum[256];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  fold_into<256>(p, pair, c, p.buf + zoff(p, pair, c), p.zlen[pair], part_sum);
}

// Rings without aliasing (n >= 2 mmax + 1, the belt at lmax <= 2 nside): every m owns the two
// bins k = m and k = n - m, so the fold is a transpose.  A 32 (m) x 32 (pair) tile goes through
// shared memory so that both the phase reads (pair-contiguous) and the spectrum writes
// (m-contiguous) are coalesced.  The bins in between are cleared by a memset beforehand.
__global__ void __launch_bounds__(256) fold_transpose_kernel(FParams p, int first_pair, int npairs_r, int n) {
  __shared__ double4 tile[32][33];
  const PhaseLayout &L = p.L;
  const int c = blockIdx.z;
  const int e0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int e = e0 + r, pr = p0 + tx;
    double4 q = make_double4(0, 0, 0, 0);
    if (e < L.nm_total && pr < npairs_r)
      q = *ph_at(p, L, L.mlist_src[e], c, L.mlist_im[e], first_pair + pr);
    tile[r][tx] = q;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int pr = p0 + r, e = e0 + tx;
    if (pr >= npairs_r || e >= L.nm_total) continue;
    const int pair = first_pair + pr;
    const int m = L.mlist[e];
    double4 q = tile[tx][r];
    double2 ph = p.shifted[pair] ? expipi(m, n) : make_double2(1.0, 0.0);
    double2 pn = cmul(make_double2(q.x, q.y), ph), ps = cmul(make_double2(q.z, q.w), ph);
    double2 *out = p.buf + zoff(p, pair, c);
    if (m == 0) out[0] = make_double2(pn.x, ps.x);
    else {
      out[m] = make_double2(pn.x - ps.y, pn.y + ps.x);
      out[n - m] = make_double2(pn.x + ps.y, -pn.y + ps.x);
    }
  }
}

// ---- Bluestein: multiply the spectrum of the chirped input by the filter spectrum
__global__ void __launch_bounds__(256) blue_mul_kernel(FParams p, int first_pair) {
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  const int len = p.zlen[pair];
  double2 *u = p.buf + zoff(p, pair, c);
  const double2 *v = p.vtab + p.zbase[pair] + (size_t)p.zidx[pair] * len;
  for (int k = threadIdx.x; k < len; k += blockDim.x) u[k] = cmul(u[k], v[k]);
}

__global__ void __launch_bounds__(256) blue_filter_kernel(FParams p, int first_pair, double2 *vtab) {
  const int pair = first_pair + blockIdx.x;
  const int n = p.nph[pair], len = p.zlen[pair];
  double2 *v = vtab + p.zbase[pair] + (size_t)p.zidx[pair] * len;
  for (int k = threadIdx.x; k < len; k += blockDim.x) {
    int j = k < n ? k : (k > len - n ? len - k : -1);
    double2 val = make_double2(0.0, 0.0);
    if (j >= 0) { double2 e = expipi((long long)j * j, n); val = make_double2(e.x, -e.y); }
    v[k] = val;
  }
}

// ---- synthesis, after the FFT: write north = Re z, south = Im z into the map
template <bool PAD = false>
__device__ __forceinline__ void scatter_from(const FParams &p, int pair, int c, const double2 *in) {
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair];
  double *mp = map_ptr(p, c);
  const double *ps = ps_ptr(p, c);
  const long long oN = p.ofsN[pair], oS = p.ofsS[pair];
  double w = p.weighted ? p.wgt[pair] : 1.0;
  if (blue) w /= (double)len;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double2 z = in[bf_pidx<PAD>(j)];
    if (blue) z = cmul(z, expipi((long long)j * j, n));
    double vn = w * z.x, vs = w * z.y;
    if (ps) { if (oN >= 0) vn *= ps[oN + j]; if (oS >= 0) vs *= ps[oS + j]; }
    if (oN >= 0) { if (p.add) mp[oN + j] += vn; else mp[oN + j] = vn; }
    if (oS >= 0) { if (p.add) mp[oS + j] += vs; else mp[oS + j] = vs; }
  }
}
__global__ void __launch_bounds__(256) scatter_kernel(FParams p, int first_pair) {
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  scatter_from(p, pair, c, p.buf + zoff(p, pair, c));
}

// ---- analysis, before the FFT: z = w (x_north + i x_south)
template <bool PAD = false>
__device__ __forceinline__ void gather_into(const FParams &p, int pair, int c, double2 *out) {
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair];
  const double *mp = map_ptr(p, c);
  const long long oN = p.ofsN[pair], oS = p.ofsS[pair];
  const double w = p.weighted ? p.wgt[pair] : 1.0;
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    double2 z = make_double2(0.0, 0.0);
    if (j < n) {
      if (oN >= 0) z.x = w * mp[oN + j];
      if (oS >= 0) z.y = w * mp[oS + j];
      // Bluestein computes DFT+ ; DFT-(z) = conj(DFT+(conj z))
      if (blue) z = cmul(make_double2(z.x, -z.y), expipi((long long)j * j, n));
    }
    out[bf_pidx<PAD>(j)] = z;
  }
}
__global__ void __launch_bounds__(256) gather_kernel(FParams p, int first_pair) {
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  gather_into(p, pair, c, p.buf + zoff(p, pair, c));
}

// ---- analysis, after the FFT: phases ph_m = c_m e^{-i m phi0} X_{m mod n}
template <bool PAD = false>
__device__ __forceinline__ void unfold_from(const FParams &p, int pair, int c, const double2 *in) {
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair], shifted = p.shifted[pair];
  const PhaseLayout &L = p.L;
  const double inv = blue ? 1.0 / (double)len : 1.0;
  for (int e = threadIdx.x; e < L.nm_total; e += blockDim.x) {
    const int m = L.mlist[e];
    const int k = m % n, k2 = (n - k) % n;
    double2 a = in[bf_pidx<PAD>(k)], b = in[bf_pidx<PAD>(k2)];
    if (blue) {
      a = cmul(a, expipi((long long)k * k, n));   a = make_double2(a.x * inv, -a.y * inv);
      b = cmul(b, expipi((long long)k2 * k2, n)); b = make_double2(b.x * inv, -b.y * inv);
    }
    // XN = (Z_k + conj Z_k2)/2 ; XS = (Z_k - conj Z_k2)/(2i)
    double2 xn = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
    double2 xs = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));
    double cm = m == 0 ? 1.0 : 2.0;
    double2 f = make_double2(cm, 0.0);
    if (shifted) { double2 t = expipi(m, n); f = make_double2(cm * t.x, -cm * t.y); }
    xn = cmul(xn, f); xs = cmul(xs, f);
    *ph_at(p, L, L.mlist_src[e], c, L.mlist_im[e], pair) = make_double4(xn.x, xn.y, xs.x, xs.y);
  }
}
__global__ void __launch_bounds__(256) unfold_kernel(FParams p, int first_pair) {
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  unfold_from(p, pair, c, p.buf + zoff(p, pair, c));
}

// ---- fused chirp-z ring transform for the Bluestein classes whose work length fits in shared memory
// (M <= 8192: 128 KB): fold | gather -> FFT_M -> .* V -> IFFT_M -> scatter | unfold in ONE kernel, one CTA per
// (ring pair, component).  The separate path moves the padded length-M array through HBM about 13 times (fold
// write, two cuFFT passes, the multiply, two more cuFFT passes, scatter read); here HBM sees the phases, the
// filter spectrum and the map once.  FFT pair: blue_fft.cuh (DIF forward, bit-reversed V, DIT inverse).
template <int DIR, int NT>
__global__ void __launch_bounds__(NT) blue_fused_kernel(FParams p, int first_pair, int M, const double2 *__restrict__ tw,
                                                        const double2 *__restrict__ vbr) {
  extern __shared__ double2 u_sm[];                 // M work elements, then the pass-major twiddle table
  __shared__ double2 part_sum[DIR == 0 ? NT : 1];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  double2 *T = u_sm + M;
  const int ntw = bf_tw_total(M);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  if (DIR == 0) fold_into<NT>(p, pair, c, u_sm, M, part_sum); else gather_into(p, pair, c, u_sm);
  __syncthreads();
  const int np = bf_num_passes(M);
  for (int k = 0; k < np; ++k) {
    int h, fused;
    bf_pass(M, k, &h, &fused);
    const double2 *Tk = T + bf_tw_offset(M, k);
    if (fused) { for (int q = threadIdx.x; q < M / 4; q += NT) dif_item4(u_sm, h, Tk, q); }
    else { for (int q = threadIdx.x; q < M / 2; q += NT) dif_item2(u_sm, q); }
    __syncthreads();
  }
  const double2 *v = vbr + p.zbase[pair] + (size_t)p.zidx[pair] * M;
  for (int k = threadIdx.x; k < M; k += NT) u_sm[k] = cmul(u_sm[k], v[k]);
  __syncthreads();
  for (int k = np - 1; k >= 0; --k) {
    int h, fused;
    bf_pass(M, k, &h, &fused);
    const double2 *Tk = T + bf_tw_offset(M, k);
    if (fused) { for (int q = threadIdx.x; q < M / 4; q += NT) dit_item4(u_sm, h, Tk, q); }
    else { for (int q = threadIdx.x; q < M / 2; q += NT) dit_item2(u_sm, q); }
    __syncthreads();
  }
  if (DIR == 0) scatter_from(p, pair, c, u_sm); else unfold_from(p, pair, c, u_sm);
}

// Register-blocked variant of the fused kernel (blue_fft.cuh, bf2_* schedule): 3-4 stages per pass on 8 or 16 elements
// held in registers, the array padded by one slot per 16 elements so that no pass has shared-memory bank conflicts, and
// the last DIF pass, the filter multiplication and the first DIT pass done in registers in one go.  An 8192-point
// convolution makes 6 shared-memory round trips and 7 barriers instead of 15 and 16.  EXPERIMENTAL: selected with
// CMDR_SHT_FFT_BLOCKED=1; verified on the host (tests/host_emul) and against the cuFFT path, not yet the default.
template <int DIR, int NT>
__global__ void __launch_bounds__(NT) blue_fused2_kernel(FParams p, int first_pair, int M, const double2 *__restrict__ tw,
                                                         const double2 *__restrict__ vbr) {
  extern __shared__ double2 u_sm[];                 // bf_padded(M) work slots, then the twiddle table
  __shared__ double2 part_sum[DIR == 0 ? NT : 1];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  double2 *T = u_sm + bf_padded(M);
  const int ntw = bf2_tw_total(M);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  if (DIR == 0) fold_into<NT, true>(p, pair, c, u_sm, M, part_sum); else gather_into<true>(p, pair, c, u_sm);
  __syncthreads();
  const int ns = bf2_num_strided(M);
  for (int k = 0; k < ns; ++k) {
    const int S = bf2_stages(M, k), h = bf2_half(M, k);
    const double2 *Tk = T + bf2_tw_offset(M, k);
    if (S == 3) { for (int q = threadIdx.x; q < (M >> 3); q += NT) dif_itemS<3, true>(u_sm, h, Tk, q); }
    else { for (int q = threadIdx.x; q < (M >> 4); q += NT) dif_itemS<4, true>(u_sm, h, Tk, q); }
    __syncthreads();
  }
  const double2 *v = vbr + p.zbase[pair] + (size_t)p.zidx[pair] * M;
  for (int q = threadIdx.x; q < (M >> 4); q += NT) conv_mid16<true>(u_sm, v, q);
  __syncthreads();
  for (int k = ns - 1; k >= 0; --k) {
    const int S = bf2_stages(M, k), h = bf2_half(M, k);
    const double2 *Tk = T + bf2_tw_offset(M, k);
    if (S == 3) { for (int q = threadIdx.x; q < (M >> 3); q += NT) dit_itemS<3, true>(u_sm, h, Tk, q); }
    else { for (int q = threadIdx.x; q < (M >> 4); q += NT) dit_itemS<4, true>(u_sm, h, Tk, q); }
    __syncthreads();
  }
  if (DIR == 0) scatter_from<true>(p, pair, c, u_sm); else unfold_from<true>(p, pair, c, u_sm);
}

// ---- in-place transforms of the register-blocked schedule (blue_fft.cuh, bf2_*) on a padded shared-memory array,
// run by all NT threads of the CTA; every pass ends with a barrier
template <int NT>
__device__ __forceinline__ void sm_dif_strided(double2 *x, int M, const double2 *T) {
  const int ns = bf2_num_strided(M);
  for (int k = 0; k < ns; ++k) {
    const int S = bf2_stages(M, k), h = bf2_half(M, k);
    const double2 *Tk = T + bf2_tw_offset(M, k);
    if (S == 3) { for (int q = threadIdx.x; q < (M >> 3); q += NT) dif_itemS<3, true>(x, h, Tk, q); }
    else { for (int q = threadIdx.x; q < (M >> 4); q += NT) dif_itemS<4, true>(x, h, Tk, q); }
    __syncthreads();
  }
}
template <int NT>
__device__ __forceinline__ void sm_dit_strided(double2 *x, int M, const double2 *T) {
  for (int k = bf2_num_strided(M) - 1; k >= 0; --k) {
    const int S = bf2_stages(M, k), h = bf2_half(M, k);
    const double2 *Tk = T + bf2_tw_offset(M, k);
    if (S == 3) { for (int q = threadIdx.x; q < (M >> 3); q += NT) dit_itemS<3, true>(x, h, Tk, q); }
    else { for (int q = threadIdx.x; q < (M >> 4); q += NT) dit_itemS<4, true>(x, h, Tk, q); }
    __syncthreads();
  }
}
// x <- M * IDFT( DFT(x) .* V ), V in bit-reversed order (global memory)
template <int NT>
__device__ __forceinline__ void sm_convolve(double2 *x, int M, const double2 *T, const double2 *__restrict__ v) {
  sm_dif_strided<NT>(x, M, T);
  for (int q = threadIdx.x; q < (M >> 4); q += NT) conv_mid16<true>(x, v, q);
  __syncthreads();
  sm_dit_strided<NT>(x, M, T);
}
// plain transforms: DIF forward (e^-, natural in, bit-reversed out) and DIT inverse (e^+, bit-reversed in, natural out)
template <int NT>
__device__ __forceinline__ void sm_fft_dif(double2 *x, int M, const double2 *T) {
  sm_dif_strided<NT>(x, M, T);
  const double2 *Tone = T + bf2_tw_total(M) - 1;
  for (int q = threadIdx.x; q < (M >> 4); q += NT) dif_itemS<4, true>(x, 8, Tone, q);
  __syncthreads();
}
template <int NT>
__device__ __forceinline__ void sm_fft_dit(double2 *x, int M, const double2 *T) {
  const double2 *Tone = T + bf2_tw_total(M) - 1;
  for (int q = threadIdx.x; q < (M >> 4); q += NT) dit_itemS<4, true>(x, 8, Tone, q);
  __syncthreads();
  sm_dit_strided<NT>(x, M, T);
}

// Folded spectrum bin k of a ring WITHOUT the phi0 shift factor: sum over m == +-k (mod n) of the phases with the
// alternating sign the shift e^{i pi m / n} leaves once e^{i pi k / n} is pulled out (ring_split.cuh), so that one
// sincospi per bin replaces one per term.
__device__ __forceinline__ double2 fold_bin(const FParams &p, const PhaseLayout &L, int c, int pair, int n, bool shifted, int k) {
  double2 acc = make_double2(0.0, 0.0);
  double sg = 1.0;
  for (int m = k; m <= L.mmax; m += n, sg = shifted ? -sg : sg) {            // m == k: p_m
    const int src = L.m2src[m];
    if (src < 0) continue;
    const double4 q = *ph_at(p, L, src, c, L.m2im[m], pair);
    if (m == 0) { acc.x += q.x; acc.y += q.z; }
    else { acc.x += sg * (q.x - q.w); acc.y += sg * (q.y + q.z); }
  }
  sg = shifted ? -1.0 : 1.0;
  for (int m = n - k; m <= L.mmax; m += n, sg = shifted ? -sg : sg) {        // m == -k, m >= 1: conj p_m
    const int src = L.m2src[m];
    if (src < 0) continue;
    const double4 q = *ph_at(p, L, src, c, L.m2im[m], pair);
    acc.x += sg * (q.x + q.w); acc.y += sg * (q.z - q.y);
  }
  return acc;
}

// The folded spectrum of a whole ring into shared memory, position-major: thread t fills position t of z, which holds
// bin k = binof(t) (the identity, or the bit reversal the DIT transform wants -- a phase element is exactly one 32-byte
// sector, so reading the phases in permuted order costs no extra DRAM traffic, while writing shared memory in
// permuted order would serialise on bank conflicts).  Rings with n > mmax (no aliasing: bin k receives at most the
// direct term of m = k and the conjugate term of m = n - k) take NB positions per thread and round with all 2 NB
// loads in flight before the first is used; the chained table look-up -> phase load -> next bin of fold_bin left the
// 16-warp CTAs of the whole-ring kernels waiting on DRAM latency.  fac(k): factor of bin k (phi0 shift, chirp).
template <int NT, int NB, class BinOf, class Slot, class Fac>
__device__ __forceinline__ void fold_positions(const FParams &p, const PhaseLayout &L, int c, int pair, int n, bool shifted, double2 *z,
                                               BinOf binof, Slot slot, Fac fac) {
  if (n <= L.mmax) {                                // aliasing: several m per bin
    for (int t = threadIdx.x; t < n; t += NT) {
      const int k = binof(t);
      z[slot(t)] = cmul(fold_bin(p, L, c, pair, n, shifted, k), fac(k));
    }
    return;
  }
  const double sgc = shifted ? -1.0 : 1.0;          // e^{i pi m / n} of m = n - k leaves (-1) next to e^{i pi k / n}
  for (int t0 = threadIdx.x; t0 < n; t0 += NT * NB) {
    double4 qd[NB], qc[NB];
    bool hd[NB], hc[NB];
    int kk[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int t = t0 + j * NT;
      hd[j] = hc[j] = false;
      kk[j] = 0;
      if (t >= n) continue;
      const int k = binof(t), m2 = n - k;
      kk[j] = k;
      if (k <= L.mmax) {
        const int src = L.m2src[k];
        if (src >= 0) { hd[j] = true; qd[j] = *ph_at(p, L, src, c, L.m2im[k], pair); }
      }
      if (m2 <= L.mmax) {                           // m2 >= 1 always (k < n)
        const int src = L.m2src[m2];
        if (src >= 0) { hc[j] = true; qc[j] = *ph_at(p, L, src, c, L.m2im[m2], pair); }
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int t = t0 + j * NT;
      if (t >= n) continue;
      double2 acc = make_double2(0.0, 0.0);
      if (hd[j]) {
        if (kk[j] == 0) acc = make_double2(qd[j].x, qd[j].z);
        else acc = make_double2(qd[j].x - qd[j].w, qd[j].y + qd[j].z);
      }
      if (hc[j]) { acc.x += sgc * (qc[j].x + qc[j].w); acc.y += sgc * (qc[j].z - qc[j].y); }
      z[slot(t)] = (hd[j] || hc[j]) ? cmul(acc, fac(kk[j])) : acc;
    }
  }
}

// phases of all m from the two DFT- bins Z_k, Z_{n-k} of z = w (x_north + i x_south)
__device__ __forceinline__ void unfold_store(const FParams &p, const PhaseLayout &L, int c, int pair, int n, bool shifted, int e,
                                             double2 a, double2 b) {
  const int m = L.mlist[e];
  // XN = (Z_k + conj Z_k2)/2 ; XS = (Z_k - conj Z_k2)/(2i)
  double2 xn = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
  double2 xs = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));
  const double cm = m == 0 ? 1.0 : 2.0;
  double2 f = make_double2(cm, 0.0);
  if (shifted) { double2 t = expipi(m, n); f = make_double2(cm * t.x, -cm * t.y); }
  xn = cmul(xn, f); xs = cmul(xs, f);
  *ph_at(p, L, L.mlist_src[e], c, L.mlist_im[e], pair) = make_double4(xn.x, xn.y, xs.x, xs.y);
}

// ---- long polar-cap rings (n = 4 i, chirp-z work length of the whole ring too large for shared memory): radix-4
// split into four length-i chirp-z transforms of work length M (ring_split.cuh).  One CTA per (ring pair, component).
// Shared memory: zbuf (n <= nmax elements, four sub-spectra r-major) | work (bf_padded(M)) | twiddles.
template <int DIR, int NT>
__global__ void __launch_bounds__(NT) ring_split_kernel(FParams p, int first_pair, int M, int nmax, const double2 *__restrict__ tw,
                                                        const double2 *__restrict__ vsub) {
  extern __shared__ double2 u_sm[];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  const int n = p.nph[pair], i = n >> 2;
  const bool shifted = p.shifted[pair];
  const PhaseLayout &L = p.L;
  double2 *zbuf = u_sm, *work = u_sm + nmax, *T = work + bf_padded(M);
  const int ntw = bf2_tw_total(M);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  const long long oN = p.ofsN[pair], oS = p.ofsS[pair];
  if (DIR == 0) {
    // folded spectrum, phi0 shift and input chirp in one factor per bin
    fold_positions<NT, 4>(p, L, c, pair, n, shifted, zbuf, [](int t) { return t; }, [i](int t) { return rs_slot(t, i); },
                          [n, shifted](int k) { return rs_expipi(rs_fold_angle(k, shifted), n); });
  } else {
    // radix-4 pass over conj(z), twiddle and input chirp
    const double *mp = map_ptr(p, c);
    const double w = p.weighted ? p.wgt[pair] : 1.0;
    for (int t = threadIdx.x; t < i; t += NT) {
      double2 y[4], o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        y[q].x = oN >= 0 ? w * mp[oN + t + q * i] : 0.0;
        y[q].y = oS >= 0 ? -w * mp[oS + t + q * i] : 0.0;
      }
      rs_butterfly4(y, o);
#pragma unroll
      for (int r = 0; r < 4; ++r) zbuf[r * i + t] = cmul(o[r], rs_expipi(rs_twiddle_angle(t, r), n));
    }
  }
  __syncthreads();
  const double2 *v = vsub + (size_t)(pair - first_pair) * M;
  for (int r = 0; r < 4; ++r) {
    for (int t = threadIdx.x; t < M; t += NT) work[bf_pidx<true>(t)] = t < i ? zbuf[r * i + t] : make_double2(0.0, 0.0);
    __syncthreads();
    sm_convolve<NT>(work, M, T, v);
    for (int t = threadIdx.x; t < i; t += NT) zbuf[r * i + t] = work[bf_pidx<true>(t)];
    __syncthreads();
  }
  const double invM = 1.0 / (double)M;
  if (DIR == 0) {
    double *mp = map_ptr(p, c);
    const double *ps = ps_ptr(p, c);
    const double w = (p.weighted ? p.wgt[pair] : 1.0) * invM;
    for (int t = threadIdx.x; t < i; t += NT) {
      double2 cr[4], o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) cr[r] = cmul(zbuf[r * i + t], rs_expipi(rs_twiddle_angle(t, r), n));
      rs_butterfly4(cr, o);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = t + q * i;
        double vn = w * o[q].x, vs = w * o[q].y;
        if (ps) { if (oN >= 0) vn *= ps[oN + j]; if (oS >= 0) vs *= ps[oS + j]; }
        if (oN >= 0) { if (p.add) mp[oN + j] += vn; else mp[oN + j] = vn; }
        if (oS >= 0) { if (p.add) mp[oS + j] += vs; else mp[oS + j] = vs; }
      }
    }
  } else {
    // Z_k = conj( chirp_a / M * conv_r[a] ),  k = 4 a + r
    auto bin = [&](int k) {
      const long long a = k >> 2;
      const double2 g = cmul(zbuf[rs_slot(k, i)], rs_expipi(4 * a * a, n));
      return make_double2(g.x * invM, -g.y * invM);
    };
    for (int e = threadIdx.x; e < L.nm_total; e += NT) {
      const int k = L.mlist[e] % n, k2 = (n - k) % n;
      unfold_store(p, L, c, pair, n, shifted, e, bin(k), bin(k2));
    }
  }
}

// filter spectra of the sub-transforms: V_i = DFT_M of v_j = e^{-i pi j^2 / i} (|j| < i, wrapped), left in the bit-reversed
// order the DIF transform produces, which is the order conv_mid16 consumes
template <int NT>
__global__ void __launch_bounds__(NT) ring_split_filter_kernel(FParams p, int first_pair, int M, const double2 *__restrict__ tw,
                                                               double2 *__restrict__ vsub) {
  extern __shared__ double2 u_sm[];
  const int pair = first_pair + blockIdx.x;
  const int i = p.nph[pair] >> 2;
  double2 *work = u_sm, *T = work + bf_padded(M);
  const int ntw = bf2_tw_total(M);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  for (int k = threadIdx.x; k < M; k += NT) {
    const long long j = k < i ? k : (k > M - i ? M - k : -1);
    double2 val = make_double2(0.0, 0.0);
    if (j >= 0) { const double2 e = rs_expipi(j * j, i); val = make_double2(e.x, -e.y); }
    work[bf_pidx<true>(k)] = val;
  }
  __syncthreads();
  sm_fft_dif<NT>(work, M, T);
  double2 *out = vsub + (size_t)(pair - first_pair) * M;
  for (int k = threadIdx.x; k < M; k += NT) out[k] = work[bf_pidx<true>(k)];
}

// ---- rings of power-of-two length n (the belt: 4 nside): one in-place FFT of the whole ring pair
template <int DIR, int NT>
__global__ void __launch_bounds__(NT) ring_pow2_kernel(FParams p, int first_pair, int n, int bits, const double2 *__restrict__ tw) {
  extern __shared__ double2 u_sm[];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  const bool shifted = p.shifted[pair];
  const PhaseLayout &L = p.L;
  double2 *work = u_sm, *T = work + bf_padded(n);
  const int ntw = bf2_tw_total(n);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  const long long oN = p.ofsN[pair], oS = p.ofsS[pair];
  if (DIR == 0) {
    // bins in natural order (coalesced phase reads), stored at their bit-reversed positions: the two-level padding of
    // bf_pidx keeps those stores free of bank conflicts
    fold_positions<NT, 4>(p, L, c, pair, n, shifted, work, [](int t) { return t; },
                          [bits](int t) { return bf_pidx<true>((int)(__brev((unsigned)t) >> (32 - bits))); },
                          [n, shifted](int k) { return shifted ? rs_expipi(k, n) : make_double2(1.0, 0.0); });
    __syncthreads();
    sm_fft_dit<NT>(work, n, T);
    double *mp = map_ptr(p, c);
    const double *ps = ps_ptr(p, c);
    const double w = p.weighted ? p.wgt[pair] : 1.0;
    for (int j = threadIdx.x; j < n; j += NT) {
      const double2 z = work[bf_pidx<true>(j)];
      double vn = w * z.x, vs = w * z.y;
      if (ps) { if (oN >= 0) vn *= ps[oN + j]; if (oS >= 0) vs *= ps[oS + j]; }
      if (oN >= 0) { if (p.add) mp[oN + j] += vn; else mp[oN + j] = vn; }
      if (oS >= 0) { if (p.add) mp[oS + j] += vs; else mp[oS + j] = vs; }
    }
  } else {
    const double *mp = map_ptr(p, c);
    const double w = p.weighted ? p.wgt[pair] : 1.0;
    for (int j = threadIdx.x; j < n; j += NT)
      work[bf_pidx<true>(j)] = make_double2(oN >= 0 ? w * mp[oN + j] : 0.0, oS >= 0 ? w * mp[oS + j] : 0.0);
    __syncthreads();
    sm_fft_dif<NT>(work, n, T);
    for (int e = threadIdx.x; e < L.nm_total; e += NT) {
      const int k = L.mlist[e] % n, k2 = (n - k) % n;
      const double2 a = work[bf_pidx<true>((int)(__brev((unsigned)k) >> (32 - bits)))];
      const double2 b = work[bf_pidx<true>((int)(__brev((unsigned)k2) >> (32 - bits)))];
      unfold_store(p, L, c, pair, n, shifted, e, a, b);
    }
  }
}

// ---- rings of power-of-two length, ONE RING per CTA through a complex transform of half the ring length
// (ring_split.cuh (3)): 64 KB + padding for the belt of nside 2048 instead of 128 KB per ring pair, so that three CTAs
// share an SM and one CTA's DRAM phases (fold, scatter) run beside the FFT passes of the others.  grid.x = 2 pairs
// (hemisphere = blockIdx.x & 1: the north and the south CTA of a pair read the two halves of the same 32-byte phase
// elements back to back, the second one from L2).
__device__ __forceinline__ double2 ph_half_load(const FParams &p, const PhaseLayout &L, int src, int c, int im, int pair, int hemi) {
  return reinterpret_cast<const double2 *>(ph_at(p, L, src, c, im, pair))[hemi];
}
// folded spectrum bin b (0 <= b < n) of one ring without the phi0 factor e^{i pi b / n} (see fold_bin)
__device__ __forceinline__ double2 fold_bin1(const FParams &p, const PhaseLayout &L, int c, int pair, int n, bool shifted, int b, int hemi) {
  double2 acc = make_double2(0.0, 0.0);
  double sg = 1.0;
  for (int m = b; m <= L.mmax; m += n, sg = shifted ? -sg : sg) {            // m == b: p_m
    const int src = L.m2src[m];
    if (src < 0) continue;
    const double2 q = ph_half_load(p, L, src, c, L.m2im[m], pair, hemi);
    if (m == 0) acc.x += q.x; else { acc.x += sg * q.x; acc.y += sg * q.y; }
  }
  sg = shifted ? -1.0 : 1.0;
  for (int m = n - b; m <= L.mmax; m += n, sg = shifted ? -sg : sg) {        // m == -b, m >= 1: conj p_m
    const int src = L.m2src[m];
    if (src < 0) continue;
    const double2 q = ph_half_load(p, L, src, c, L.m2im[m], pair, hemi);
    acc.x += sg * q.x; acc.y -= sg * q.y;
  }
  return acc;
}

template <int DIR, int NT>
__global__ void __launch_bounds__(NT) ring_half_kernel(FParams p, int first_pair, int n, int bitsh, const double2 *__restrict__ tw) {
  extern __shared__ double2 u_sm[];
  const int pair = first_pair + (blockIdx.x >> 1), hemi = blockIdx.x & 1, c = blockIdx.y;
  const int h = n >> 1;
  const bool shifted = p.shifted[pair];
  const PhaseLayout &L = p.L;
  const long long o = hemi ? p.ofsS[pair] : p.ofsN[pair];
  if (DIR == 0 && o < 0) return;                       // ring absent (equator mirror, ring subsets): nothing to write
  double2 *work = u_sm, *T = work + bf_padded(h);
  const int ntw = bf2_tw_total(h);
  for (int k = threadIdx.x; k < ntw; k += NT) T[k] = tw[k];
  const double w = p.weighted ? p.wgt[pair] : 1.0;
  auto slot = [bitsh](int k) { return bf_pidx<true>((int)(__brev((unsigned)k) >> (32 - bitsh))); };
  if (DIR == 0) {
    // position k holds Y_k = e^{i pi k / n} [ (X_k + f X_{k+h}) + i w^k (X_k - f X_{k+h}) ], X without the phi0 factor,
    // f = e^{i pi h / n} = i for shifted rings
    if (L.mmax < h) {                                  // no aliasing: X_k = p_k, X_{k+h} = -+ conj p_{h-k}; loads in flight
      constexpr int NB = 4;
      const double sgc = shifted ? -1.0 : 1.0;
      for (int t0 = threadIdx.x; t0 < h; t0 += NT * NB) {
        double2 qd[NB], qc[NB];
        bool hd[NB], hc[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int k = t0 + j * NT, m2 = h - k;
          hd[j] = hc[j] = false;
          if (k >= h) continue;
          if (k <= L.mmax) {
            const int src = L.m2src[k];
            if (src >= 0) { hd[j] = true; qd[j] = ph_half_load(p, L, src, c, L.m2im[k], pair, hemi); }
          }
          if (m2 <= L.mmax) {                          // m2 >= 1 always
            const int src = L.m2src[m2];
            if (src >= 0) { hc[j] = true; qc[j] = ph_half_load(p, L, src, c, L.m2im[m2], pair, hemi); }
          }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int k = t0 + j * NT;
          if (k >= h) continue;
          double2 Xk = make_double2(0.0, 0.0), Xh = make_double2(0.0, 0.0);
          if (hd[j]) Xk = k == 0 ? make_double2(qd[j].x, 0.0) : qd[j];
          if (hc[j]) { Xh = make_double2(sgc * qc[j].x, -sgc * qc[j].y); if (shifted) Xh = bf_mul_pi(Xh); }
          double2 y = rh_pack(Xk, Xh, rs_expipi(k, h));
          if (shifted) y = cmul(y, rs_expipi(k, n));
          work[slot(k)] = y;
        }
      }
    } else {
      for (int k = threadIdx.x; k < h; k += NT) {
        const double2 Xk = fold_bin1(p, L, c, pair, n, shifted, k, hemi);
        double2 Xh = fold_bin1(p, L, c, pair, n, shifted, k + h, hemi);
        if (shifted) Xh = bf_mul_pi(Xh);
        double2 y = rh_pack(Xk, Xh, rs_expipi(k, h));
        if (shifted) y = cmul(y, rs_expipi(k, n));
        work[slot(k)] = y;
      }
    }
    __syncthreads();
    sm_fft_dit<NT>(work, h, T);
    double *mp = map_ptr(p, c) + o;
    const double *ps = ps_ptr(p, c);
    if (ps) ps += o;
    if ((reinterpret_cast<uintptr_t>(mp) & 15) == 0 && (!ps || (reinterpret_cast<uintptr_t>(ps) & 15) == 0)) {
      double2 *mp2 = reinterpret_cast<double2 *>(mp);
      const double2 *ps2 = reinterpret_cast<const double2 *>(ps);
      for (int q = threadIdx.x; q < h; q += NT) {
        const double2 z = work[bf_pidx<true>(q)];
        double2 v = make_double2(w * z.x, w * z.y);
        if (ps) { const double2 f = ps2[q]; v.x *= f.x; v.y *= f.y; }
        if (p.add) { const double2 old = mp2[q]; v.x += old.x; v.y += old.y; }
        mp2[q] = v;
      }
    } else {
      for (int j = threadIdx.x; j < n; j += NT) {
        const double2 z = work[bf_pidx<true>(j >> 1)];
        double v = w * ((j & 1) ? z.y : z.x);
        if (ps) v *= ps[j];
        if (p.add) mp[j] += v; else mp[j] = v;
      }
    }
  } else {
    if (o < 0) {                                       // absent ring: its phases are zero
      for (int e = threadIdx.x; e < L.nm_total; e += NT)
        reinterpret_cast<double2 *>(ph_at(p, L, L.mlist_src[e], c, L.mlist_im[e], pair))[hemi] = make_double2(0.0, 0.0);
      return;
    }
    const double *mp = map_ptr(p, c) + o;
    if ((reinterpret_cast<uintptr_t>(mp) & 15) == 0) {
      const double2 *mp2 = reinterpret_cast<const double2 *>(mp);
      for (int q = threadIdx.x; q < h; q += NT) { const double2 z = mp2[q]; work[bf_pidx<true>(q)] = make_double2(w * z.x, w * z.y); }
    } else {
      for (int q = threadIdx.x; q < h; q += NT) work[bf_pidx<true>(q)] = make_double2(w * mp[2 * q], w * mp[2 * q + 1]);
    }
    __syncthreads();
    sm_fft_dif<NT>(work, h, T);
    for (int e = threadIdx.x; e < L.nm_total; e += NT) {
      const int m = L.mlist[e], b = m % n, k = b & (h - 1), k2 = (h - k) & (h - 1);
      double2 wb = rs_expipi(2LL * b, n);
      wb.y = -wb.y;
      double2 X = rh_unpack(work[slot(k)], work[slot(k2)], wb);
      const double cm = m == 0 ? 1.0 : 2.0;
      double2 f = make_double2(cm, 0.0);
      if (shifted) { const double2 t = rs_expipi(m, n); f = make_double2(cm * t.x, -cm * t.y); }
      X = cmul(X, f);
      reinterpret_cast<double2 *>(ph_at(p, L, L.mlist_src[e], c, L.mlist_im[e], pair))[hemi] = X;
    }
  }
}

// twiddles of one fused pass with leading half-size h: T[j] = exp(-i pi j / h), j < h/2 (blue_fft.cuh)
__global__ void twiddle_kernel(double2 *T, int h, int count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) { double2 e = expipi(j, h); T[j] = make_double2(e.x, -e.y); }
}

// filter spectra of the fused classes in bit-reversed order
__global__ void __launch_bounds__(256) bitrev_filter_kernel(FParams p, int first_pair, int bits, double2 *vbr) {
  const int pair = first_pair + blockIdx.x;
  const int len = p.zlen[pair];
  const size_t o = p.zbase[pair] + (size_t)p.zidx[pair] * len;
  for (int k = threadIdx.x; k < len; k += blockDim.x) vbr[o + k] = p.vtab[o + bf_bitrev((unsigned)k, bits)];
}

// ---------------------------------------------------------------- host side
static cufftHandle get_plan(sharp_geom_info *g, int region, int ncomp) {
  long long key = ((long long)region << 8) | ncomp;
  auto it = g->plans.find(key);
  if (it != g->plans.end()) return (cufftHandle)it->second;
  const FftRegion &R = g->regions[region];
  cufftHandle h;
  int n[1] = {R.len};
  CMDR_CUFFT_CHECK(cufftPlanMany(&h, 1, n, nullptr, 1, R.len, nullptr, 1, R.len, CUFFT_Z2Z, ncomp * R.np));
  g->plans[key] = (int)h;
  return h;
}

static FParams base_params(sharp_geom_info *g, int ncomp, const PhaseLayout &L, double2 *buf) {
  FParams p;
  p.npairs = g->npairs; p.ncomp = ncomp;
  p.nph = g->d_nph; p.shifted = g->d_shifted; p.zidx = g->d_zidx; p.znp = g->d_znp;
  p.zlen = g->d_zlen; p.zblue = g->d_zblue; p.zbase = g->d_zbase; p.ofsN = g->d_ofsN; p.ofsS = g->d_ofsS;
  p.wgt = g->d_wgt; p.buf = buf; p.vtab = reinterpret_cast<const double2 *>(g->d_vtab);
  p.L = L; p.ph = nullptr; p.map0 = p.map1 = p.map2 = nullptr; p.weighted = 0; p.add = 0;
  p.ps0 = p.ps1 = p.ps2 = nullptr;
  p.ring_major = phase_ring_major() ? 1 : 0;
  return p;
}

// How each FFT region (ring pairs of one transform length) is executed:
//   FUSED  whole-ring chirp-z in shared memory (work length <= 8192), blue_fused[2]_kernel
//   SPLIT  radix-4 split into four length-n/4 chirp-z transforms (ring_split_kernel): the classes whose whole-ring
//          work length (16384) does not fit in shared memory -- 75 % of the cap pixels at nside 2048
//   POW2   power-of-two rings (the belt, 4 nside <= 8192): one in-place FFT per ring pair (ring_pow2_kernel)
//   CUFFT  everything else: fold / gather -> batched cuFFT (+ chirp-z through HBM) -> scatter / unfold
// Switches for cross-checks and tuning: CMDR_SHT_FUSED_BLUE=0, CMDR_SHT_RING_SPLIT=0, CMDR_SHT_BELT_FUSED=0 send the
// respective classes to cuFFT; CMDR_SHT_SPLIT_MIN=<len> also splits the shorter classes with whole-ring length >= len.
enum RegionKind { RK_CUFFT = 0, RK_FUSED, RK_SPLIT, RK_POW2 };

static int env_or(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static int split_sub_len(int whole_len) { return whole_len / 4 < 1024 ? 1024 : whole_len / 4; }   // M of the length-n/4 chirp-z

static RegionKind region_kind(const sharp_geom_info *g, size_t r) {
  static const bool fused_on = env_or("CMDR_SHT_FUSED_BLUE", 1) != 0, split_on = env_or("CMDR_SHT_RING_SPLIT", 1) != 0,
                    pow2_on = env_or("CMDR_SHT_BELT_FUSED", 1) != 0;
  static const int split_min = env_or("CMDR_SHT_SPLIT_MIN", 16384);
  const FftRegion &R = g->regions[r];
  if (R.bluestein) {
    if (split_on && R.len >= split_min && R.len <= 16384) return RK_SPLIT;     // n <= 8192, sub work length <= 4096
    if (fused_on && R.len <= 8192) return RK_FUSED;
    return RK_CUFFT;
  }
  if (pow2_on && is_pow2(R.len) && R.len >= 1024 && R.len <= 8192) return RK_POW2;
  return RK_CUFFT;
}

// complex elements of FFT work buffer (per component) the cuFFT regions need: 0 when every region is fused
size_t ringfft_scratch_elems(const sharp_geom_info *g) {
  size_t need = 0;
  for (size_t r = 0; r < g->regions.size(); ++r)
    if (g->regions[r].np > 0 && region_kind(g, r) == RK_CUFFT)
      need = (size_t)g->regions[r].base + (size_t)g->regions[r].np * g->regions[r].len;
  return need;
}

static const double2 *twiddle_table(int M, cudaStream_t st) {
  static std::map<long long, double2 *> tabs;       // (device, M)
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  const long long key = ((long long)dev << 32) | (unsigned)M;
  auto it = tabs.find(key);
  if (it != tabs.end()) return it->second;
  double2 *t = nullptr;
  CMDR_CUDA_CHECK(cudaMalloc(&t, sizeof(double2) * (size_t)(bf_tw_total(M) + 1)));
  const int np = bf_num_passes(M);
  for (int k = 0; k < np; ++k) {
    int h, fused;
    bf_pass(M, k, &h, &fused);
    if (!fused) continue;
    twiddle_kernel<<<(h / 2 + 255) / 256, 256, 0, st>>>(t + bf_tw_offset(M, k), h, h / 2);
    count_launch();
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
  tabs[key] = t;
  return t;
}

// twiddles of the register-blocked variant (bf2_* schedule): per strided pass st entries exp(-i pi j / h), then {1}
static const double2 *twiddle_table2(int M, cudaStream_t st) {
  static std::map<long long, double2 *> tabs;       // (device, M)
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  const long long key = ((long long)dev << 32) | (unsigned)M;
  auto it = tabs.find(key);
  if (it != tabs.end()) return it->second;
  double2 *t = nullptr;
  const int ntw = bf2_tw_total(M);
  CMDR_CUDA_CHECK(cudaMalloc(&t, sizeof(double2) * (size_t)ntw));
  for (int k = 0; k < bf2_num_strided(M); ++k) {
    const int h = bf2_half(M, k), cnt = h >> (bf2_stages(M, k) - 1);
    twiddle_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(t + bf2_tw_offset(M, k), h, cnt);
    count_launch();
  }
  twiddle_kernel<<<1, 32, 0, st>>>(t + ntw - 1, 1, 1);        // exp(0) = 1
  count_launch();
  CMDR_CUDA_CHECK(cudaGetLastError());
  tabs[key] = t;
  return t;
}

// side streams of the fused classes: the classes are independent of each other and of the belt path, and the
// small ones are latency bound (few CTAs, long m chains per bin), so all of them run concurrently
static cudaStream_t class_stream(int k) {
  static std::map<long long, cudaStream_t> cs;       // (device, k)
  int dev = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  const long long key = ((long long)dev << 8) | k;
  auto it = cs.find(key);
  if (it != cs.end()) return it->second;
  cudaStream_t s;
  CMDR_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cs[key] = s;
  return s;
}

// One fused region on stream `cs` (DIR 0 synthesis, 1 analysis).
template <int DIR>
static void launch_region(sharp_geom_info *g, int ncomp, size_t r, RegionKind kind, const FParams &p, cudaStream_t cs) {
  const FftRegion &R = g->regions[r];
  const dim3 grid(R.np, ncomp);
  if (kind == RK_SPLIT) {
    static bool attr = false;
    if (!attr) {
      attr = true;
      const int smem = (int)(sizeof(double2) * (8192 + bf_padded(4096) + bf2_tw_total(4096)));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_split_kernel<DIR, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_split_kernel<DIR, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_split_kernel<DIR, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int M = split_sub_len(R.len), nmax = R.len / 2;          // rings of this class have n <= len / 2 points
    const double2 *tw = twiddle_table2(M, cs);
    const double2 *vsub = reinterpret_cast<const double2 *>(g->d_vsub[(int)r]);
    const size_t smem = sizeof(double2) * (size_t)(nmax + bf_padded(M) + bf2_tw_total(M));
    static const int nt = env_or("CMDR_SHT_SPLIT_NT", 0);
    const int NT = nt ? nt : (M >= 4096 ? 512 : (M >= 2048 ? 256 : 128));
    if (NT <= 128) ring_split_kernel<DIR, 128><<<grid, 128, smem, cs>>>(p, R.first, M, nmax, tw, vsub);
    else if (NT <= 256) ring_split_kernel<DIR, 256><<<grid, 256, smem, cs>>>(p, R.first, M, nmax, tw, vsub);
    else ring_split_kernel<DIR, 512><<<grid, 512, smem, cs>>>(p, R.first, M, nmax, tw, vsub);
  } else if (kind == RK_POW2) {
    static bool attr = false;
    if (!attr) {
      attr = true;
      const int smem = (int)(sizeof(double2) * (bf_padded(8192) + bf2_tw_total(8192)));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_pow2_kernel<DIR, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_pow2_kernel<DIR, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_pow2_kernel<DIR, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int n = R.len;
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    // CMDR_SHT_BELT_HALF=1: one ring per CTA through a transform of half the length (three CTAs per SM at n = 8192).
    // Measured at nside 2048: ring stage 6.48 ms per pair against 5.99 ms with the whole-pair kernel below (the
    // 16-byte halves of the 32-byte phase elements cost twice the sector requests, two sincospi per bin), so it is
    // off by default; kept as a second, independently written path for the tests.
    static const bool half = env_or("CMDR_SHT_BELT_HALF", 0) != 0;
    if (half && n >= 2048) {
      static bool hattr = false;
      if (!hattr) {
        hattr = true;
        const int smh = (int)(sizeof(double2) * (bf_padded(4096) + bf2_tw_total(4096)));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_half_kernel<DIR, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smh));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_half_kernel<DIR, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smh));
      }
      const int h = n / 2;
      const double2 *twh = twiddle_table2(h, cs);
      const size_t smh = sizeof(double2) * (size_t)(bf_padded(h) + bf2_tw_total(h));
      const dim3 gh(2 * R.np, ncomp);
      if (h <= 2048) ring_half_kernel<DIR, 128><<<gh, 128, smh, cs>>>(p, R.first, n, bits - 1, twh);
      else ring_half_kernel<DIR, 256><<<gh, 256, smh, cs>>>(p, R.first, n, bits - 1, twh);
      count_launch();
      return;
    }
    const double2 *tw = twiddle_table2(n, cs);
    const size_t smem = sizeof(double2) * (size_t)(bf_padded(n) + bf2_tw_total(n));
    if (n <= 2048) ring_pow2_kernel<DIR, 128><<<grid, 128, smem, cs>>>(p, R.first, n, bits, tw);
    else if (n <= 4096) ring_pow2_kernel<DIR, 256><<<grid, 256, smem, cs>>>(p, R.first, n, bits, tw);
    else ring_pow2_kernel<DIR, 512><<<grid, 512, smem, cs>>>(p, R.first, n, bits, tw);
  } else {   // RK_FUSED
    const double2 *vbr = reinterpret_cast<const double2 *>(g->d_vtab_br);
    static const bool blocked = env_or("CMDR_SHT_FFT_BLOCKED", 1) != 0;
    if (blocked) {
      static bool attr = false;
      if (!attr) {
        attr = true;
        const int smem2 = (int)(sizeof(double2) * (bf_padded(8192) + bf2_tw_total(8192)));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused2_kernel<DIR, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused2_kernel<DIR, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused2_kernel<DIR, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      }
      const double2 *tw = twiddle_table2(R.len, cs);
      const size_t smem = sizeof(double2) * (size_t)(bf_padded(R.len) + bf2_tw_total(R.len));
      if (R.len <= 2048) blue_fused2_kernel<DIR, 128><<<grid, 128, smem, cs>>>(p, R.first, R.len, tw, vbr);
      else if (R.len <= 4096) blue_fused2_kernel<DIR, 256><<<grid, 256, smem, cs>>>(p, R.first, R.len, tw, vbr);
      else blue_fused2_kernel<DIR, 512><<<grid, 512, smem, cs>>>(p, R.first, R.len, tw, vbr);
    } else {
      static bool attr = false;
      if (!attr) {
        attr = true;
        const int max_smem = (int)(sizeof(double2) * (8192 + bf_tw_total(8192)));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused_kernel<DIR, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused_kernel<DIR, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(blue_fused_kernel<DIR, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      }
      const double2 *tw = twiddle_table(R.len, cs);
      const size_t smem = sizeof(double2) * (size_t)(R.len + bf_tw_total(R.len));
      if (R.len <= 2048) blue_fused_kernel<DIR, 256><<<grid, 256, smem, cs>>>(p, R.first, R.len, tw, vbr);
      else if (R.len <= 4096) blue_fused_kernel<DIR, 512><<<grid, 512, smem, cs>>>(p, R.first, R.len, tw, vbr);
      else blue_fused_kernel<DIR, 1024><<<grid, 1024, smem, cs>>>(p, R.first, R.len, tw, vbr);
    }
  }
  count_launch();
}

// Launches every fused region on its own side stream forked from `st` (the classes are independent, and the short
// ones are latency bound, so they overlap), longest class first; returns the number of streams to join.
template <int DIR>
static int launch_fused_regions(sharp_geom_info *g, int ncomp, const FParams &p, cudaStream_t st) {
  std::vector<size_t> todo;
  for (size_t r = g->regions.size(); r-- > 0;)
    if (g->regions[r].np > 0 && region_kind(g, r) != RK_CUFFT) todo.push_back(r);
  if (todo.empty()) return 0;
  cudaEvent_t e0 = pooled_event(150);
  CMDR_CUDA_CHECK(cudaEventRecord(e0, st));
  int ns = 0;
  for (size_t r : todo) {
    cudaStream_t cs = class_stream(ns);
    CMDR_CUDA_CHECK(cudaStreamWaitEvent(cs, e0, 0));
    launch_region<DIR>(g, ncomp, r, region_kind(g, r), p, cs);
    CMDR_CUDA_CHECK(cudaEventRecord(pooled_event(151 + ns), cs));
    ++ns;
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
  return ns;
}
static void join_fused(int ns, cudaStream_t st) {
  for (int k = 0; k < ns; ++k) CMDR_CUDA_CHECK(cudaStreamWaitEvent(st, pooled_event(151 + k), 0));
}

static void run_fft_region(sharp_geom_info *g, int ncomp, size_t r, double2 *buf, int direct_dir, FParams &p, cudaStream_t st) {
  const FftRegion &R = g->regions[r];
  cufftHandle h = get_plan(g, (int)r, ncomp);
  CMDR_CUFFT_CHECK(cufftSetStream(h, st));
  cufftDoubleComplex *d = reinterpret_cast<cufftDoubleComplex *>(buf + (size_t)ncomp * R.base);
  if (!R.bluestein) {
    CMDR_CUFFT_CHECK(cufftExecZ2Z(h, d, d, direct_dir));
    count_launch();
  } else {
    CMDR_CUFFT_CHECK(cufftExecZ2Z(h, d, d, CUFFT_FORWARD));
    blue_mul_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
    CMDR_CUFFT_CHECK(cufftExecZ2Z(h, d, d, CUFFT_INVERSE));
    count_launch(3);
  }
}

// Tables built once per geometry: chirp-z filter spectra of the cuFFT and whole-ring fused classes (natural order, and
// bit-reversed for the fused kernels) and of the sub-transforms of the split classes; twiddle tables of every length.
static void ensure_vtab(sharp_geom_info *g, cudaStream_t st) {
  if (g->vtab_ready) return;
  g->vtab_ready = true;
  PhaseLayout L;
  FParams p = base_params(g, 1, L, nullptr);
  size_t vneed = 0, brneed = 0;
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (!R.bluestein || R.np == 0) continue;
    const RegionKind k = region_kind(g, r);
    if (k == RK_CUFFT || k == RK_FUSED) vneed = (size_t)R.base + (size_t)R.np * R.len;
    if (k == RK_FUSED) brneed = (size_t)R.base + (size_t)R.np * R.len;
  }
  if (vneed) CMDR_CUDA_CHECK(cudaMalloc(&g->d_vtab, sizeof(double2) * vneed));
  if (brneed) CMDR_CUDA_CHECK(cudaMalloc(&g->d_vtab_br, sizeof(double2) * brneed));
  p.vtab = reinterpret_cast<const double2 *>(g->d_vtab);
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (R.np == 0) continue;
    const RegionKind k = region_kind(g, r);
    if (k == RK_POW2) { twiddle_table2(R.len, st); continue; }
    if (!R.bluestein) continue;
    if (k == RK_SPLIT) {
      const int M = split_sub_len(R.len);
      double2 *vs = nullptr;
      CMDR_CUDA_CHECK(cudaMalloc(&vs, sizeof(double2) * (size_t)R.np * M));
      g->d_vsub[(int)r] = reinterpret_cast<double *>(vs);
      const double2 *tw = twiddle_table2(M, st);
      static bool attr = false;
      if (!attr) {
        attr = true;
        CMDR_CUDA_CHECK(cudaFuncSetAttribute(ring_split_filter_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(double2) * (bf_padded(4096) + bf2_tw_total(4096)))));
      }
      ring_split_filter_kernel<256><<<R.np, 256, sizeof(double2) * (size_t)(bf_padded(M) + bf2_tw_total(M)), st>>>(p, R.first, M, tw, vs);
      count_launch();
      continue;
    }
    blue_filter_kernel<<<R.np, 256, 0, st>>>(p, R.first, reinterpret_cast<double2 *>(g->d_vtab));
    cufftHandle h = get_plan(g, (int)r, 1);
    CMDR_CUFFT_CHECK(cufftSetStream(h, st));
    cufftDoubleComplex *d = reinterpret_cast<cufftDoubleComplex *>(g->d_vtab) + R.base;
    CMDR_CUFFT_CHECK(cufftExecZ2Z(h, d, d, CUFFT_FORWARD));
    count_launch(2);
    if (k == RK_FUSED) {     // the fused kernels read the spectrum in bit-reversed order
      int bits = 0;
      while ((1 << bits) < R.len) ++bits;
      bitrev_filter_kernel<<<R.np, 256, 0, st>>>(p, R.first, bits, reinterpret_cast<double2 *>(g->d_vtab_br));
      count_launch();
      if (env_or("CMDR_SHT_FFT_BLOCKED", 1) != 0) twiddle_table2(R.len, st); else twiddle_table(R.len, st);
    }
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
}

void ringfft_synth(sharp_geom_info *g, int ncomp, const PhaseLayout &L, const double4 *ph,
                   double *const *map, bool weighted, bool add, cudaStream_t st, const double *const *pixscale) {
  if (g->npairs == 0) return;
  ensure_vtab(g, st);
  const size_t zneed = ringfft_scratch_elems(g);
  double2 *buf = zneed ? static_cast<double2 *>(scratch_get("fftbuf", sizeof(double2) * zneed * ncomp)) : nullptr;
  FParams p = base_params(g, ncomp, L, buf);
  p.ph = const_cast<double4 *>(ph);
  p.map0 = map[0]; p.map1 = ncomp > 1 ? map[1] : nullptr; p.map2 = ncomp > 2 ? map[2] : nullptr;
  if (pixscale) { p.ps0 = pixscale[0]; p.ps1 = ncomp > 1 ? pixscale[1] : nullptr; p.ps2 = ncomp > 2 ? pixscale[2] : nullptr; }
  p.weighted = weighted; p.add = add;
  const int nside_streams = launch_fused_regions<0>(g, ncomp, p, st);
  for (size_t r = 0; r < g->regions.size(); ++r) {             // what is left goes through cuFFT on the caller's stream
    const FftRegion &R = g->regions[r];
    if (R.np == 0 || region_kind(g, r) != RK_CUFFT) continue;
    if (!R.bluestein && R.len >= 2 * L.mmax + 1) {             // no aliasing: coalesced transpose
      CMDR_CUDA_CHECK(cudaMemsetAsync(buf + (size_t)ncomp * R.base, 0, sizeof(double2) * (size_t)ncomp * R.np * R.len, st));
      dim3 grid((R.np + 31) / 32, (L.nm_total + 31) / 32, ncomp);
      fold_transpose_kernel<<<grid, 256, 0, st>>>(p, R.first, R.np, R.len);
    } else {
      fold_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
    }
    run_fft_region(g, ncomp, r, buf, CUFFT_INVERSE, p, st);
    scatter_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
    count_launch(2);
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
  join_fused(nside_streams, st);
}

void ringfft_anal(sharp_geom_info *g, int ncomp, const PhaseLayout &L, double4 *ph,
                  const double *const *map, bool weighted, cudaStream_t st) {
  if (g->npairs == 0) return;
  ensure_vtab(g, st);
  const size_t zneed = ringfft_scratch_elems(g);
  double2 *buf = zneed ? static_cast<double2 *>(scratch_get("fftbuf", sizeof(double2) * zneed * ncomp)) : nullptr;
  FParams p = base_params(g, ncomp, L, buf);
  p.ph = ph;
  p.map0 = const_cast<double *>(map[0]);
  p.map1 = ncomp > 1 ? const_cast<double *>(map[1]) : nullptr;
  p.map2 = ncomp > 2 ? const_cast<double *>(map[2]) : nullptr;
  p.weighted = weighted;
  const int nside_streams = launch_fused_regions<1>(g, ncomp, p, st);
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (R.np == 0 || region_kind(g, r) != RK_CUFFT) continue;
    gather_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
    run_fft_region(g, ncomp, r, buf, CUFFT_FORWARD, p, st);
    unfold_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
    count_launch(2);
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
  join_fused(nside_streams, st);
}

void destroy_plans(sharp_geom_info *g) {
  for (auto &kv : g->plans) cufftDestroy((cufftHandle)kv.second);
  g->plans.clear();
}

}  // namespace cmdr
