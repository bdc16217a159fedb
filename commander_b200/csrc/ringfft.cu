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

#include "kernels.h"

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
  int weighted, add;
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
__device__ __forceinline__ double *map_ptr(const FParams &p, int c) {
  return c == 0 ? p.map0 : (c == 1 ? p.map1 : p.map2);
}

// ---- synthesis, before the FFT: fold phases into the spectrum Z = X_north + i X_south
// Generic gather form (any ring length, aliasing allowed): bin k sums the m == +-k (mod n).
// Short rings have few bins and long m-chains, so the chain of each bin is split over
// `parts` threads and combined through shared memory.
__device__ __forceinline__ double2 fold_one(const FParams &p, const PhaseLayout &L, int c, int pair, int n,
                                            bool shifted, int m, bool conj_term) {
  int src = L.m2src[m];
  if (src < 0) return make_double2(0.0, 0.0);
  double4 q = p.ph[((size_t)(src * L.ncomp_tot + L.comp0 + c) * L.NML + L.m2im[m]) * L.NPL + L.pair0 + pair];
  double2 e = shifted ? expipi(m, n) : make_double2(1.0, 0.0);
  double2 pn = cmul(make_double2(q.x, q.y), e), ps = cmul(make_double2(q.z, q.w), e);
  if (conj_term) return make_double2(pn.x + ps.y, -pn.y + ps.x);       // conj p_m
  if (m == 0) return make_double2(pn.x, ps.x);
  return make_double2(pn.x - ps.y, pn.y + ps.x);
}

__global__ void __launch_bounds__(256) fold_kernel(FParams p, int first_pair) {
  __shared__ double2 part_sum[256];
  const int pair = first_pair + blockIdx.x, c = blockIdx.y;
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair], shifted = p.shifted[pair];
  double2 *out = p.buf + zoff(p, pair, c);
  const PhaseLayout &L = p.L;
  int parts = 1;
  while (parts * 2 * n <= 256) parts *= 2;          // power of two, parts * n <= 256
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
    for (int kk = threadIdx.x; kk < len; kk += blockDim.x) {
      double2 tot = make_double2(0.0, 0.0);
      if (kk < n) {
        for (int q = 0; q < parts; ++q) { tot.x += part_sum[kk + q * n].x; tot.y += part_sum[kk + q * n].y; }
        if (blue) tot = cmul(tot, expipi((long long)kk * kk, n));
      }
      out[kk] = tot;
    }
    return;
  }
  for (int k = threadIdx.x; k < len; k += blockDim.x) {
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
    out[k] = acc;
  }
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
      q = p.ph[((size_t)(L.mlist_src[e] * L.ncomp_tot + L.comp0 + c) * L.NML + L.mlist_im[e]) * L.NPL + L.pair0 + first_pair + pr];
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
__global__ void __launch_bounds__(256) scatter_kernel(FParams p) {
  const int pair = blockIdx.x, c = blockIdx.y;
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair];
  const double2 *in = p.buf + zoff(p, pair, c);
  double *mp = map_ptr(p, c);
  const long long oN = p.ofsN[pair], oS = p.ofsS[pair];
  double w = p.weighted ? p.wgt[pair] : 1.0;
  if (blue) w /= (double)len;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double2 z = in[j];
    if (blue) z = cmul(z, expipi((long long)j * j, n));
    if (oN >= 0) { if (p.add) mp[oN + j] += w * z.x; else mp[oN + j] = w * z.x; }
    if (oS >= 0) { if (p.add) mp[oS + j] += w * z.y; else mp[oS + j] = w * z.y; }
  }
}

// ---- analysis, before the FFT: z = w (x_north + i x_south)
__global__ void __launch_bounds__(256) gather_kernel(FParams p) {
  const int pair = blockIdx.x, c = blockIdx.y;
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair];
  double2 *out = p.buf + zoff(p, pair, c);
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
    out[j] = z;
  }
}

// ---- analysis, after the FFT: phases ph_m = c_m e^{-i m phi0} X_{m mod n}
__global__ void __launch_bounds__(256) unfold_kernel(FParams p) {
  const int pair = blockIdx.x, c = blockIdx.y;
  const int n = p.nph[pair], len = p.zlen[pair];
  const bool blue = p.zblue[pair], shifted = p.shifted[pair];
  const double2 *in = p.buf + zoff(p, pair, c);
  const PhaseLayout &L = p.L;
  const double inv = blue ? 1.0 / (double)len : 1.0;
  for (int e = threadIdx.x; e < L.nm_total; e += blockDim.x) {
    const int m = L.mlist[e];
    const int k = m % n, k2 = (n - k) % n;
    double2 a = in[k], b = in[k2];
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
    p.ph[((size_t)(L.mlist_src[e] * L.ncomp_tot + L.comp0 + c) * L.NML + L.mlist_im[e]) * L.NPL + L.pair0 + pair] =
        make_double4(xn.x, xn.y, xs.x, xs.y);
  }
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
  return p;
}

static void run_ffts(sharp_geom_info *g, int ncomp, double2 *buf, int direct_dir, FParams &p, cudaStream_t st) {
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (R.np == 0) continue;
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
  CMDR_CUDA_CHECK(cudaGetLastError());
}

static void ensure_vtab(sharp_geom_info *g, cudaStream_t st) {
  if (g->vtab_ready) return;
  g->vtab_ready = true;
  if (g->vlen_total == 0) return;
  CMDR_CUDA_CHECK(cudaMalloc(&g->d_vtab, sizeof(double2) * (size_t)g->vlen_total));
  PhaseLayout L;
  FParams p = base_params(g, 1, L, nullptr);
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (!R.bluestein || R.np == 0) continue;
    blue_filter_kernel<<<R.np, 256, 0, st>>>(p, R.first, reinterpret_cast<double2 *>(g->d_vtab));
    cufftHandle h = get_plan(g, (int)r, 1);
    CMDR_CUFFT_CHECK(cufftSetStream(h, st));
    cufftDoubleComplex *d = reinterpret_cast<cufftDoubleComplex *>(g->d_vtab) + R.base;
    CMDR_CUFFT_CHECK(cufftExecZ2Z(h, d, d, CUFFT_FORWARD));
    count_launch(2);
  }
  CMDR_CUDA_CHECK(cudaGetLastError());
}

void ringfft_synth(sharp_geom_info *g, int ncomp, const PhaseLayout &L, const double4 *ph,
                   double *const *map, bool weighted, bool add, cudaStream_t st) {
  if (g->npairs == 0) return;
  ensure_vtab(g, st);
  double2 *buf = static_cast<double2 *>(scratch_get("fftbuf", sizeof(double2) * (size_t)g->zlen_total * ncomp));
  FParams p = base_params(g, ncomp, L, buf);
  p.ph = const_cast<double4 *>(ph);
  p.map0 = map[0]; p.map1 = ncomp > 1 ? map[1] : nullptr; p.map2 = ncomp > 2 ? map[2] : nullptr;
  p.weighted = weighted; p.add = add;
  // all pairs that need the generic (aliasing) fold go in ONE launch: the blocks of short rings are
  // latency bound (long m chains per bin) and must overlap the big ones instead of queueing
  int gen_first = -1, gen_np = 0;
  for (size_t r = 0; r < g->regions.size(); ++r) {
    const FftRegion &R = g->regions[r];
    if (R.np == 0) continue;
    if (!R.bluestein && R.len >= 2 * L.mmax + 1) {      // no aliasing: coalesced transpose
      CMDR_CUDA_CHECK(cudaMemsetAsync(buf + (size_t)ncomp * R.base, 0, sizeof(double2) * (size_t)ncomp * R.np * R.len, st));
      dim3 grid((R.np + 31) / 32, (L.nm_total + 31) / 32, ncomp);
      fold_transpose_kernel<<<grid, 256, 0, st>>>(p, R.first, R.np, R.len);
      count_launch();
    } else if (gen_first < 0 || R.first == gen_first + gen_np) {
      if (gen_first < 0) gen_first = R.first;
      gen_np += R.np;
    } else {                                             // non-contiguous (not produced by build_regions)
      fold_kernel<<<dim3(R.np, ncomp), 256, 0, st>>>(p, R.first);
      count_launch();
    }
  }
  if (gen_np > 0) {
    fold_kernel<<<dim3(gen_np, ncomp), 256, 0, st>>>(p, gen_first);
    count_launch();
  }
  run_ffts(g, ncomp, buf, CUFFT_INVERSE, p, st);
  scatter_kernel<<<dim3(g->npairs, ncomp), 256, 0, st>>>(p);
  count_launch();
  CMDR_CUDA_CHECK(cudaGetLastError());
}

void ringfft_anal(sharp_geom_info *g, int ncomp, const PhaseLayout &L, double4 *ph,
                  const double *const *map, bool weighted, cudaStream_t st) {
  if (g->npairs == 0) return;
  ensure_vtab(g, st);
  double2 *buf = static_cast<double2 *>(scratch_get("fftbuf", sizeof(double2) * (size_t)g->zlen_total * ncomp));
  FParams p = base_params(g, ncomp, L, buf);
  p.ph = ph;
  p.map0 = const_cast<double *>(map[0]);
  p.map1 = ncomp > 1 ? const_cast<double *>(map[1]) : nullptr;
  p.map2 = ncomp > 2 ? const_cast<double *>(map[2]) : nullptr;
  p.weighted = weighted;
  gather_kernel<<<dim3(g->npairs, ncomp), 256, 0, st>>>(p);
  count_launch();
  run_ffts(g, ncomp, buf, CUFFT_FORWARD, p, st);
  unfold_kernel<<<dim3(g->npairs, ncomp), 256, 0, st>>>(p);
  count_launch();
  CMDR_CUDA_CHECK(cudaGetLastError());
}

void destroy_plans(sharp_geom_info *g) {
  for (auto &kv : g->plans) cufftDestroy((cufftHandle)kv.second);
  g->plans.clear();
}

}  // namespace cmdr
