// conviqt.cu -- the convolution cube of comm_conviqt (commander3/src/comm_conviqt_mod.f90).
//
// precompute_sky (:207-292) builds, for one sky and one beam,
//     c(pix, psi_k) = sum_{j=-bmax}^{bmax} M_j(pix) e^{i j psi_k},   psi_k = 2 pi k / (2 bmax),
// where M_0 is a spin-0 synthesis and (M_j, M_-j) the two components of a spin-j synthesis of
// sky x beam coefficient products (get_alms, :294-357).  The reference runs bmax+1 sharp_execute
// calls through host arrays, then one FFTW c2r of length 2*bmax per pixel on the host, and stores the
// cube in single precision.  Here everything between the sky a_lm and the cube stays in HBM:
//
//   conviqt_alms_kernel   sky a_lm, beam b_l,j  ->  the (1 or 2) a_lm columns of beam index j
//   spin-j synthesis      the library's own Legendre + ring-FFT stages (any registered comm)
//   conviqt_psi_kernel    M[-bmax..bmax][pix] (FP64)  ->  cube[psi][pix] (float or double)
//
// The psi transform is a direct real DFT per pixel (2*bmax <= 64 points): each thread owns one pixel,
// parks its 2*bmax+1 inputs in a private shared-memory column (coalesced loads, no barrier) and
// writes 2*bmax coalesced outputs, four planes per inner loop (the planes k, n-k, k+bmax, bmax-k share their
// products).  Algorithmic bytes per pixel: 8 (2 bmax + 1) read + 4 * 2 bmax written.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

using namespace cmdr;

namespace {

constexpr int PSI_BLOCK = 128;

// get_alms, commander3/src/comm_conviqt_mod.f90:294-357, one thread per local (l, m >= 0).
// sky: the nmaps columns of the local real-packed a_lm; beam: alm_beam%a(nmaps, nalm_tot) as the
// reference holds it (:94-115), single-precision complex at 1-based index l(l+1)/2 + m + 1.
__global__ void conviqt_alms_kernel(int j, int nmaps, int lmax, const int *__restrict__ mval,
                                    const long long *__restrict__ mvstart, const double *__restrict__ s0,
                                    const double *__restrict__ s1, const double *__restrict__ s2,
                                    const float2 *__restrict__ beam, double *__restrict__ o0,
                                    double *__restrict__ o1) {
  const int im = blockIdx.y;
  const int m = mval[im];
  const int l = m + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const long long mvs = mvstart[im];
  const long long i = (m == 0) ? mvs + l : mvs + 2 * (long long)l;
  const double rsqrt2 = 0.70710678118654752440, sqrt2 = 1.41421356237309504880;
  double pr = 0.0, pi = 0.0, nr = 0.0, ni = 0.0;
  if (l >= j) {                                              // :311
    const double spinsign = j ? -1.0 : 1.0;                  // :303
    const double mfac = (m & 1) ? -1.0 : 1.0;                // :312
    const double lnorm = 0.5 * sqrt(4.0 * M_PI / (2.0 * l + 1.0));   // :96-98
    const float2 *b = beam + ((long long)l * (l + 1) / 2 + j) * nmaps;   // :314
    const double *s[3] = {s0, s1, s2};
    double v1r = 0.0, v1i = 0.0, v2r = 0.0, v2i = 0.0;
    for (int c = 0; c < nmaps; ++c) {
      double sr, si;                                         // get_alm_TEB, comm_map_mod.f90:1523-1546
      if (m == 0) { sr = s[c][i]; si = 0.0; }
      else { sr = rsqrt2 * s[c][i]; si = rsqrt2 * s[c][i + 1]; }
      const double br = (double)b[c].x, bi = (double)b[c].y;
      v1r += sr * br - si * bi; v1i += sr * bi + si * br;    // sum(alm_s * alm_b)            :318
      v2r += sr * br + si * bi; v2i += sr * bi - si * br;    // sum(conjg(alm_s) * alm_b)     :323
    }
    v2r *= mfac; v2i *= mfac;
    const double cr = v2r * mfac, ci = -v2i * mfac;          // conjg(v2) * mfac
    const double f = spinsign * lnorm;
    pr = f * (v1r + cr); pi = f * (v1i + ci);                // positive spin                :326
    const double xr = f * (v1r - cr), xi = f * (v1i - ci);   // -i * (xr + i xi) = xi - i xr  :336
    nr = xi; ni = -xr;
  }
  if (m == 0) {
    o0[i] = pr;
    if (o1) o1[i] = nr;
  } else {
    o0[i] = pr * sqrt2; o0[i + 1] = pi * sqrt2;
    if (o1) { o1[i] = nr * sqrt2; o1[i + 1] = ni * sqrt2; }
  }
}

// The psi transform of precompute_sky, commander3/src/comm_conviqt_mod.f90:267-282: FFTW c2r of
// dv(0) = M_0, dv(j) = (M_j, M_-j); backward sign, unnormalised, imaginary parts of dv(0) and of the
// Nyquist entry dv(bmax) ignored:
//   c_k = M_0 + (-1)^k M_bmax + 2 sum_{j=1}^{bmax-1} (M_j cos(2 pi j k / n) - M_-j sin(2 pi j k / n)),  n = 2 bmax.
// marr rows: 0 -> M_0, 2j-1 -> M_j, 2j -> M_-j.
template <typename OUT>
__global__ void conviqt_psi_kernel(int bmax, long long np, const double *__restrict__ marr, OUT *__restrict__ cube) {
  extern __shared__ double sm[];
  const int n = 2 * bmax;
  double *twc = sm, *tws = sm + n, *col = sm + 2 * n + threadIdx.x;
  for (int t = threadIdx.x; t < n; t += blockDim.x) sincospi(2.0 * t / n, &tws[t], &twc[t]);
  __syncthreads();
  for (long long p0 = (long long)blockIdx.x * PSI_BLOCK; p0 < np; p0 += (long long)gridDim.x * PSI_BLOCK) {
    const long long p = p0 + threadIdx.x;
    if (p >= np) continue;
    for (int r = 0; r <= n; ++r) col[r * PSI_BLOCK] = marr[(long long)r * np + p];
    const double m0 = col[0], mb = col[(n - 1) * PSI_BLOCK];
    // Planes k, n-k, k+bmax and bmax-k share their products: cos(2 pi j (n-k)/n) = cos(2 pi j k/n), the sine flips
    // sign, and a shift by bmax multiplies both by (-1)^j.  Four partial sums (even / odd j, cosine / sine part)
    // per k in [0, bmax/2] give all four planes: a quarter of the shared-memory loads of the plain double loop.
    for (int k = 0; k <= bmax / 2; ++k) {
      double ce = 0.0, co = 0.0, se = 0.0, so = 0.0;
      int t = 0;
      for (int j = 1; j < bmax; j += 2) {
        t += k; if (t >= n) t -= n;                          // t = j k mod n, j odd
        co = fma(col[(2 * j - 1) * PSI_BLOCK], twc[t], co);
        so = fma(col[(2 * j) * PSI_BLOCK], tws[t], so);
        if (j + 1 < bmax) {
          t += k; if (t >= n) t -= n;                        // j + 1, even
          ce = fma(col[(2 * j + 1) * PSI_BLOCK], twc[t], ce);
          se = fma(col[(2 * j + 2) * PSI_BLOCK], tws[t], se);
        }
      }
      const double nyq0 = (k & 1) ? -mb : mb, nyq1 = ((k + bmax) & 1) ? -mb : mb;
      const double c0 = ce + co, s0 = se + so, c1 = ce - co, s1 = se - so;
      const int k1 = k + bmax, k2 = (n - k) % n, k3 = bmax - k;
      cube[(long long)k * np + p] = (OUT)(m0 + nyq0 + 2.0 * (c0 - s0));
      cube[(long long)k2 * np + p] = (OUT)(m0 + nyq0 + 2.0 * (c0 + s0));
      cube[(long long)k1 * np + p] = (OUT)(m0 + nyq1 + 2.0 * (c1 - s1));
      cube[(long long)k3 * np + p] = (OUT)(m0 + nyq1 + 2.0 * (c1 + s1));
    }
  }
}

}  // namespace

extern "C" {

void cmdr_sht_conviqt_cube(int comm, int nmaps, int bmax, const double *const *sky_alm, const float *beam,
                           const sharp_geom_info *geom_T, const sharp_alm_info *alm_info, void *cube,
                           int cube_f64, void *stream) {
  if (nmaps < 1 || nmaps > 3) { fprintf(stderr, "cmdr_sht_conviqt_cube: nmaps %d unsupported (1..3)\n", nmaps); abort(); }
  if (bmax < 1 || bmax > CMDR_MAX_SPIN) {
    fprintf(stderr, "cmdr_sht_conviqt_cube: bmax %d outside 1..%d\n", bmax, CMDR_MAX_SPIN);
    abort();
  }
  sharp_alm_info *a = const_cast<sharp_alm_info *>(alm_info);
  if (!a->real_packed) { fprintf(stderr, "cmdr_sht_conviqt_cube: needs a real-packed alm_info\n"); abort(); }
  cudaStream_t st = (cudaStream_t)stream;
  ensure_alm_device(a);
  const long long nalm = a->nalm, np = geom_T->npix;
  const int lmax = a->lmax, n = 2 * bmax;
  const long long ntri = (long long)(lmax + 1) * (lmax + 2) / 2;

  // inputs to the device (no copy when the caller already keeps them there)
  const double *sd[3] = {nullptr, nullptr, nullptr};
  bool sky_on_dev = true;
  for (int c = 0; c < nmaps; ++c) sky_on_dev = sky_on_dev && (nalm == 0 || is_device_ptr(sky_alm[c]));
  if (sky_on_dev) {
    for (int c = 0; c < nmaps; ++c) sd[c] = sky_alm[c];
  } else {
    double *buf = static_cast<double *>(scratch_get("cvq_sky", sizeof(double) * (size_t)std::max<long long>(1, nalm) * nmaps));
    for (int c = 0; c < nmaps; ++c) {
      if (nalm) CMDR_CUDA_CHECK(cudaMemcpyAsync(buf + (size_t)c * nalm, sky_alm[c], sizeof(double) * nalm, cudaMemcpyHostToDevice, st));
      sd[c] = buf + (size_t)c * nalm;
    }
  }
  const float2 *bd = reinterpret_cast<const float2 *>(beam);
  if (!is_device_ptr(beam)) {
    float2 *buf = static_cast<float2 *>(scratch_get("cvq_beam", sizeof(float2) * (size_t)ntri * nmaps));
    CMDR_CUDA_CHECK(cudaMemcpyAsync(buf, beam, sizeof(float2) * (size_t)ntri * nmaps, cudaMemcpyHostToDevice, st));
    bd = buf;
  }
  double *aj = static_cast<double *>(scratch_get("cvq_alm", sizeof(double) * (size_t)std::max<long long>(1, nalm) * 2));
  double *marr = static_cast<double *>(scratch_get("cvq_marr", sizeof(double) * (size_t)std::max<long long>(1, np) * (n + 1)));

  for (int j = 0; j <= bmax; ++j) {
    double *ad[2] = {aj, aj + nalm};
    if (a->nm > 0) {
      dim3 grid((unsigned)((lmax + 128) / 128), (unsigned)a->nm);
      conviqt_alms_kernel<<<grid, 128, 0, st>>>(j, nmaps, lmax, a->d_mval, a->d_mvstart, sd[0], sd[1], sd[2], bd,
                                                ad[0], j ? ad[1] : nullptr);
      CMDR_CUDA_CHECK(cudaGetLastError());
      count_launch(1);
    }
    // :247-252 (j = 0, spin 0, one column) and :254-272 (spin j, two columns: M_j, M_-j)
    double *md[2] = {marr + (size_t)(j ? 2 * j - 1 : 0) * np, marr + (size_t)(2 * j) * np};
    cmdr_sht_execute_dist(comm, SHARP_Y, j, ad, md, geom_T, alm_info, SHARP_DP, st);
  }

  if (np > 0) {
    const size_t esz = cube_f64 ? sizeof(double) : sizeof(float);
    const bool cube_on_dev = is_device_ptr(cube);
    void *cd = cube_on_dev ? cube : scratch_get("cvq_cube", esz * (size_t)np * n);
    int dev = 0, nsm = 148;
    CMDR_CUDA_CHECK(cudaGetDevice(&dev));
    CMDR_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sizeof(double) * ((size_t)2 * n + (size_t)(n + 1) * PSI_BLOCK);
    const long long tiles = (np + PSI_BLOCK - 1) / PSI_BLOCK;
    const int grid = (int)std::min<long long>(tiles, (long long)nsm * 8);
    if (cube_f64) {
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(conviqt_psi_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conviqt_psi_kernel<double><<<grid, PSI_BLOCK, smem, st>>>(bmax, np, marr, static_cast<double *>(cd));
    } else {
      CMDR_CUDA_CHECK(cudaFuncSetAttribute(conviqt_psi_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conviqt_psi_kernel<float><<<grid, PSI_BLOCK, smem, st>>>(bmax, np, marr, static_cast<float *>(cd));
    }
    CMDR_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    if (!cube_on_dev) CMDR_CUDA_CHECK(cudaMemcpyAsync(cube, cd, esz * (size_t)np * n, cudaMemcpyDeviceToHost, st));
  }
  // host inputs may be rewritten and host outputs read by the caller after return
  if (!sky_on_dev || !is_device_ptr(beam) || (np > 0 && !is_device_ptr(cube))) CMDR_CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // extern "C"
