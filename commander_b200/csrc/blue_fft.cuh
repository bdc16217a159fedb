// blue_fft.cuh -- power-of-two FFT pair that works in place on one array (shared memory on the device),
// used by the fused chirp-z (Bluestein) ring kernels of ringfft.cu.
//
// A circular convolution needs  u -> IFFT( FFT(u) .* V ).  With a decimation-in-frequency forward transform
// (natural order in, bit-reversed order out), V stored in bit-reversed order and a decimation-in-time inverse
// (bit-reversed in, natural out) no reordering pass is needed at all.  Two radix-2 stages are fused per pass
// (four elements in registers, one shared-memory round trip per two stages); an odd stage count ends (DIF) or
// starts (DIT) with one plain radix-2 pass.
//
// The functions process ONE work item; the caller distributes items over threads and separates passes with a
// barrier.  Written without CUDA-only constructs so that tests/host_emul can run the same code on the host
// (define BLUE_FFT_HOST before including).
#pragma once

#ifdef BLUE_FFT_HOST
struct double2 { double x, y; };
#define BF_HD inline
#else
#include <cuda_runtime.h>
#define BF_HD __host__ __device__ __forceinline__
#endif

namespace cmdr {

BF_HD double2 bf_add(double2 a, double2 b) { double2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
BF_HD double2 bf_sub(double2 a, double2 b) { double2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
BF_HD double2 bf_mul(double2 a, double2 b) { double2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
BF_HD double2 bf_mulc(double2 a, double2 b) { double2 r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r; }   // a conj(b)
BF_HD double2 bf_mul_mi(double2 a) { double2 r; r.x = a.y; r.y = -a.x; return r; }   // -i a
BF_HD double2 bf_mul_pi(double2 a) { double2 r; r.x = -a.y; r.y = a.x; return r; }   // +i a

// Twiddles are stored pass-major: the fused pass with leading half-size h owns the h/2 contiguous entries
// T[j] = exp(-2 pi i j / (2h)), j < h/2 (bf_tw_offset gives its start; M/3 entries in total), so that consecutive
// work items read consecutive entries.  The second stage's factor exp(-2 pi i j / h) is T[j]^2.  The single
// radix-2 pass that ends an odd stage count has h = 1 and needs no twiddle.

// DIF stages of half-sizes h and h/2 on the four elements of work item q in [0, M/4); h >= 2.
BF_HD void dif_item4(double2 *x, int h, const double2 *T, int q) {
  const int hh = h >> 1;
  const int j = q & (hh - 1);
  const int i = (q - j) * 4 + j;                  // block (q / hh) of 2h elements, offset j
  const double2 w1 = T[j], w2 = bf_mul(w1, w1);
  const double2 a0 = x[i], a1 = x[i + hh], a2 = x[i + h], a3 = x[i + h + hh];
  const double2 b0 = bf_add(a0, a2), b2 = bf_mul(bf_sub(a0, a2), w1);
  const double2 b1 = bf_add(a1, a3), b3 = bf_mul_mi(bf_mul(bf_sub(a1, a3), w1));
  x[i] = bf_add(b0, b1);
  x[i + hh] = bf_mul(bf_sub(b0, b1), w2);
  x[i + h] = bf_add(b2, b3);
  x[i + h + hh] = bf_mul(bf_sub(b2, b3), w2);
}

// the last DIF stage (h = 1) on work item q in [0, M/2)
BF_HD void dif_item2(double2 *x, int q) {
  const double2 a = x[2 * q], b = x[2 * q + 1];
  x[2 * q] = bf_add(a, b);
  x[2 * q + 1] = bf_sub(a, b);
}

// inverse (DIT, conjugate twiddles) of dif_item4: stages h/2 then h
BF_HD void dit_item4(double2 *x, int h, const double2 *T, int q) {
  const int hh = h >> 1;
  const int j = q & (hh - 1);
  const int i = (q - j) * 4 + j;
  const double2 w1 = T[j], w2 = bf_mul(w1, w1);
  const double2 c0 = x[i], c1 = bf_mulc(x[i + hh], w2), c2 = x[i + h], c3 = bf_mulc(x[i + h + hh], w2);
  const double2 b0 = bf_add(c0, c1), b1 = bf_sub(c0, c1);
  const double2 b2 = bf_mulc(bf_add(c2, c3), w1), b3 = bf_mul_pi(bf_mulc(bf_sub(c2, c3), w1));
  x[i] = bf_add(b0, b2);
  x[i + h] = bf_sub(b0, b2);
  x[i + hh] = bf_add(b1, b3);
  x[i + h + hh] = bf_sub(b1, b3);
}

BF_HD void dit_item2(double2 *x, int q) { dif_item2(x, q); }   // h = 1: the butterfly is its own inverse (up to 2)

// Pass schedule.  DIF: h = M/2, M/8, ... fused while h >= 2, then a single stage when h == 1 remains.
// pass k of n: *h = leading half-size, *fused = two stages in this pass.  DIT runs the same list backwards.
BF_HD int bf_num_passes(int M) {
  int n = 0;
  for (int h = M >> 1; h >= 1; h >>= 2) ++n;
  return n;
}
BF_HD void bf_pass(int M, int k, int *h, int *fused) {
  int hh = M >> 1;
  for (int i = 0; i < k; ++i) hh >>= 2;
  *h = hh;
  *fused = hh >= 2;
}

// start of pass k's twiddles in the pass-major table, and the table's total length (entries)
BF_HD int bf_tw_offset(int M, int k) {
  int o = 0, hh = M >> 1;
  for (int i = 0; i < k; ++i) { o += hh >> 1; hh >>= 2; }
  return o;
}
BF_HD int bf_tw_total(int M) { return bf_tw_offset(M, bf_num_passes(M)); }

// ---------------------------------------------------------------------------------------------------------
// Register-blocked variant: S radix-2 stages per pass on 2^S elements held in registers (the same butterfly
// network, so the same bit-reversed spectrum order), with the array padded by one slot per 16 elements.
//
// Pass with leading half-size h: work item q in [0, M / 2^S) owns the elements i + r st, r < 2^S, st = h / 2^(S-1),
// i = (q / st) 2h + (q mod st).  Stage s (half-size h / 2^s, register distance d = 2^(S-1-s)) multiplies the lower
// output of the pair (r, r + d) by  W_{2h}^{j 2^s} * exp(-i pi (r mod d) / d):  a per-item factor obtained by
// repeated squaring of T[j] = W_{2h}^j (one table load per item) times a compile-time root of unity.
// Schedule (bf2_*): passes of 3 or 4 stages with st >= 16 down to half-size 16, then ONE pass of 4 stages on 16
// consecutive elements per item (st = 1).  With the padding a thread's 16 consecutive elements are 17 slots from its
// neighbour's, and every strided pass reads 8 consecutive slots per quarter warp: no bank conflicts in any pass.
// A second, coarser pad (one slot per 256 elements) makes the bit-reversed accesses of the whole-ring transforms
// conflict free as well: consecutive bins sit M/32 elements apart after bit reversal, a multiple of 256, which
// the 17/16 pad alone maps to one bank group (ring_pow2_kernel's fold and unfold ran 32-way conflicted).
template <bool PAD>
BF_HD int bf_pidx(int i) { return PAD ? i + (i >> 4) + (i >> 8) : i; }
BF_HD int bf_padded(int M) { return M + (M >> 4) + (M >> 8) + 1; }    // slots of a padded array of M elements

BF_HD double2 bf_root(int k, int d) {          // exp(-i pi k / d), d in {1, 2, 4, 8}, 0 <= k < d (constants after unrolling)
  const double c8[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                        0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
  const double s8[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                        1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  const int t = k * (8 / d);                   // exp(-i pi t / 8)
  double2 r; r.x = c8[t]; r.y = -s8[t];
  return r;
}

template <int S, bool PAD>
BF_HD void dif_itemS(double2 *x, int h, const double2 *T, int q) {
  constexpr int R = 1 << S;
  const int st = h >> (S - 1);
  const int j = q & (st - 1);
  const int i = (q - j) * R + j;
  double2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = x[bf_pidx<PAD>(i + r * st)];
  double2 w = T[j];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int d = R >> (s + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & d) continue;
      const int k = r & (d - 1);
      const double2 a = v[r], b = v[r + d];
      v[r] = bf_add(a, b);
      double2 t = bf_mul(bf_sub(a, b), w);
      if (k != 0) t = (2 * k == d) ? bf_mul_mi(t) : bf_mul(t, bf_root(k, d));
      v[r + d] = t;
    }
    w = bf_mul(w, w);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) x[bf_pidx<PAD>(i + r * st)] = v[r];
}

template <int S, bool PAD>
BF_HD void dit_itemS(double2 *x, int h, const double2 *T, int q) {
  constexpr int R = 1 << S;
  const int st = h >> (S - 1);
  const int j = q & (st - 1);
  const int i = (q - j) * R + j;
  double2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = x[bf_pidx<PAD>(i + r * st)];
  double2 wp[S];                                // T[j]^(2^s)
  wp[0] = T[j];
#pragma unroll
  for (int s = 1; s < S; ++s) wp[s] = bf_mul(wp[s - 1], wp[s - 1]);
#pragma unroll
  for (int s = S - 1; s >= 0; --s) {
    const int d = R >> (s + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & d) continue;
      const int k = r & (d - 1);
      double2 t = bf_mulc(v[r + d], wp[s]);
      if (k != 0) t = (2 * k == d) ? bf_mul_pi(t) : bf_mulc(t, bf_root(k, d));
      const double2 a = v[r];
      v[r] = bf_add(a, t);
      v[r + d] = bf_sub(a, t);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) x[bf_pidx<PAD>(i + r * st)] = v[r];
}

// The middle of the convolution for the register-blocked variant: the last DIF pass (half-sizes 8, 4, 2, 1), the
// multiplication by the bit-reversed filter spectrum and the first DIT pass all act on the 16 consecutive elements
// [16 q, 16 q + 16) of one work item, so they are done in registers in one go: one shared-memory round trip and no
// barrier instead of three passes.  No per-item twiddle (st = 1: T[j] = 1).
template <bool PAD>
BF_HD void conv_mid16(double2 *x, const double2 *v, int q) {
  constexpr int R = 16;
  const int i = q * R;
  double2 u[R];
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = x[bf_pidx<PAD>(i + r)];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int d = R >> (s + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & d) continue;
      const int k = r & (d - 1);
      const double2 a = u[r], b = u[r + d];
      u[r] = bf_add(a, b);
      double2 t = bf_sub(a, b);
      if (k != 0) t = (2 * k == d) ? bf_mul_mi(t) : bf_mul(t, bf_root(k, d));
      u[r + d] = t;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = bf_mul(u[r], v[i + r]);
#pragma unroll
  for (int s = 3; s >= 0; --s) {
    const int d = R >> (s + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & d) continue;
      const int k = r & (d - 1);
      double2 t = u[r + d];
      if (k != 0) t = (2 * k == d) ? bf_mul_pi(t) : bf_mulc(t, bf_root(k, d));
      const double2 a = u[r];
      u[r] = bf_add(a, t);
      u[r + d] = bf_sub(a, t);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) x[bf_pidx<PAD>(i + r)] = u[r];
}

// schedule of the register-blocked variant for M in {1024, 2048, 4096, 8192}: stage counts of the strided passes
// (leading half-size M/2 downwards, ending at half-size 16); the final pass has 4 stages from half-size 8.
BF_HD int bf2_num_strided(int M) { return M >= 8192 ? 3 : 2; }
BF_HD int bf2_stages(int M, int k) {             // stages of strided pass k
  if (M >= 8192) return 3;                       // 13 = 3 + 3 + 3 + 4
  if (M >= 4096) return 4;                       // 12 = 4 + 4 + 4
  if (M >= 2048) return k == 0 ? 4 : 3;          // 11 = 4 + 3 + 4
  return 3;                                      // 10 = 3 + 3 + 4
}
BF_HD int bf2_half(int M, int k) {               // leading half-size of strided pass k
  int h = M >> 1;
  for (int i = 0; i < k; ++i) h >>= bf2_stages(M, i);
  return h;
}
// twiddle table of the variant: pass k owns st_k = h_k / 2^(S_k - 1) entries T[j] = exp(-i pi j / h_k); the final
// pass has st = 1 and T = {1}
BF_HD int bf2_tw_offset(int M, int k) {
  int o = 0;
  for (int i = 0; i < k; ++i) o += bf2_half(M, i) >> (bf2_stages(M, i) - 1);
  return o;
}
BF_HD int bf2_tw_total(int M) { return bf2_tw_offset(M, bf2_num_strided(M)) + 1; }

BF_HD unsigned bf_bitrev(unsigned v, int bits) {
  unsigned r = 0;
  for (int b = 0; b < bits; ++b) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

}  // namespace cmdr
