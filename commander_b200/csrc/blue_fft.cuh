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

BF_HD unsigned bf_bitrev(unsigned v, int bits) {
  unsigned r = 0;
  for (int b = 0; b < bits; ++b) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

}  // namespace cmdr
