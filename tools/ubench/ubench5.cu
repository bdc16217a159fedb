// ubench5.cu -- the tensor-core question for the spin-2 analysis Legendre kernel, as a kernel-shaped microbenchmark
// (VERDICT r01 item 4b): per warp and per group of 8 l's,
//   variant DFMA : what anal2_kernel does per (8 l x 32 ring pairs): 32 recurrence DFMA + 64 accumulate DFMA per lane-set
//                  (12 DFMA per (l, ring pair): 4 recurrence + 8 accumulate) and the 5-stage shuffle butterfly;
//   variant DMMA : the same recurrence (one ring pair per lane), lambda values re-laid out through shared memory into
//                  mma.sync.m8n8k4.f64 A fragments (rows = 8 l, k = 4 ring pairs), 8 ring blocks x 2 matrices (P, M)
//                  = 16 DMMA per group; the B fragments (ring data) stay in registers; the K dimension replaces the
//                  butterfly.  Only 4 of the 8 N columns carry data: {zpN.re, zpN.im, zmS.re, zmS.im} for P and the
//                  mirrored set for M -- a single map offers no more right-hand sides that share one lambda matrix.
// Reports cycles per (l, ring pair) triple per SM sub-partition for both, at 2-4 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// DFMA variant: R = 4 ring pairs per lane (as the product kernel), 4 l per step, 16 partial sums reduced by a butterfly
__global__ void __launch_bounds__(32, 12) k_dfma(double *out, const double *in, int ngroups) {
  const int lane = threadIdx.x;
  double x[4], Pa[4], Pap[4], Pb[4], Pbp[4], w[4][8];
  for (int r = 0; r < 4; ++r) {
    x[r] = in[lane * 4 + r] * 1e-3; Pa[r] = in[r] + lane; Pb[r] = in[r + 4] - lane; Pap[r] = Pbp[r] = 0.5;
    for (int q = 0; q < 8; ++q) w[r][q] = in[8 + q] * (1 + r);
  }
  double tot = 0;
#pragma unroll 1
  for (int g = 0; g < ngroups; ++g) {
    double s[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0;
    const double A = in[64 + (g & 15)], C = in[80 + (g & 15)];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double va = Pa[r], vb = Pb[r];
        const double ua = fma(x[r], A, C), ub = fma(x[r], A, -C);
        const double na = fma(va, ua, -Pap[r]), nb = fma(vb, ub, -Pbp[r]);
        s[j][0] = fma(va, w[r][0], s[j][0]); s[j][1] = fma(va, w[r][1], s[j][1]);
        s[j][2] = fma(va, w[r][6], s[j][2]); s[j][3] = fma(va, w[r][7], s[j][3]);
        s[j][2] = fma(vb, w[r][2], s[j][2]); s[j][3] = fma(vb, w[r][3], s[j][3]);
        s[j][0] = fma(vb, w[r][4], s[j][0]); s[j][1] = fma(vb, w[r][5], s[j][1]);
        Pap[r] = va; Pa[r] = na * 1e-3; Pbp[r] = vb; Pb[r] = nb * 1e-3;
      }
    // 5-stage butterfly reduce-scatter of the 16 sums over the warp (16 shuffles + 16 adds)
    double v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[4 * j] = s[j][0]; v[4 * j + 1] = s[j][1]; v[4 * j + 2] = s[j][2]; v[4 * j + 3] = s[j][3]; }
#pragma unroll
    for (int st = 0, h = 8, bit = 16; st < 4; ++st, h >>= 1, bit >>= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < h) {
        const bool hi = lane & bit;
        const double send = hi ? v[i] : v[i + h], keep = hi ? v[i + h] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    tot += v[0];
  }
  out[blockIdx.x * 32 + lane] = tot + Pa[0] + Pb[1];
}

// DMMA variant: one ring pair per lane, 8 l per group
__global__ void __launch_bounds__(32, 12) k_dmma(double *out, const double *in, int ngroups) {
  __shared__ double lamP[8][33], lamM[8][33];      // [l][ring], padded
  const int lane = threadIdx.x;
  const double x = in[lane] * 1e-3;
  double Pa = in[0] + lane, Pb = in[4] - lane, Pap = 0.5, Pbp = 0.5;
  // B fragments: lane holds B[k = lane % 4][n = lane / 4] of every ring block (8 blocks x 2 matrices), kept in registers
  double bP[8], bM[8];
  for (int b = 0; b < 8; ++b) { bP[b] = (lane / 4 < 4) ? in[8 + b] * (1 + lane % 4) : 0.0; bM[b] = (lane / 4 < 4) ? in[16 + b] * (1 + lane % 4) : 0.0; }
  double tot = 0;
#pragma unroll 1
  for (int g = 0; g < ngroups; ++g) {
    const double A = in[64 + (g & 15)], C = in[80 + (g & 15)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {                   // recurrence: 4 DFMA per l
      lamP[j][lane] = Pa; lamM[j][lane] = Pb;
      const double ua = fma(x, A, C), ub = fma(x, A, -C);
      const double na = fma(Pa, ua, -Pap), nb = fma(Pb, ub, -Pbp);
      Pap = Pa; Pa = na * 1e-3; Pbp = Pb; Pb = nb * 1e-3;
    }
    __syncwarp();
    double d1a = 0, d1b = 0, d2a = 0, d2b = 0;      // D1 = P X_P, D2 = M X_M : rows l = lane / 4, cols 2 (lane % 4) + {0,1}
#pragma unroll
    for (int b = 0; b < 8; ++b) {                   // A fragment: A[row = lane / 4][k = lane % 4] = lambda[l = lane / 4][ring = 4 b + lane % 4]
      const double aP = lamP[lane >> 2][4 * b + (lane & 3)], aM = lamM[lane >> 2][4 * b + (lane & 3)];
      dmma(d1a, d1b, aP, bP[b]);
      dmma(d2a, d2b, aM, bM[b]);
    }
    __syncwarp();
    // S1 = D1[:, 0:2] +- D2[:, 2:4], S2 = D2[:, 0:2] +- D1[:, 2:4]: one exchange between column pairs
    const double o1 = __shfl_xor_sync(0xffffffffu, d2a, 2), o2 = __shfl_xor_sync(0xffffffffu, d1a, 2);
    tot += d1a + o1 + d2b + o2 + d1b + d2a;
  }
  out[blockIdx.x * 32 + lane] = tot + Pa + Pb;
}

int main() {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  double *out, *in; CK(cudaMalloc(&out, sizeof(double) * nsm * 64 * 32)); CK(cudaMalloc(&in, sizeof(double) * 256));
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 1.0 + 1e-3 * i;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int ng = 20000;
  for (int wps = 2; wps <= 4; ++wps) {             // warps per sub-partition (one warp per CTA, as the product kernels)
    const int nblk = nsm * 4 * wps;
    for (int variant = 0; variant < 2; ++variant) {
      float best = 1e30f;
      for (int t = 0; t < 4; ++t) {
        CK(cudaEventRecord(e0));
        if (variant == 0) k_dfma<<<nblk, 32>>>(out, in, ng); else k_dmma<<<nblk, 32>>>(out, in, ng);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (t && ms < best) best = ms;
      }
      const double cycles = best * 1e-3 * clk * 1e3;
      // triples per group per warp: DFMA variant 4 l x 128 ring pairs = 512, DMMA variant 8 l x 32 ring pairs = 256
      const double triples = (variant == 0 ? 512.0 : 256.0) * ng * wps;
      printf("%s warps/SMSP %d : %.3f cycles per (l, ring pair) triple per SMSP  (%.2f ms)\n", variant == 0 ? "DFMA+butterfly" : "DMMA m8n8k4   ", wps,
             cycles / triples, best);
    }
  }
  printf("reference: 12 DFMA per triple at 2 cycles per warp-DFMA = 0.75 cycles per triple per SMSP; the product's anal2_kernel runs at ~1.10\n");
  return 0;
}
