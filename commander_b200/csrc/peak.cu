// peak.cu -- FP64 FMA throughput probe.  MEASURED_PEAKS.json carries HBM and bf16 peaks only;
// the Legendre kernels are bound by the FP64 pipe, so bench.py measures that denominator live
// with this kernel (dependent DFMA chains, 8 independent chains per thread, no memory traffic).
#include <cstdio>

#include "kernels.h"

namespace cmdr {

// MODE 0: x = fma(u, w, x) with u warp-uniform (uniform register): two vector-register operands per
//         DFMA -- the pattern that reaches the pipe's issue rate of one warp-DFMA per 2 cycles per
//         scheduler (tools/ubench: 2.00 cycles).  This is the roofline denominator.
// MODE 1: x = fma(v, w, x), three distinct vector-register operands: 3.0 cycles per warp-DFMA on
//         B200 (the register file delivers one 64-bit operand per cycle per scheduler; operands
//         served by the reuse cache, uniform registers or constants are free).  Reported next to the
//         peak because ptxas-scheduled Legendre code sits between the two.
template <int MODE>
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters, double a, double b) {
  double x[8], w[8], v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; w[i] = 1e-9 * (threadIdx.x + i); v[i] = 1e-9 * (threadIdx.x + 2 * i + 1); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = MODE == 0 ? fma(a, w[q], x[q]) : fma(v[q], w[q], x[q]);
    }
  }
  double s = b;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += x[q];
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

// mode 0: DMMA (mma.sync m8n8k4 f64) only; mode 1: DMMA and DFMA interleaved 1:4 (same FMA count each);
// used once to decide whether a tensor-core Legendre variant could pay off.
__global__ void __launch_bounds__(256) dmma_probe_kernel(double *out, int iters, int mode, double a, double b) {
  double c0 = threadIdx.x * 1e-3, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
  double x0 = c0, x1 = c1, x2 = c2, x3 = c3, x4 = c4, x5 = c5, x6 = c6, x7 = c7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2), "+d"(c3) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c4), "+d"(c5) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c6), "+d"(c7) : "d"(a), "d"(b));
      if (mode == 1) {   // 4 DMMA = 1024 FMA per warp ; 32 warp-DFMA = 1024 FMA per warp
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
          x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
      }
    }
  }
  double s = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7 + x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}

}  // namespace cmdr

extern "C" double cmdr_sht_measure_dmma_tflops(int iters, int reps, int mode) {
  using namespace cmdr;
  int dev = 0, sms = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  CMDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = static_cast<double *>(scratch_get("probe", 64));
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t a, b;
  CMDR_CUDA_CHECK(cudaEventCreate(&a)); CMDR_CUDA_CHECK(cudaEventCreate(&b));
  dmma_probe_kernel<<<blocks, threads>>>(out, iters, mode, 0.999999, 1e-9);
  CMDR_CUDA_CHECK(cudaDeviceSynchronize());
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    CMDR_CUDA_CHECK(cudaEventRecord(a));
    dmma_probe_kernel<<<blocks, threads>>>(out, iters, mode, 0.999999, 1e-9);
    CMDR_CUDA_CHECK(cudaEventRecord(b));
    CMDR_CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CMDR_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    // per thread per iter: 8*4 DMMA * 256 FMA / 32 lanes = 256 FMA (+ 256 DFMA in mode 1)
    double fma_per_thread = 256.0 * iters * (mode == 1 ? 2.0 : 1.0);
    double tf = 2.0 * fma_per_thread * blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    count_launch();
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

template <int MODE>
static double measure_dfma(int iters, int reps) {
  using namespace cmdr;
  int dev = 0, sms = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  CMDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = static_cast<double *>(scratch_get("probe", 64));
  const int blocks = sms * 4, threads = 256;
  cudaEvent_t a, b;
  CMDR_CUDA_CHECK(cudaEventCreate(&a)); CMDR_CUDA_CHECK(cudaEventCreate(&b));
  dfma_probe_kernel<MODE><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  CMDR_CUDA_CHECK(cudaDeviceSynchronize());
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    CMDR_CUDA_CHECK(cudaEventRecord(a));
    dfma_probe_kernel<MODE><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    CMDR_CUDA_CHECK(cudaEventRecord(b));
    CMDR_CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    CMDR_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    count_launch();
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

// FP64 FMA peak (two vector-register operands per DFMA): the roofline denominator
extern "C" double cmdr_sht_measure_fp64_tflops(int iters, int reps) { return measure_dfma<0>(iters, reps); }
// the same with three distinct vector-register operands per DFMA (register-bandwidth bound)
extern "C" double cmdr_sht_measure_fp64_tflops_3op(int iters, int reps) { return measure_dfma<1>(iters, reps); }
