// peak.cu -- FP64 FMA throughput probe.  MEASURED_PEAKS.json carries HBM and bf16 peaks only;
// the Legendre kernels are bound by the FP64 pipe, so bench.py measures that denominator live
// with this kernel (dependent DFMA chains, 8 independent chains per thread, no memory traffic).
#include <cstdio>

#include "kernels.h"

namespace cmdr {

__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace cmdr

extern "C" double cmdr_sht_measure_fp64_tflops(int iters, int reps) {
  using namespace cmdr;
  int dev = 0, sms = 0;
  CMDR_CUDA_CHECK(cudaGetDevice(&dev));
  CMDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = static_cast<double *>(scratch_get("probe", 64));
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t a, b;
  CMDR_CUDA_CHECK(cudaEventCreate(&a)); CMDR_CUDA_CHECK(cudaEventCreate(&b));
  dfma_probe_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  CMDR_CUDA_CHECK(cudaDeviceSynchronize());
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    CMDR_CUDA_CHECK(cudaEventRecord(a));
    dfma_probe_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
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
