// ubench3.cu -- DFMA throughput vs operand reuse on sm_100a: does a DFMA with three distinct
// 64-bit register operands issue every 2 cycles per SMSP, or is it register-bandwidth limited?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// MODE 0: v[i] = fma(v[i], a, b)           (2 operands shared by all DFMAs)
// MODE 1: v[i] = fma(u[i], a, v[i])        (1 shared operand)
// MODE 2: v[i] = fma(u[i], w[i], v[i])     (no shared operand: 3 distinct register pairs each)
// MODE 3: v[i] = fma(u[i/4], w[i], v[i])   (first operand shared by 4 consecutive DFMAs, as in the Legendre accumulate)
// MODE 4: v[i] = fma(u[i], w[i/4], v[i])   (second operand shared by 4 consecutive)
template <int MODE, int N>
__global__ void k(double *out, const double *in, int iters) {
  double v[N], u[N], w[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { v[i] = in[i] + threadIdx.x; u[i] = in[N + i] * 1e-3; w[i] = in[2 * N + i] * 1e-3; }
  const double a = in[100], b = in[101];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (MODE == 0) v[i] = fma(v[i], a, b);
        if (MODE == 1) v[i] = fma(u[i], a, v[i]);
        if (MODE == 2) v[i] = fma(u[i], w[i], v[i]);
        if (MODE == 3) v[i] = fma(u[i / 4], w[i], v[i]);
        if (MODE == 4) v[i] = fma(u[i], w[i / 4], v[i]);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int N>
int run(int warps_per_smsp, double *out, double *in) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  const int iters = 20000;
  dim3 grid(nsm), block(128 * warps_per_smsp);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE, N><<<grid, block>>>(out, in, 100);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int t = 0; t < 3; ++t) {
    CK(cudaEventRecord(e0));
    k<MODE, N><<<grid, block>>>(out, in, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double cycles = best * 1e-3 * clk * 1e3;
  printf("MODE %d N %2d warps/SMSP %d : %.3f cycles per DFMA per SMSP\n", MODE, N, warps_per_smsp,
         cycles / ((double)iters * 4 * N * warps_per_smsp));
  return 0;
}

int main() {
  double *out, *in; CK(cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4)); CK(cudaMalloc(&in, sizeof(double) * 256));
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w = 1; w <= 4; ++w) {
    if (w == 3) continue;
    run<0, 16>(w, out, in); run<1, 16>(w, out, in); run<2, 16>(w, out, in); run<3, 16>(w, out, in); run<4, 16>(w, out, in);
    run<2, 8>(w, out, in); run<2, 24>(w, out, in);
  }
  return 0;
}
