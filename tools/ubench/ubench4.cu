// ubench4.cu -- calibration of the operand-reuse model: DFMA streams in which G consecutive instructions
// share one vector-register operand (the pattern of the Legendre accumulate FMAs).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// 16 accumulators v[i]; operand u[i / G] shared by G consecutive FMAs; w[i] distinct.
// SLOT 0: fma(u, w, v)  SLOT 1: fma(w, u, v) (the compiler may commute; check the SASS)
// ALU > 0: one independent integer op after every FMA (does an intervening instruction break reuse?)
template <int G, int ALU>
__global__ void k(double *out, const double *in, int iters, int c) {
  double v[16], u[16], w[16];
  int q[4] = {(int)threadIdx.x, c, c + 1, c + 2};
#pragma unroll
  for (int i = 0; i < 16; ++i) { v[i] = in[i] + threadIdx.x; u[i] = in[16 + i] * 0.999 - threadIdx.x * 1e-9; w[i] = in[32 + i] * 1e-3 + threadIdx.x * 1e-9; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = fma(u[i / G], v[i], w[i]);
        if (ALU) q[i & 3] = (q[i & 3] & c) ^ it;
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + q[0] + q[1] + q[2] + q[3];
}

template <int G, int ALU>
int run(int warps_per_smsp, double *out, double *in) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  const int iters = 20000;
  dim3 grid(nsm), block(128 * warps_per_smsp);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<G, ALU><<<grid, block>>>(out, in, 100, 0x7fffffff);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int t = 0; t < 3; ++t) {
    CK(cudaEventRecord(e0));
    k<G, ALU><<<grid, block>>>(out, in, iters, 0x7fffffff);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double cycles = best * 1e-3 * clk * 1e3;
  printf("share-group %2d  alu %d  warps/SMSP %d : %.3f cycles per DFMA per SMSP   (model: %.3f)\n", G, ALU, warps_per_smsp,
         cycles / ((double)iters * 64 * warps_per_smsp), (3.0 + 2.0 * (G - 1)) / G);
  return 0;
}

int main() {
  double *out, *in; CK(cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4)); CK(cudaMalloc(&in, sizeof(double) * 256));
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w = 1; w <= 4; w *= 2) {
    run<1, 0>(w, out, in); run<2, 0>(w, out, in); run<4, 0>(w, out, in); run<8, 0>(w, out, in); run<16, 0>(w, out, in);
    run<4, 1>(w, out, in); run<8, 1>(w, out, in);
  }
  return 0;
}
