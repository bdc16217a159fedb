// ubench2.cu -- does non-FP64 work issue in the shadow of DFMA on sm_100a?  Independent ops only.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// per loop body: 8 independent DFMA, NALU independent LOP3, NSHFL independent SHFL.32, NLDS independent LDS.32(+LOP3)
template <int NALU, int NSHFL, int NLDS, int NDFMA>
__global__ void k(double *out, int iters, double a, double b, int c) {
  __shared__ int sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * c;
  __syncthreads();
  double v[8];
  int w[16], s[16], q[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = a + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) { w[i] = threadIdx.x * (i + 1); s[i] = threadIdx.x + i; q[i] = i; }
  const int *smp = sm + (threadIdx.x & 31);
#pragma unroll 1
  for (int it0 = 0; it0 < iters; it0 += 4) {
#pragma unroll
   for (int u = 0; u < 4; ++u) {
    const int it = it0 + u;
#pragma unroll
    for (int i = 0; i < NDFMA; ++i) v[i] = fma(v[i], a, b);
#pragma unroll
    for (int i = 0; i < NALU; ++i) w[i] = (w[i] & c) ^ it;           // 1 LOP3
#pragma unroll
    for (int i = 0; i < NSHFL; ++i) s[i] = __shfl_xor_sync(0xffffffffu, s[i], 1 + i);
#pragma unroll
    for (int i = 0; i < NLDS; ++i) q[i] ^= smp[i * 32 + (it & 31) * 32];
   }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += v[i];
  int t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += w[i] + s[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + t;
}

template <int NALU, int NSHFL, int NLDS, int NDFMA = 8>
int run(int warps_per_smsp, double *out) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  const int iters = 40000;
  dim3 grid(nsm), block(128 * warps_per_smsp);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<NALU, NSHFL, NLDS, NDFMA><<<grid, block>>>(out, 200, 1.0000001, 1e-9, 0x7fffffff);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int t = 0; t < 3; ++t) {
    CK(cudaEventRecord(e0));
    k<NALU, NSHFL, NLDS, NDFMA><<<grid, block>>>(out, iters, 1.0000001, 1e-9, 0x7fffffff);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double cycles = best * 1e-3 * clk * 1e3;
  printf("DFMA %d ALU %2d SHFL %2d LDS %2d | warps/SMSP %d : %.2f cycles per loop body per SMSP (per warp-body: %.2f)\n", NDFMA, NALU, NSHFL, NLDS,
         warps_per_smsp, cycles / iters / warps_per_smsp, cycles / iters);
  return 0;
}

int main() {
  double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4));
  for (int w = 1; w <= 4; w *= 2) {
    if (w == 1) { run<0, 0, 0>(1, out); run<4, 0, 0>(1, out); run<8, 0, 0>(1, out); run<16, 0, 0>(1, out);
      run<0, 4, 0>(1, out); run<0, 8, 0>(1, out); run<0, 16, 0>(1, out); run<0, 0, 4>(1, out); run<0, 0, 8>(1, out); run<0, 0, 16>(1, out); run<8, 8, 0>(1, out); }
    if (w == 2) { run<0, 0, 0>(2, out); run<4, 0, 0>(2, out); run<8, 0, 0>(2, out); run<16, 0, 0>(2, out);
      run<0, 4, 0>(2, out); run<0, 8, 0>(2, out); run<0, 16, 0>(2, out); run<0, 0, 4>(2, out); run<0, 0, 8>(2, out); run<0, 0, 16>(2, out); run<8, 8, 0>(2, out); }
    if (w == 4) { run<0, 0, 0>(4, out); run<4, 0, 0>(4, out); run<8, 0, 0>(4, out); run<16, 0, 0>(4, out);
      run<0, 4, 0>(4, out); run<0, 8, 0>(4, out); run<0, 16, 0>(4, out); run<0, 0, 4>(4, out); run<0, 0, 8>(4, out); run<0, 0, 16>(4, out); run<8, 8, 0>(4, out); }
  }
  // no DFMA at all: raw cost of the other ops
  run<16, 0, 0, 0>(4, out); run<0, 16, 0, 0>(4, out); run<0, 0, 16, 0>(4, out);
  return 0;
}
