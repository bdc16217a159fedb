// ubench.cu -- FP64 issue-model microbenchmarks for sm_100a (tuning aid, not product code).
// Answers: DFMA latency / throughput per SMSP, and whether ALU / SHFL / LDS instructions issue
// "in the shadow" of DFMA or take issue slots away from it.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP, int NALU, int NSHFL, int NLDS>
__global__ void k(double *out, int iters, double a, double b, int sel) {
  __shared__ double sm[256];
  sm[threadIdx.x & 255] = threadIdx.x;
  __syncthreads();
  double v[ILP];
  int w[4] = {(int)threadIdx.x, sel, sel + 1, sel + 2};
  double s = threadIdx.x, acc = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        v[i] = fma(v[i], a, b);
        if (i < NALU) w[i & 3] = (w[i & 3] ^ (w[(i + 1) & 3] + 0x9e3779b9)) + (w[(i + 2) & 3] >> 3);   // ~3 ALU ops
        if (i < NSHFL) s = __shfl_xor_sync(0xffffffffu, s, 1 + (i & 15));                          // 2 SHFL
        if (i < NLDS) acc += sm[(w[0] + i * 32 + rep) & 255];                                       // LDS + DADD
      }
    }
  }
  double r = s + acc;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + w[0] + w[1] + w[2] + w[3];
}

template <int ILP, int NALU, int NSHFL, int NLDS>
int run(const char *name, int warps_per_smsp, double *out) {
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  const int iters = 4000;
  dim3 grid(nsm), block(128 * warps_per_smsp);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<ILP, NALU, NSHFL, NLDS><<<grid, block>>>(out, 200, 1.0000001, 1e-9, 3);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int t = 0; t < 3; ++t) {
    CK(cudaEventRecord(e0));
    k<ILP, NALU, NSHFL, NLDS><<<grid, block>>>(out, iters, 1.0000001, 1e-9, 3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double dfma_per_warp = (double)iters * 8 * ILP;
  double cycles = best * 1e-3 * clk * 1e3;
  double per_smsp = dfma_per_warp * warps_per_smsp;
  printf("%-34s ILP %2d warps/SMSP %d : %.3f cycles per DFMA per SMSP (%.2f TFLOP/s)  [clk %d kHz]\n", name, ILP, warps_per_smsp,
         cycles / per_smsp, per_smsp * 4 * nsm * 64.0 / (best * 1e-3) / 1e12, clk);
  return 0;
}

int main() {
  double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4));
  // latency: 1 warp, ILP 1
  run<1, 0, 0, 0>("dfma chain (latency)", 1, out);
  run<2, 0, 0, 0>("dfma", 1, out);
  run<4, 0, 0, 0>("dfma", 1, out);
  run<8, 0, 0, 0>("dfma", 1, out);
  run<16, 0, 0, 0>("dfma", 1, out);
  run<8, 0, 0, 0>("dfma", 2, out);
  run<8, 0, 0, 0>("dfma", 4, out);
  run<4, 0, 0, 0>("dfma", 2, out);
  run<4, 0, 0, 0>("dfma", 3, out);
  run<2, 0, 0, 0>("dfma", 3, out);
  run<1, 0, 0, 0>("dfma", 3, out);
  run<1, 0, 0, 0>("dfma", 4, out);
  run<1, 0, 0, 0>("dfma", 8, out);
  // ALU interleave: per 8 DFMA, NALU x ~3 ALU ops
  run<8, 2, 0, 0>("dfma + 2x3 ALU per 8", 2, out);
  run<8, 4, 0, 0>("dfma + 4x3 ALU per 8", 2, out);
  run<8, 8, 0, 0>("dfma + 8x3 ALU per 8", 2, out);
  run<8, 8, 0, 0>("dfma + 8x3 ALU per 8", 4, out);
  // SHFL interleave
  run<8, 0, 1, 0>("dfma + 1 shfl per 8", 2, out);
  run<8, 0, 2, 0>("dfma + 2 shfl per 8", 2, out);
  run<8, 0, 4, 0>("dfma + 4 shfl per 8", 2, out);
  run<8, 0, 4, 0>("dfma + 4 shfl per 8", 4, out);
  // LDS interleave
  run<8, 0, 0, 2>("dfma + 2 (lds+dadd) per 8", 2, out);
  run<8, 0, 0, 4>("dfma + 4 (lds+dadd) per 8", 2, out);
  run<8, 4, 2, 0>("dfma + 4x3 ALU + 2 shfl per 8", 3, out);
  return 0;
}
