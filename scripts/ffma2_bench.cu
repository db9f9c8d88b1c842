// microbenchmark: FFMA vs FFMA2 issue throughput on sm_100a, alone and mixed with ALU work
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float s, int iters, unsigned salt) {
  float2 a[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i); u[i] = threadIdx.x * 77u + i; }
  float2 m = make_float2(s, s * 0.5f), c = make_float2(s * 0.25f, s * 0.125f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((MODE & 1) == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        else a[i] = __ffma2_rn(a[i], m, c);
        if (MODE & 2) { u[i] = (u[i] ^ salt) + (u[(i + 1) & 7] & 0x55555555u); }   // ALU-pipe work (LOP3 + IADD3)
      }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(float* out, int wpb, int iters) {
  k<MODE><<<148, wpb * 32>>>(out, 1.0001f, iters, 0x9e3779b9u);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 32 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 10000;
  const char* names[4] = {"FFMA       ", "FFMA2      ", "FFMA +2ALU ", "FFMA2+2ALU "};
  for (int mode = 0; mode < 4; ++mode)
    for (int wpb = 4; wpb <= 16; wpb *= 2) {   // warps per SM
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) run<0>(out, wpb, iters); else if (mode == 1) run<1>(out, wpb, iters);
        else if (mode == 2) run<2>(out, wpb, iters); else run<3>(out, wpb, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 2.0 * 64 * iters * 148.0 * wpb * 32;   // scalar FMAs
      printf("%s warps/SM %2d: %.3f ms, scalar FMA per SM per clk @1.965GHz: %.1f  (cycles per float2-update per SMSP-warp: %.2f)\n", names[mode], wpb, ms,
             fma / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (64.0 * iters * wpb / 4));
    }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
