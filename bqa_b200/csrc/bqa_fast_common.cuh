// bqa_fast_common.cuh -- device helpers shared by the specialised complex64 kernels.
#pragma once
#include <cuda_runtime.h>

namespace bqa {
namespace fast {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b
__device__ __forceinline__ void fma_c(float2& acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void fma_cc(float2& acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
}

// 16 complex (128 bytes) shared -> registers
__device__ __forceinline__ void lds_tile(float2 (&r)[16], const unsigned char* p) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 16 * i);
    r[2 * i] = make_float2(v.x, v.y);
    r[2 * i + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void sts_tile(unsigned char* p, const float2 (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(p + 16 * i) = make_float4(r[2 * i].x, r[2 * i].y, r[2 * i + 1].x, r[2 * i + 1].y);
}

}  // namespace fast
}  // namespace bqa
