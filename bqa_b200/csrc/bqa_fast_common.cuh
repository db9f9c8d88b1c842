// bqa_fast_common.cuh -- device helpers shared by the specialised complex64 kernels.
#pragma once
#include <cuda_runtime.h>

#include "bqa_f32x2.cuh"

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

// ---- packed (FFMA2) complex arithmetic: a complex number is one fp32 pair (re, im) -----------------------------
using x2::p2;

// 16 complex (128 bytes) shared <-> register pairs
__device__ __forceinline__ void lds_tile(p2 (&r)[16], const unsigned char* p) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + 16 * i);
    r[2 * i] = v.x;
    r[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ void sts_tile(unsigned char* p, const p2 (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<ulonglong2*>(p + 16 * i) = make_ulonglong2(r[2 * i], r[2 * i + 1]);
}

// Two-accumulator complex multiply-add: with s = (sx, sy) used as broadcast scalars and v as a pair,
//   A += sx * v,  B += sy * v     (2 FFMA2 per complex multiply-add, no operand preparation)
// and at the end  sum s v = (A.x - B.y, A.y + B.x),  sum conj(s) v = (A.x + B.y, A.y - B.x).
struct CAcc {
  p2 A, B;
};
template <bool FIRST>
__device__ __forceinline__ void cmac(CAcc& acc, p2 s, p2 v) {
  const float2 f = x2::unpk(s);
  if (FIRST) {
    acc.A = x2::mul2s(f.x, v);
    acc.B = x2::mul2s(f.y, v);
  } else {
    acc.A = x2::fma2s(f.x, v, acc.A);
    acc.B = x2::fma2s(f.y, v, acc.B);
  }
}
// the same with a REAL scalar s (the diagonal of a Hermitian message): only the first accumulator moves.  B_FIRST marks
// the first term that touches B when the chain started with a real scalar.
template <bool FIRST>
__device__ __forceinline__ void cmac_real(CAcc& acc, p2 s, p2 v) {
  const float2 f = x2::unpk(s);
  acc.A = FIRST ? x2::mul2s(f.x, v) : x2::fma2s(f.x, v, acc.A);
}
template <bool B_FIRST>
__device__ __forceinline__ void cmac_bfirst(CAcc& acc, p2 s, p2 v) {
  const float2 f = x2::unpk(s);
  acc.A = x2::fma2s(f.x, v, acc.A);
  acc.B = B_FIRST ? x2::mul2s(f.y, v) : x2::fma2s(f.y, v, acc.B);
}
// (one FFMA2 each: the second accumulator enters half-swapped and multiplied by (-1, 1) or (1, -1); x * (+-1) is exact,
// so the results equal the separately rounded additions)
__device__ __forceinline__ p2 cfinish(const CAcc& acc) {          // sum s v = (A.x - B.y, A.y + B.x)
  return x2::fma2(x2::swap(acc.B), x2::pk(-1.f, 1.f), acc.A);
}
__device__ __forceinline__ float2 cfinish_conj(const CAcc& acc) {  // sum conj(s) v = (A.x + B.y, A.y - B.x)
  return x2::unpk(x2::fma2(x2::swap(acc.B), x2::pk(1.f, -1.f), acc.A));
}

}  // namespace fast
}  // namespace bqa
