// bqa_f32x2.cuh -- packed fp32 pairs for sm_100a (FFMA2 / FMUL2 / FADD2: one issue slot, two fp32 operations).
//
// A pair is carried as a 64-bit value so that the register allocator keeps it in an aligned register pair (the float2
// intrinsics of sm_100_rt.h let the halves drift apart and pay two MOVs per operand).  ptxas folds a pair built from the
// same scalar twice into the broadcast operand form (`R7.F32`) and a pair negated on both halves into the operand
// negation of FFMA2, so `bcast(s)` and the `fnma2*` forms below cost no instruction (checked with cuobjdump -sass).
// Measured on B200 (scripts/ffma2_bench.cu): FFMA2 takes one issue slot and two FMA-pipe cycles, i.e. it halves the
// issue pressure of fp32 arithmetic at unchanged pipe throughput.
#pragma once
#include <cuda_runtime.h>

namespace bqa {
namespace x2 {

typedef unsigned long long p2;       // (lo, hi) fp32 pair

__device__ __forceinline__ p2 pk(float lo, float hi) {
  p2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ p2 pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ p2 bcast(float s) { return pk(s, s); }
__device__ __forceinline__ float2 unpk(p2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ p2 swap(p2 v) { const float2 r = unpk(v); return pk(r.y, r.x); }   // (hi, lo): an operand swizzle
__device__ __forceinline__ float hsum(p2 v) { const float2 r = unpk(v); return r.x + r.y; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) {       // a * b + c
  p2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ p2 fnma2(p2 a, p2 b, p2 c) {      // -a * b + c
  p2 d;
  asm("{ .reg .b64 t; .reg .f32 lo, hi; mov.b64 {lo, hi}, %1; neg.f32 lo, lo; neg.f32 hi, hi; mov.b64 t, {lo, hi};\n"
      "fma.rn.f32x2 %0, t, %2, %3; }"
      : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ p2 mul2(p2 a, p2 b) {
  p2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ p2 add2(p2 a, p2 b) {
  p2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// scalar-times-pair forms (the scalar is broadcast to both halves by the instruction)
__device__ __forceinline__ p2 fma2s(float s, p2 b, p2 c) { return fma2(bcast(s), b, c); }
__device__ __forceinline__ p2 fnma2s(float s, p2 b, p2 c) { return fnma2(bcast(s), b, c); }
__device__ __forceinline__ p2 mul2s(float s, p2 b) { return mul2(bcast(s), b); }

}  // namespace x2
}  // namespace bqa
