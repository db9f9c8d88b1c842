// bqa_fast_gram_d3D8.cuh -- the per-node contraction of the BP update (node_gram, bqa_core.cuh; reference
// Tensor._apply_msgs_but_one / _compute_msg, src/bqa/backends.py:381-408) for degree 3, D = 8, complex64, one warp per node.
//
// The generic routine works out of a per-warp global-memory workspace with runtime strides: at D = 8 it is bound by load
// instructions and their latency (profiles/r2_generic_D8_ncu_full.md: FMA pipe 16 %, 19 % of the instructions global
// loads, DRAM traffic 7.5 x the algorithmic bytes).  Here the node tensor and the two partial products live in shared
// memory, one physical half at a time (the Gram parts of p = 0 and p = 1 are independent):
//
//   T[p], P, E : 8 x 8 x 8 complex64 each, element (a, b, c) at a * 78 + b * 10 + c  (4.9 KB per array)
//
// The padding keeps the 16-byte accesses of the routine conflict free (a step of 78 or 10 elements moves by an odd number
// of 16-byte bank groups: eight such steps hit the eight groups once each) and keeps pairs along c 16-byte aligned.  Same sequence of
// contractions as node_gram (2 + 3 mode products, 3 closing contractions):
//   k = 0: E = m1 . P (leg 1), E = m2 . E (leg 2), gram_0 = T^H E, P = m0 . P (leg 0)
//   k = 1: E = m2 . P (leg 2),                     gram_1 = T^H E, P = m1 . P (leg 1)
//   k = 2:                                         gram_2 = T^H P
// A lane owns two whole fibres of a mode product (in-place safe) and two adjacent entries (x, 2 yq), (x, 2 yq + 1) of a
// Gram matrix; arithmetic is packed (FFMA2, bqa_f32x2.cuh): a complex multiply-add is two FFMA2.
// Output: gram[k][p][x][y] like node_gram, so the epilogues (emit_bp_msg / emit_ext_msg) are shared.
#pragma once
#include <cuda_runtime.h>

#include "bqa_f32x2.cuh"

namespace bqa {
namespace gram8 {

constexpr int kSA = 78, kSB = 10;                     // element strides of legs 0 and 1 (leg 2: 1); 7 * kSB + 7 < kSA
constexpr int kArr = 624;                             // elements per array (last index 7 * 78 + 7 * 10 + 7 = 623)
constexpr int kWarpElems = 3 * kArr + 3 * 64;         // T, P, E, the three messages
constexpr int kWarpBytes = kWarpElems * 8;            // 16 512 B

using x2::p2;

// out[x] = sum_b m[x][b] v[b] for the two fibres of this lane; fibre element b sits at base + b * S in src, result x at
// base + x * S in dst (dst == src allowed: all loads precede the stores, and a lane owns its fibres)
template <int S>
__device__ __forceinline__ void mode_product(const float2* src, float2* dst, int base0, int base1, const float2* m) {
  p2 v0[8], r0[8], v1[8], r1[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const float2 a = src[base0 + b * S], c = src[base1 + b * S];
    v0[b] = x2::pk(a.x, a.y); r0[b] = x2::pk(-a.y, a.x);        // i * v
    v1[b] = x2::pk(c.x, c.y); r1[b] = x2::pk(-c.y, c.x);
  }
#pragma unroll 1                     // rolled: fully unrolled, the routine's five products and three contractions came to about
  for (int x = 0; x < 8; ++x) {      // 75 KB of straight-line code and "no instruction" was the largest stall (r2 capture)
    p2 a0 = x2::pk(0.f, 0.f), a1 = a0;
#pragma unroll
    for (int b2 = 0; b2 < 4; ++b2) {
      const float4 mm = *reinterpret_cast<const float4*>(m + x * 8 + 2 * b2);     // m[x][2 b2], m[x][2 b2 + 1]
      a0 = x2::fma2s(mm.x, v0[2 * b2], a0); a0 = x2::fma2s(mm.y, r0[2 * b2], a0);
      a1 = x2::fma2s(mm.x, v1[2 * b2], a1); a1 = x2::fma2s(mm.y, r1[2 * b2], a1);
      a0 = x2::fma2s(mm.z, v0[2 * b2 + 1], a0); a0 = x2::fma2s(mm.w, r0[2 * b2 + 1], a0);
      a1 = x2::fma2s(mm.z, v1[2 * b2 + 1], a1); a1 = x2::fma2s(mm.w, r1[2 * b2 + 1], a1);
    }
    dst[base0 + x * S] = x2::unpk(a0);
    dst[base1 + x * S] = x2::unpk(a1);
  }
}

// acc0 += conj(t) e0, acc1 += conj(t) e1:  conj(t) e = e.x (t.x, -t.y) + e.y (t.y, t.x)
__device__ __forceinline__ void cmacc2(p2& acc0, p2& acc1, float2 t, float2 e0, float2 e1) {
  const p2 tc = x2::pk(t.x, -t.y), ts = x2::pk(t.y, t.x);
  acc0 = x2::fma2s(e0.x, tc, acc0); acc0 = x2::fma2s(e0.y, ts, acc0);
  acc1 = x2::fma2s(e1.x, tc, acc1); acc1 = x2::fma2s(e1.y, ts, acc1);
}

// gram[x][y] = sum over the two other legs of conj(T[.. x ..]) E[.. y ..] for y = 2 yq, 2 yq + 1; XS = stride of the open
// leg (0: kSA, 1: kSB), OS = stride of the other outer leg; the innermost leg c runs contiguously
template <int XS, int OS>
__device__ __forceinline__ void gram_outer(const float2* T, const float2* E, int x, int yq, float2* out) {
  p2 acc0 = x2::pk(0.f, 0.f), acc1 = acc0;
  const float2* tp = T + x * XS;
  const float2* e0p = E + (2 * yq) * XS;
  const float2* e1p = e0p + XS;
#pragma unroll 1
  for (int o = 0; o < 8; ++o) {
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      const float4 t = *reinterpret_cast<const float4*>(tp + o * OS + 2 * c2);
      const float4 a = *reinterpret_cast<const float4*>(e0p + o * OS + 2 * c2);
      const float4 b = *reinterpret_cast<const float4*>(e1p + o * OS + 2 * c2);
      cmacc2(acc0, acc1, make_float2(t.x, t.y), make_float2(a.x, a.y), make_float2(b.x, b.y));
      cmacc2(acc0, acc1, make_float2(t.z, t.w), make_float2(a.z, a.w), make_float2(b.z, b.w));
    }
  }
  out[x * 8 + 2 * yq] = x2::unpk(acc0);
  out[x * 8 + 2 * yq + 1] = x2::unpk(acc1);
}

// open leg = 2 (innermost): gram[x][y] = sum_{a, b} conj(T[a, b, x]) E[a, b, y]
__device__ __forceinline__ void gram_inner(const float2* T, const float2* E, int x, int yq, float2* out) {
  p2 acc0 = x2::pk(0.f, 0.f), acc1 = acc0;
#pragma unroll 1
  for (int a = 0; a < 8; ++a) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const float2 t = T[a * kSA + b * kSB + x];
      const float4 e = *reinterpret_cast<const float4*>(E + a * kSA + b * kSB + 2 * yq);
      cmacc2(acc0, acc1, t, make_float2(e.x, e.y), make_float2(e.z, e.w));
    }
  }
  out[x * 8 + 2 * yq] = x2::unpk(acc0);
  out[x * 8 + 2 * yq + 1] = x2::unpk(acc1);
}

// T: the node tensor (2 x 8 x 8 x 8, global), m0..m2: its incoming messages (8 x 8, global), gram: 3 x 2 x 8 x 8 output,
// sm: this warp's kWarpElems elements of shared memory (16-byte aligned)
__device__ __forceinline__ void node_gram_d3D8(const float2* __restrict__ T, const float2* __restrict__ m0,
                                               const float2* __restrict__ m1, const float2* __restrict__ m2,
                                               float2* gram, float2* sm) {
  const int lane = threadIdx.x & 31;
  float2* Ts = sm;
  float2* Ps = sm + kArr;
  float2* Es = sm + 2 * kArr;
  float2* Ms = sm + 3 * kArr;                          // m0 | m1 | m2
  __syncwarp();                                        // the previous node's readers of this scratch are done
  for (int i = lane; i < 64; i += 32) {
    Ms[i] = m0[i];
    Ms[64 + i] = m1[i];
    Ms[128 + i] = m2[i];
  }
  const int x = lane >> 2, yq = lane & 3;
  // fibres of this lane in the three mode products (f = lane, lane + 32 -> two (outer, inner) pairs)
  const int f0 = lane, f1 = lane + 32;
  const int l0a = (f0 >> 3) * kSB + (f0 & 7), l0b = (f1 >> 3) * kSB + (f1 & 7);     // leg 0: fixed (b, c)
  const int l1a = (f0 >> 3) * kSA + (f0 & 7), l1b = (f1 >> 3) * kSA + (f1 & 7);     // leg 1: fixed (a, c)
  const int l2a = (f0 >> 3) * kSA + (f0 & 7) * kSB, l2b = (f1 >> 3) * kSA + (f1 & 7) * kSB;   // leg 2: fixed (a, b)
#pragma unroll 1
  for (int p = 0; p < 2; ++p) {
    __syncwarp();
    const float4* src = reinterpret_cast<const float4*>(T + p * 512);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i2 = lane + 32 * j;                    // elements 2 i2, 2 i2 + 1: same (a, b), c even
      const float4 v = src[i2];
      const int at = (i2 >> 5) * kSA + ((i2 >> 2) & 7) * kSB + (i2 & 3) * 2;
      *reinterpret_cast<float4*>(Ts + at) = v;
      *reinterpret_cast<float4*>(Ps + at) = v;
    }
    __syncwarp();
    float2* g0 = gram + p * 64;
    // k = 0
    mode_product<kSB>(Ps, Es, l1a, l1b, Ms + 64);
    __syncwarp();
    mode_product<1>(Es, Es, l2a, l2b, Ms + 128);
    __syncwarp();
    gram_outer<kSA, kSB>(Ts, Es, x, yq, g0);
    __syncwarp();
    mode_product<kSA>(Ps, Ps, l0a, l0b, Ms);
    __syncwarp();
    // k = 1
    mode_product<1>(Ps, Es, l2a, l2b, Ms + 128);
    __syncwarp();
    gram_outer<kSB, kSA>(Ts, Es, x, yq, g0 + 128);
    __syncwarp();
    mode_product<kSB>(Ps, Ps, l1a, l1b, Ms + 64);
    __syncwarp();
    // k = 2
    gram_inner(Ts, Ps, x, yq, g0 + 256);
  }
  __syncwarp();
}

}  // namespace gram8
}  // namespace bqa
