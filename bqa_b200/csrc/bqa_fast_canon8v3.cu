// bqa_fast_canon8v3.cu -- canonicalizers of bond dimension 4 (extended dimension n = 8) in complex64, third layout.
//
// Same algorithm as bqa_fast_canon8v2.cu (replaces _get_canonicalizers, src/bqa/state.py:171-200: Cholesky factor +
// column Jacobi for the two eigenproblems of an edge, backends.py:483-490; one-sided Jacobi on the stacked matrix
// [ker ; conj(ul_b)] for the SVD, state.py:186-200) -- what changes is who holds what:
//
//   v2: ONE lane per 8 x 8 matrix (128 registers of matrix, 255 in all, 8 warps per SM).  ncu: FMA pipe 57 % busy,
//       "wait" (fixed-latency dependencies) the largest stall -- at 255 registers the compiler cannot keep independent
//       chains apart, and two warps per sub-partition do not cover it.
//   v3: TWO lanes per matrix, rows 0-3 and 4-7 (64 registers of matrix, <= 128 in all, 16 warps per SM).  A column
//       inner product is two partial sums and one xor shuffle; the four rotations of a round are derived two per lane
//       and fetched by shuffle, so no lane repeats another lane's parameter arithmetic.  Same FMA work per matrix,
//       twice the warps to hide its latencies.
//
// A quad of lanes owns an edge (8 edges per warp iteration): lane q = 2 m + h holds row half h of matrix m; eigen phase:
// m = 0 is m_f, m = 1 is m_b; SVD phase: m = 0 is ker (leaders: derive the rotations), m = 1 is the stacked block
// conj(ul_b) (followers: take the leaders' rotations).  One copy of the sweep code serves both phases.
#include <cuda_runtime.h>

#include <cstdlib>

#include "bqa_core.cuh"
#include "bqa_f32x2.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace canon8v3 {

using x2::p2;

#ifndef BQA_CANON3_WARPS
#define BQA_CANON3_WARPS 16
#endif
constexpr int kWarps = BQA_CANON3_WARPS;
constexpr int kEdges = 8;                  // edges per warp iteration (4 lanes each)
constexpr int kMat = 528;                  // 8 rows x 64 bytes + 16 (an odd number of 16-byte units: bank spreading)
constexpr int kPad = 512;                  // the 16 spare bytes of a matrix slot: row permutation of its owner
constexpr int kPair = 3 * kMat;            // per edge: F (A_f published) | B (input m_f, then A_b) | Q (input m_b)
constexpr int kWarpBytes = kEdges * kPair;
constexpr int kBars = kWarps * kWarpBytes;
constexpr int kSmem = kBars + kWarps * 8;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

__device__ unsigned long long g_stats[7];  // as in bqa_fast_canon8v2.cu

struct Half {                              // column j, k = 0, 1: rows (4 h + 2 k, 4 h + 2 k + 1); X = real, Y = imaginary parts
  p2 X[8][2], Y[8][2];
};

__device__ __forceinline__ float rsqrt_nr(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (1.5f - 0.5f * x * y * y);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Rot {
  float c, s, phx, phy, dw;
};
__device__ __forceinline__ Rot rot_params(float al, float be, float gr, float gi, float nul, float tol2, bool frozen,
                                          float& mxg2, float& mxs2) {
  const float g2 = gr * gr + gi * gi;
  const float ab = al * be;
  const bool act = !frozen && !(al <= nul || be <= nul || g2 <= tol2 * ab);
  const float ig = rsqrt_nr(g2 * 0x1p60f) * 0x1p30f;
  const float ag = g2 * ig;
  const float zeta = 0.5f * (be - al) * ig;
  const float z2 = 1.f + zeta * zeta;
  const float t0 = __fdividef(1.f, fabsf(zeta) + z2 * rsqrt_nr(z2));
  const float t = act ? copysignf(t0, zeta) : 0.f;
  Rot r;
  r.c = rsqrt_nr(1.f + t * t);
  r.s = r.c * t;
  r.phx = act ? gr * ig : 1.f;
  r.phy = act ? -gi * ig : 0.f;
  r.dw = act ? t * ag : 0.f;
  mxg2 = fmaxf(mxg2, act ? __fdividef(g2, ab) : 0.f);
  mxs2 = fmaxf(mxs2, r.s * r.s);
  return r;
}

// conj(a_p) . a_q over this lane's 4 rows
template <int P, int Q>
__device__ __forceinline__ void gamma_half(const Half& A, float& re, float& im) {
  p2 r = x2::mul2(A.X[P][0], A.X[Q][0]);
  p2 i = x2::mul2(A.X[P][0], A.Y[Q][0]);
  r = x2::fma2(A.Y[P][0], A.Y[Q][0], r);
  i = x2::fnma2(A.Y[P][0], A.X[Q][0], i);
  r = x2::fma2(A.X[P][1], A.X[Q][1], r);
  i = x2::fma2(A.X[P][1], A.Y[Q][1], i);
  r = x2::fma2(A.Y[P][1], A.Y[Q][1], r);
  i = x2::fnma2(A.Y[P][1], A.X[Q][1], i);
  re = x2::hsum(r);
  im = x2::hsum(i);
}

template <int P, int Q>
__device__ __forceinline__ void rot_cols(Half& A, const Rot& r) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const p2 qx = x2::fnma2s(r.phy, A.Y[Q][k], x2::mul2s(r.phx, A.X[Q][k]));
    const p2 qy = x2::fma2s(r.phy, A.X[Q][k], x2::mul2s(r.phx, A.Y[Q][k]));
    const p2 px = A.X[P][k], py = A.Y[P][k];
    A.X[Q][k] = x2::fnma2s(r.s, qx, x2::mul2s(r.c, px));
    A.Y[Q][k] = x2::fnma2s(r.s, qy, x2::mul2s(r.c, py));
    A.X[P][k] = x2::fma2s(r.c, qx, x2::mul2s(r.s, px));
    A.Y[P][k] = x2::fma2s(r.c, qy, x2::mul2s(r.s, py));
  }
}

// One round = NP (4 or 3) disjoint pairs of neighbouring columns.  Inner products: partial sums over the lane's rows + one
// xor shuffle with the other row half.  Lane half h derives the rotations of pairs 2 h and 2 h + 1 from the leaders'
// sums; every lane then fetches pair p's five scalars from lane `src0 + (p >> 1)`: the two halves of its own matrix in
// the eigen phase, of the leader matrix (m = 0) for everybody in the SVD phase.
template <int NP, int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
__device__ __forceinline__ void round_step(Half& A, float (&w)[8], float nul, float tol2, bool frozen, float& mxg2,
                                           float& mxs2, int h, int src0) {
  float gr[4], gi[4];
  gamma_half<P0, Q0>(A, gr[0], gi[0]);
  gamma_half<P1, Q1>(A, gr[1], gi[1]);
  gamma_half<P2, Q2>(A, gr[2], gi[2]);
  if (NP == 4) gamma_half<P3, Q3>(A, gr[3], gi[3]); else { gr[3] = 0.f; gi[3] = 0.f; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    gr[k] += __shfl_xor_sync(0xffffffffu, gr[k], 1);
    gi[k] += __shfl_xor_sync(0xffffffffu, gi[k], 1);
  }
  // this lane's two pairs (pair 3 of a 3-pair round: zero norms = no rotation)
  const float alA = h ? w[P2] : w[P0], beA = h ? w[Q2] : w[Q0];
  const float alB = h ? (NP == 4 ? w[P3] : 0.f) : w[P1], beB = h ? (NP == 4 ? w[Q3] : 0.f) : w[Q1];
  const Rot mA = rot_params(alA, beA, h ? gr[2] : gr[0], h ? gi[2] : gi[0], nul, tol2, frozen, mxg2, mxs2);
  const Rot mB = rot_params(alB, beB, h ? gr[3] : gr[1], h ? gi[3] : gi[1], nul, tol2, frozen, mxg2, mxs2);
  Rot r[4];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int src = src0 + (p >> 1);
    const Rot& mine = (p & 1) ? mB : mA;
    r[p].c = __shfl_sync(0xffffffffu, mine.c, src);
    r[p].s = __shfl_sync(0xffffffffu, mine.s, src);
    r[p].phx = __shfl_sync(0xffffffffu, mine.phx, src);
    r[p].phy = __shfl_sync(0xffffffffu, mine.phy, src);
    r[p].dw = __shfl_sync(0xffffffffu, mine.dw, src);
  }
  rot_cols<P0, Q0>(A, r[0]);
  rot_cols<P1, Q1>(A, r[1]);
  rot_cols<P2, Q2>(A, r[2]);
  if (NP == 4) rot_cols<P3, Q3>(A, r[3]);
  float t;
  t = w[P0]; w[P0] = w[Q0] + r[0].dw; w[Q0] = t - r[0].dw;
  t = w[P1]; w[P1] = w[Q1] + r[1].dw; w[Q1] = t - r[1].dw;
  t = w[P2]; w[P2] = w[Q2] + r[2].dw; w[Q2] = t - r[2].dw;
  if (NP == 4) { t = w[P3]; w[P3] = w[Q3] + r[3].dw; w[Q3] = t - r[3].dw; }
}

// squared column norms of the whole matrix (both row halves)
__device__ __forceinline__ void col_norms(const Half& A, float (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p2 s = x2::mul2(A.X[j][0], A.X[j][0]);
    s = x2::fma2(A.Y[j][0], A.Y[j][0], s);
    s = x2::fma2(A.X[j][1], A.X[j][1], s);
    s = x2::fma2(A.Y[j][1], A.Y[j][1], s);
    w[j] = x2::hsum(s);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] += __shfl_xor_sync(0xffffffffu, w[j], 1);
}

// One-sided Jacobi on the columns (odd-even ordering with column exchange, see bqa_fast_canon8.cu).  paired: the lanes of
// matrix m = 1 follow the rotations of matrix m = 0 of their quad.
__device__ __forceinline__ void jacobi8(Half& A, float (&w)[8], int& sweeps, bool paired, int& own_sweeps, float conv) {
  const int lane = threadIdx.x & 31;
  const int h = lane & 1;
  const int src0 = paired ? (lane & ~3) : (lane & ~1);      // lane holding row half 0 of the matrix whose rotations apply
  const float eps = 1.1920929e-07f;
  const float tol = eps * 2.f * 2.8284271f;
  const float tol2 = tol * tol;
  col_norms(A, w);
  if (paired) {                                              // followers carry the leaders' norms (they steer nothing)
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = __shfl_sync(0xffffffffu, w[j], src0);
  }
  const float fro2 = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
  const float nul = eps * eps * fro2;
  bool frozen = false;
  int done = 0;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    float mxg2 = 0.f, mxs2 = 0.f;
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
      round_step<4, 0, 1, 2, 3, 4, 5, 6, 7>(A, w, nul, tol2, frozen, mxg2, mxs2, h, src0);
      round_step<3, 1, 2, 3, 4, 5, 6, 0, 0>(A, w, nul, tol2, frozen, mxg2, mxs2, h, src0);
    }
    col_norms(A, w);
    ++done;
    own_sweeps += frozen ? 0 : 1;
    mxg2 = fmaxf(mxg2, __shfl_xor_sync(0xffffffffu, mxg2, 1));          // the two halves derived different pairs
    mxs2 = fmaxf(mxs2, __shfl_xor_sync(0xffffffffu, mxs2, 1));
    const bool fin = frozen || conv * mxg2 * mxs2 < tol2;                // xGESVJ's quadratic-convergence test
    frozen = __shfl_sync(0xffffffffu, fin ? 1 : 0, src0) != 0;
    if (paired) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = __shfl_sync(0xffffffffu, w[j], src0);
    }
    if (!__any_sync(0xffffffffu, !frozen)) break;
  }
  sweeps += done;
  if (done & 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const p2 ax = A.X[j][k], ay = A.Y[j][k];
        A.X[j][k] = A.X[7 - j][k]; A.Y[j][k] = A.Y[7 - j][k];
        A.X[7 - j][k] = ax; A.Y[7 - j][k] = ay;
      }
      const float wj = w[j];
      w[j] = w[7 - j];
      w[7 - j] = wj;
    }
  }
}

__device__ __forceinline__ void ranks_desc(const float (&v)[8], int (&rank)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += (v[i] > v[j] || (v[i] == v[j] && i < j)) ? 1 : 0;
    rank[j] = r;
  }
}

// Cholesky factor of the message in `src` (diagonal pre-sorted, zero columns for pivots at the rounding level): both
// lanes of the matrix factor it, each keeps its row half of the columns
__device__ __forceinline__ void eig_load(const unsigned char* src, unsigned char* pad, Half& A, int h) {
  float d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = *reinterpret_cast<const float*>(src + i * 64 + i * 8);
  {
    int rk[8];
    ranks_desc(d, rk);
#pragma unroll
    for (int i = 0; i < 8; ++i) pad[rk[i]] = (unsigned char)i;             // both lanes write the same bytes
  }
  __syncwarp();
  const uint2 pw = *reinterpret_cast<const uint2*>(pad);
  int roff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) roff[i] = (int)(((i < 4 ? pw.x : pw.y) >> (8 * (i & 3))) & 0xffu);
  float2 G[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) G[i][j] = *reinterpret_cast<const float2*>(src + roff[i] * 64 + roff[j] * 8);
  __syncwarp();                                             // every lane holds its input: the slots may be overwritten
  const float thr = G[0][0].x * 0x1p-22f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float dk = G[k][k].x;
    const float r = dk > thr ? rsqrt_nr(dk) : 0.f;
    G[k][k] = make_float2(dk * r, 0.f);
#pragma unroll
    for (int i = k + 1; i < 8; ++i) { G[i][k].x *= r; G[i][k].y *= r; }
#pragma unroll
    for (int j = k + 1; j < 8; ++j) {
      const float2 b = G[j][k];
#pragma unroll
      for (int i = j; i < 8; ++i) {
        const float2 a = G[i][k];
        G[i][j].x -= a.x * b.x + a.y * b.y;
        G[i][j].y -= a.y * b.x - a.x * b.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // rows 2 k, 2 k + 1 (h = 0) or 4 + 2 k, 5 + 2 k (h = 1); entries above the diagonal are zero
      const int a0 = 2 * k, a1 = 2 * k + 1, b0 = 4 + 2 * k, b1 = 5 + 2 * k;
      const float x0 = h ? (b0 >= j ? G[b0][j].x : 0.f) : (a0 >= j ? G[a0][j].x : 0.f);
      const float x1 = h ? (b1 >= j ? G[b1][j].x : 0.f) : (a1 >= j ? G[a1][j].x : 0.f);
      const float y0 = h ? (b0 >= j ? G[b0][j].y : 0.f) : (a0 >= j ? G[a0][j].y : 0.f);
      const float y1 = h ? (b1 >= j ? G[b1][j].y : 0.f) : (a1 >= j ? G[a1][j].y : 0.f);
      A.X[j][k] = x2::pk(x0, x1);
      A.Y[j][k] = x2::pk(y0, y1);
    }
}

// dst[row][c] = A[row][column of rank c]: rows back in the original order, columns by descending eigenvalue w, masked
__device__ __forceinline__ void eig_publish(const Half& A, const float (&w)[8], const unsigned char* pad, unsigned char* dst,
                                            float pinv_eps, int h) {
  int rk[8];
  ranks_desc(w, rk);
  const uint2 pw2 = *reinterpret_cast<const uint2*>(pad);
  const unsigned pwh = h ? pw2.y : pw2.x;                   // perm of rows 4 h .. 4 h + 3
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool keep = w[j] > pinv_eps && w[j] > 1.4210855e-14f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float2 xr = x2::unpk(A.X[j][k]), yi = x2::unpk(A.Y[j][k]);
      const int r0 = (int)((pwh >> (16 * k)) & 0xffu), r1 = (int)((pwh >> (16 * k + 8)) & 0xffu);
      *reinterpret_cast<float2*>(dst + r0 * 64 + rk[j] * 8) = keep ? make_float2(xr.x, yi.x) : make_float2(0.f, 0.f);
      *reinterpret_cast<float2*>(dst + r1 * 64 + rk[j] * 8) = keep ? make_float2(xr.y, yi.y) : make_float2(0.f, 0.f);
    }
  }
}

__device__ __forceinline__ void lds_row(float2 (&r)[8], const unsigned char* p) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 16 * i);
    r[2 * i] = make_float2(v.x, v.y);
    r[2 * i + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void lds_half_row(float2 (&r)[4], const unsigned char* p) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 16 * i);
    r[2 * i] = make_float2(v.x, v.y);
    r[2 * i + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void inv_col_norms(const unsigned char* m, float (&inv)[8]) {
  float w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float2 row[8];
    lds_row(row, m + r * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] += row[j].x * row[j].x + row[j].y * row[j].y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) inv[j] = w[j] > 0.f ? 1.f / w[j] : 0.f;
}

__global__ void __launch_bounds__(kWarps * 32, 1) k_canon8v3(long long L, const float2* __restrict__ ext,
                                                             float2* __restrict__ canon, float* __restrict__ lmbds,
                                                             float* __restrict__ colmax, float pinv_eps, int ncols, int nphases,
                                                             float conv) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int quad = lane >> 2, q = lane & 3, h = q & 1, m = q >> 1;
  unsigned char* wbase = smem + wib * kWarpBytes;
  unsigned char* F = wbase + quad * kPair;
  unsigned char* Bm = F + kMat;
  unsigned char* Qm = Bm + kMat;
  const long long groups = (L + kEdges - 1) / kEdges;
  const long long nwarps = (long long)gridDim.x * kWarps;
  const unsigned char* gext = reinterpret_cast<const unsigned char*>(ext);
  float cm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cm[j] = 0.f;
  int n_sweeps = 0, n_ker = 0, n_jac = 0, own_eig = 0, own_ker = 0, n_iter = 0;

  // inputs of a group by TMA: lane t < 8 copies m_f of edge t into slot B of quad t, lane 8 <= t < 16 copies m_b of
  // edge t - 8 into slot Q (512 contiguous bytes each); the warp's mbarrier completes when all 8 KB have landed
  const unsigned bar = smem_u32(smem + kBars + wib * 8);
  if (lane == 0) mbar_init(bar, 1);
  fence_proxy_async();
  __syncwarp();
  unsigned bar_parity = 0;
  auto prefetch = [&](long long g) {
    long long e = g * kEdges + (lane & 7);
    e = e < L ? e : L - 1;
    const unsigned dst = smem_u32(wbase) + (lane & 7) * kPair + (lane < 8 ? kMat : 2 * kMat);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(bar, 16 * 512);
    __syncwarp();
    if (lane < 16) bulk_g2s(dst, gext + (size_t)(lane < 8 ? e : e + L) * 512, 512, bar);
  };
  long long g = (long long)blockIdx.x * kWarps + wib;
  if (g < groups) prefetch(g);
#pragma unroll 1
  for (; g < groups; g += nwarps) {
    long long e = g * kEdges + quad;
    const bool live = e < L;
    e = live ? e : L - 1;
    mbar_wait(bar, bar_parity);
    bar_parity ^= 1;
    Half A;
    float w[8];
#pragma unroll 1
    for (int phase = 0; phase < nphases; ++phase) {
      if (phase == 0) {
        eig_load(m ? Qm : Bm, (m ? Qm : Bm) + kPad, A, h);
      } else if (m == 0) {
        // ker rows of this half: ker[i][j] = conj(sum_r A_f[r][i] A_b[r][j])
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int k = 0; k < 2; ++k) { A.X[j][k] = x2::pk(0.f, 0.f); A.Y[j][k] = x2::pk(0.f, 0.f); }
#pragma unroll 2
        for (int r = 0; r < 8; ++r) {
          float2 f[4], b[8];
          lds_half_row(f, F + r * 64 + h * 32);
          lds_row(b, Bm + r * 64);
          p2 FX[2], FY[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) { FX[k] = x2::pk(f[2 * k].x, f[2 * k + 1].x); FY[k] = x2::pk(f[2 * k].y, f[2 * k + 1].y); }
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              A.X[j][k] = x2::fma2s(b[j].x, FX[k], A.X[j][k]);
              A.X[j][k] = x2::fnma2s(b[j].y, FY[k], A.X[j][k]);
              A.Y[j][k] = x2::fnma2s(b[j].y, FX[k], A.Y[j][k]);
              A.Y[j][k] = x2::fnma2s(b[j].x, FY[k], A.Y[j][k]);
            }
        }
      } else {
        // stacked block rows of this half: conj(ul_b)[r][j] = conj(A_b[r][j]) / |column j|^2
        float inv[8];
        inv_col_norms(Bm, inv);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float2 r0[8], r1[8];
          lds_row(r0, Bm + (4 * h + 2 * k) * 64);
          lds_row(r1, Bm + (4 * h + 2 * k + 1) * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            A.X[j][k] = x2::pk(r0[j].x * inv[j], r1[j].x * inv[j]);
            A.Y[j][k] = x2::pk(-r0[j].y * inv[j], -r1[j].y * inv[j]);
          }
        }
      }
      if (phase == 1) {
        __syncwarp();
        if (g + nwarps < groups) prefetch(g + nwarps);      // slots B and Q are free: next inputs behind the SVD phase
      }
      const int before = n_sweeps;
      int own = 0;
      jacobi8(A, w, n_sweeps, phase == 1, own, conv);
      n_jac += 1;
      if (phase == 0) {
        own_eig += own;
        eig_publish(A, w, (m ? Qm : Bm) + kPad, m ? Bm : F, pinv_eps, h);
        __syncwarp();
      } else {
        own_ker += own;
        n_ker += n_sweeps - before;
      }
    }
    // ---- epilogue: w = S^2 per column (the leaders' values on every lane of the quad)
    int rk[8];
    ranks_desc(w, rk);
    float sig[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sig[j] = sqrtf(w[j]);
    if (m == 0) {
      float nrm2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) nrm2 += sig[j] > pinv_eps ? w[j] : 0.f;
      const float inrm = 1.f / sqrtf(nrm2);
      if (h == 0) {
        float sorted[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) v = rk[j] == c ? (sig[j] > pinv_eps ? sig[j] * inrm : 0.f) : v;
          sorted[c] = v;
        }
        if (live) {
          float4* lo = reinterpret_cast<float4*>(lmbds + (size_t)e * 8);
          lo[0] = make_float4(sorted[0], sorted[1], sorted[2], sorted[3]);
          lo[1] = make_float4(sorted[4], sorted[5], sorted[6], sorted[7]);
#pragma unroll
          for (int c = 0; c < 8; ++c) cm[c] = fmaxf(cm[c], sorted[c]);
        }
      }
    }
    // followers hold conj(C_b): slot e, column position = rank of column j, rows of this half
    if (m == 1) {
      float2* cb = canon + (size_t)e * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool keep = sig[j] > pinv_eps;
        if (live && rk[j] < ncols) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float2 xr = x2::unpk(A.X[j][k]), yi = x2::unpk(A.Y[j][k]);
            cb[(4 * h + 2 * k) * 8 + rk[j]] = keep ? make_float2(xr.x, -yi.x) : make_float2(0.f, 0.f);
            cb[(4 * h + 2 * k + 1) * 8 + rk[j]] = keep ? make_float2(xr.y, -yi.y) : make_float2(0.f, 0.f);
          }
        }
      }
    }
    // leaders: G[i][j] = (ker W)[i][j] / (S_j |A_f column i|^2) for the rows i of their half, then
    // C_f[r][j] = sum_i A_f[r][i] G[i][j] for the rows r of their half -- the other half of G comes by shuffle, so the
    // loop runs on every lane (the followers' registers are dead by now; only leaders store)
    {
      float invf[8];
      inv_col_norms(F, invf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float is = sig[j] > pinv_eps ? 1.f / sig[j] : 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float i0 = h ? invf[4 + 2 * k] : invf[2 * k], i1 = h ? invf[5 + 2 * k] : invf[2 * k + 1];
          const p2 sc = x2::pk(i0 * is, i1 * is);
          A.X[j][k] = x2::mul2(A.X[j][k], sc);
          A.Y[j][k] = x2::mul2(A.Y[j][k], sc);
        }
      }
    }
    {
      float2* cf = canon + (size_t)(e + L) * 64;
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = 4 * h + rr;
        float2 f[8];
        lds_row(f, F + r * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float re = 0.f, im = 0.f;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            // own half of G: rows 4 h + 2 k (+ 1); partner half: rows 4 (1 - h) + 2 k (+ 1)
            const p2 ox = A.X[j][k], oy = A.Y[j][k];
            const p2 px = __shfl_xor_sync(0xffffffffu, ox, 1), py = __shfl_xor_sync(0xffffffffu, oy, 1);
            const float2 gox = x2::unpk(ox), goy = x2::unpk(oy), gpx = x2::unpk(px), gpy = x2::unpk(py);
            const float2 fo0 = h ? f[4 + 2 * k] : f[2 * k], fo1 = h ? f[5 + 2 * k] : f[2 * k + 1];
            const float2 fp0 = h ? f[2 * k] : f[4 + 2 * k], fp1 = h ? f[2 * k + 1] : f[5 + 2 * k];
            re += fo0.x * gox.x - fo0.y * goy.x + fo1.x * gox.y - fo1.y * goy.y;
            im += fo0.x * goy.x + fo0.y * gox.x + fo1.x * goy.y + fo1.y * gox.y;
            re += fp0.x * gpx.x - fp0.y * gpy.x + fp1.x * gpx.y - fp1.y * gpy.y;
            im += fp0.x * gpy.x + fp0.y * gpx.x + fp1.x * gpy.y + fp1.y * gpx.y;
          }
          if (m == 0 && live && rk[j] < ncols) cf[r * 8 + rk[j]] = make_float2(re, im);
        }
      }
    }
    n_iter += live ? 1 : 0;
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) cm[c] = fmaxf(cm[c], __shfl_xor_sync(0xffffffffu, cm[c], o));
  }
  if (lane < 8) {
    float v = cm[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) v = lane == c ? cm[c] : v;
    atomicMax(reinterpret_cast<unsigned int*>(colmax + lane), __float_as_uint(v));
  }
  {
    unsigned long long se = h == 0 ? (unsigned long long)own_eig : 0ull, sk = q == 0 ? (unsigned long long)own_ker : 0ull;
    unsigned long long ne = h == 0 ? (unsigned long long)n_iter : 0ull, nk = q == 0 ? (unsigned long long)n_iter : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o); sk += __shfl_xor_sync(0xffffffffu, sk, o);
      ne += __shfl_xor_sync(0xffffffffu, ne, o); nk += __shfl_xor_sync(0xffffffffu, nk, o);
    }
    if (lane == 0) {
      atomicAdd(&g_stats[3], se); atomicAdd(&g_stats[4], sk); atomicAdd(&g_stats[5], ne); atomicAdd(&g_stats[6], nk);
      atomicAdd(&g_stats[0], (unsigned long long)n_jac);
      atomicAdd(&g_stats[1], (unsigned long long)n_sweeps);
      atomicAdd(&g_stats[2], (unsigned long long)n_ker);
    }
  }
}

}  // namespace canon8v3

void canon8v3_stats(unsigned long long* out7) {
  cudaMemcpyFromSymbol(out7, canon8v3::g_stats, sizeof(unsigned long long) * 7);
}

int launch_fast_canon8v3(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                         int ncols, cudaStream_t st) {
  using namespace canon8v3;
  if (L == 0) return 0;
  if (ncols < 1 || ncols > 8) return set_error("canonicalize: %d canonicalizer columns requested for n = 8", ncols);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_canon8v3, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(k_canon8v3): %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const long long groups = (L + kEdges - 1) / kEdges;
  long long grid = (groups + kWarps - 1) / kWarps;
  if (grid > sms) grid = sms;
  static const float conv = [] { const char* e = getenv("BQA_B200_CANON_CONV"); return e ? (float)atof(e) : 64.f; }();
  k_canon8v3<<<(int)grid, kWarps * 32, kSmem, st>>>(L, (const float2*)ext, (float2*)canon, (float*)lmbds, (float*)colmax,
                                                   (float)pinv_eps, ncols, 2, conv);
  return after_launch("canonicalize(n=8, v3)");
}

}  // namespace bqa
