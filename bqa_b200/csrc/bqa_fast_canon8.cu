// bqa_fast_canon8.cu -- canonicalizers of bond dimension 4 (extended dimension n = 8) in complex64.
//
// replaces _get_canonicalizers (src/bqa/state.py:171-200) for the headline shape: per undirected edge
//   m_f = V_f L_f V_f^H, m_b = V_b L_b V_b^H        masked eigendecompositions   (backends.py:483-490, 709-727)
//   ker = L_f^1/2 V_f^H conj(V_b) L_b^1/2            (state.py:186-187)
//   ker = U S W^H                                    masked SVD                   (state.py:189)
//   C_f = V_f L_f^-1/2 U (slot e + L),  C_b = V_b L_b^-1/2 conj(W) (slot e),  lambda = S / |S|   (state.py:196-200)
//
// Four lanes own an edge: lane q holds ROWS q and q + 4 of the working matrix and of the accumulated rotations, so
// a one-sided (Hestenes) Jacobi rotation of columns (p, q) is thread-local once the column inner product has been
// reduced over the 4 lanes.  Rotations follow the odd-even ordering (8 rounds of 4 / 3 neighbouring pairs per sweep).  Per
// round the 8 inner-product components (re, im of 4 pairs) are reduce-scattered with xor shuffles so that lane q
// ends up with BOTH components of pair q, derives that pair's rotation alone, and the four parameter sets are
// broadcast -- no lane repeats another lane's parameter arithmetic.  Column norms are carried along
// (alpha' = alpha - t |g|, beta' = beta + t |g|) and recomputed exactly once per sweep.  A warp holds 8 edges;
// matrices move between the "rows per lane" and "column per lane" views through padded shared-memory tiles.
// Arithmetic is packed (sm_100 FFMA2 / FMUL2: one issue slot for two fp32 FMAs, measured on B200 in
// scripts/ffma2_bench.cu): the two rows of a lane are held as pairs X[j] = (re A[q][j], re A[q+4][j]),
// Y[j] = (im A[q][j], im A[q+4][j]), so every rotation and inner product works on both rows at once.
// A sweep loop ends, per edge, with LAPACK xGESVJ's quadratic-convergence test (n max|cos(a_p, a_q)| max|sin(rotation)|
// < tol over the sweep: the next sweep would rotate by less than the tolerance), which saves the last, confirming sweep.
// A finished edge is frozen (all its later rotations are exact identities) while other edges of the warp still sweep,
// so an edge's result does not depend on which edges share its warp (single-GPU and partitioned runs stay bit-identical).
// (A warm start from the previous step's rotations was tried and dropped: A V_prev loses the relative accuracy that
// one-sided Jacobi keeps on the graded extended messages -- mean Bloch error 1.1e-4 instead of 1.3e-5 -- and the extra
// product and traffic made the step slower, not faster.)
// Code size is kept inside the 32 KB instruction cache: TWO round bodies (even and odd pairs), executed 4 times per
// sweep, and ONE Jacobi instance looped over the three matrices of an edge.
#include <cuda_runtime.h>

#include "bqa_core.cuh"
#include "bqa_f32x2.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace canon8 {

using x2::p2;

constexpr int kWarps = 4;
constexpr int kEdges = 8;                  // edges per warp (4 lanes each)
constexpr int kRow = 80;                   // 64-byte row of 8 complex + 16 bytes of padding
constexpr int kMat = 8 * kRow;             // one 8 x 8 complex tile
constexpr int kEdgeBytes = 2 * kMat + 128; // two tiles + eigenvalues of m_f, m_b (16 floats) + sorted sigma (8) + permutation (8 ints)
constexpr int kWarpBytes = kEdges * kEdgeBytes;

// statistics: [0] Jacobi problems solved (per warp: 8 matrices at a time), [1] sweeps summed over them,
// [2] of which spent on the SVD of ker (the other two problems per edge are the message eigendecompositions)
__device__ unsigned long long g_stats[3];

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float red4(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// squared column norms of the 8 x 8 matrix whose rows q, q + 4 live in this lane (packed: X = re, Y = im of both rows)
__device__ __forceinline__ void col_norms(const p2 (&X)[8], const p2 (&Y)[8], float (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = red4(x2::hsum(x2::fma2(Y[j], Y[j], x2::mul2(X[j], X[j]))));
}

// rsqrt with one Newton step (MUFU.RSQ is good to ~2 ulp; rotations must stay orthonormal to rounding)
// (MUFU.RSQ directly: `rsqrtf` wraps it in a denormal-input path of ~8 instructions; the one caller whose argument
// can be denormal pre-scales it by an even power of two, which leaves every mantissa bit unchanged)
__device__ __forceinline__ float rsqrt_nr(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (1.5f - 0.5f * x * y * y);
}

// Rotation parameters of one column pair, computed from the reduced inner product (gr, gi) and the column norms.
// Branch-free: an inactive pair (converged, or a numerically-zero column) gets the identity through selects.
struct Rot {
  float c, s, phx, phy, dw;      // cos, sin, unimodular phase conj(g)/|g|, norm transfer t |g|
};
// `mxg2` / `mxs2` collect, over the pairs this lane handled in a sweep, the largest squared cosine between two
// columns and the largest squared rotation sine (both 0 when no pair needed a rotation).  `frozen`: the edge is done.
__device__ __forceinline__ Rot rot_params(float al, float be, float gr, float gi, float nul, float tol2, bool frozen,
                                          float& mxg2, float& mxs2) {
  const float g2 = gr * gr + gi * gi;
  const float ab = al * be;
  const bool act = !frozen && !(al <= nul || be <= nul || g2 <= tol2 * ab);
  const float ig = rsqrt_nr(g2 * 0x1p60f) * 0x1p30f;   // 1 / |g|   (inf / nan when inactive: discarded below)
  const float ag = g2 * ig;
  const float zeta = 0.5f * (be - al) * ig;
  const float z2 = 1.f + zeta * zeta;
  const float t0 = __fdividef(1.f, fabsf(zeta) + z2 * rsqrt_nr(z2));   // accuracy of t only affects how well g is zeroed
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

// both rows of this lane at once: a_q <- phase a_q, then (a_p, a_q) <- (s a_p + c a_q, c a_p - s a_q): the rotation
// AND the exchange of the two columns (odd-even ordering, see jacobi8); 12 packed ops, the rotation scalars enter as
// broadcast operands
template <int P, int Q>
__device__ __forceinline__ void rot_cols(p2 (&X)[8], p2 (&Y)[8], const Rot& r) {
  const p2 qx = x2::fnma2s(r.phy, Y[Q], x2::mul2s(r.phx, X[Q]));
  const p2 qy = x2::fma2s(r.phy, X[Q], x2::mul2s(r.phx, Y[Q]));
  const p2 px = X[P], py = Y[P];
  X[Q] = x2::fnma2s(r.s, qx, x2::mul2s(r.c, px));
  Y[Q] = x2::fnma2s(r.s, qy, x2::mul2s(r.c, py));
  X[P] = x2::fma2s(r.c, qx, x2::mul2s(r.s, px));
  Y[P] = x2::fma2s(r.c, qy, x2::mul2s(r.s, py));
}
template <int P, int Q>
__device__ __forceinline__ void apply_rot(p2 (&AX)[8], p2 (&AY)[8], p2 (&VX)[8], p2 (&VY)[8], float (&w)[8],
                                          const Rot& r) {
  rot_cols<P, Q>(AX, AY, r);
  rot_cols<P, Q>(VX, VY, r);
  const float al = w[P], be = w[Q];
  w[Q] = al - r.dw;
  w[P] = be + r.dw;
}

// conj(a_p) a_q summed over the two rows of this lane
#define BQA_GAMMA(P, Q, RE, IM)                                                    \
  RE = x2::hsum(x2::fma2(AY[P], AY[Q], x2::mul2(AX[P], AX[Q])));                    \
  IM = x2::hsum(x2::fnma2(AY[P], AX[Q], x2::mul2(AX[P], AY[Q])));

// One round = NP (4 or 3) disjoint pairs of neighbouring positions (P0,Q0) .. (P3,Q3), pair k handled by lane k of the edge.
template <int NP, int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
__device__ __forceinline__ void jacobi_round(p2 (&AX)[8], p2 (&AY)[8], p2 (&VX)[8], p2 (&VY)[8],
                                             float (&w)[8], float nul, float tol2, bool frozen, float& mxg2, float& mxs2, int q) {
  float g[8];
  BQA_GAMMA(P0, Q0, g[0], g[1])
  BQA_GAMMA(P1, Q1, g[2], g[3])
  BQA_GAMMA(P2, Q2, g[4], g[5])
  if (NP == 4) {
    BQA_GAMMA(P3, Q3, g[6], g[7])
  } else {
    g[6] = 0.f; g[7] = 0.f;
  }
  const bool b1 = q & 2, b0 = q & 1;
  float h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {                             // reduce-scatter over the 4 lanes: 6 shuffles
    const float send = b1 ? g[i] : g[4 + i], keep = b1 ? g[4 + i] : g[i];
    h[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float gr = (b0 ? h[2] : h[0]) + __shfl_xor_sync(0xffffffffu, b0 ? h[0] : h[2], 1);
  const float gi = (b0 ? h[3] : h[1]) + __shfl_xor_sync(0xffffffffu, b0 ? h[1] : h[3], 1);
  // lane q now holds the inner product of pair q; its column norms (lane 3 of a 3-pair round: zero norms = no rotation)
  const float al = b1 ? (b0 ? (NP == 4 ? w[P3] : 0.f) : w[P2]) : (b0 ? w[P1] : w[P0]);
  const float be = b1 ? (b0 ? (NP == 4 ? w[Q3] : 0.f) : w[Q2]) : (b0 ? w[Q1] : w[Q0]);
  const Rot mine = rot_params(al, be, gr, gi, nul, tol2, frozen, mxg2, mxs2);
  const int base = (threadIdx.x & 31) & ~3;
  Rot R[4];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    R[k].c = __shfl_sync(0xffffffffu, mine.c, base + k);
    R[k].s = __shfl_sync(0xffffffffu, mine.s, base + k);
    R[k].phx = __shfl_sync(0xffffffffu, mine.phx, base + k);
    R[k].phy = __shfl_sync(0xffffffffu, mine.phy, base + k);
    R[k].dw = __shfl_sync(0xffffffffu, mine.dw, base + k);
  }
  apply_rot<P0, Q0>(AX, AY, VX, VY, w, R[0]);
  apply_rot<P1, Q1>(AX, AY, VX, VY, w, R[1]);
  apply_rot<P2, Q2>(AX, AY, VX, VY, w, R[2]);
  if (NP == 4) apply_rot<P3, Q3>(AX, AY, VX, VY, w, R[3]);
}

// one-sided Jacobi SVD on the packed rows: on exit A = U diag(sigma) (rows q, q + 4 of this lane), V = right singular
// vectors (same rows), w = sigma^2 per column (all lanes).  Odd-even (transposition) ordering: a sweep is 4 x [pairs
// (0,1) (2,3) (4,5) (6,7) | pairs (1,2) (3,4) (5,6)] by POSITION, and every rotation also exchanges its two columns
// (for free: the two results are written to each other's registers).  After the 8 rounds every pair of columns has met
// exactly once and the column order is reversed -- no register shuffling between rounds (the round-robin ordering used
// before spent 22 % of a round's instructions moving columns) and two round bodies of code.  An odd number of sweeps
// is undone by one reversal at the end, so the column order on exit does not depend on the sweep count.
__device__ __forceinline__ void jacobi8(p2 (&AX)[8], p2 (&AY)[8], p2 (&VX)[8], p2 (&VY)[8],
                                        float (&w)[8], int q, int& sweeps) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    VX[j] = x2::pk(j == q ? 1.f : 0.f, j == q + 4 ? 1.f : 0.f);
    VY[j] = x2::pk(0.f, 0.f);
  }
  const float eps = 1.1920929e-07f;
  const float tol = eps * 2.f * 2.8284271f;                 // eps * 2 * sqrt(n), like the generic kernel
  const float tol2 = tol * tol;
  col_norms(AX, AY, w);
  const float fro2 = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
  const float nul = eps * eps * fro2;                       // columns below eps |A|_F are numerically zero
  bool frozen = false;
  int done = 0;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    float mxg2 = 0.f, mxs2 = 0.f;
#pragma unroll 2                                            // (the back edge costs ~22 register-pair moves: amortised over 4 rounds)
    for (int rr = 0; rr < 4; ++rr) {
      jacobi_round<4, 0, 1, 2, 3, 4, 5, 6, 7>(AX, AY, VX, VY, w, nul, tol2, frozen, mxg2, mxs2, q);
      jacobi_round<3, 1, 2, 3, 4, 5, 6, 0, 0>(AX, AY, VX, VY, w, nul, tol2, frozen, mxg2, mxs2, q);
    }
    col_norms(AX, AY, w);                                   // exact norms once per sweep
    ++done;
    // per edge (4 lanes): done when nothing rotated, or when the next sweep's rotations would be below the tolerance
    mxg2 = fmaxf(mxg2, __shfl_xor_sync(0xffffffffu, mxg2, 1));
    mxs2 = fmaxf(mxs2, __shfl_xor_sync(0xffffffffu, mxs2, 1));
    mxg2 = fmaxf(mxg2, __shfl_xor_sync(0xffffffffu, mxg2, 2));
    mxs2 = fmaxf(mxs2, __shfl_xor_sync(0xffffffffu, mxs2, 2));
    frozen = frozen || 64.f * mxg2 * mxs2 < tol2;
    if (!__any_sync(0xffffffffu, !frozen)) break;
  }
  sweeps += done;
  if (done & 1) {                                           // warp-uniform: restore the original column order
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const p2 ax = AX[j], ay = AY[j], vx = VX[j], vy = VY[j];
      const float wj = w[j];
      AX[j] = AX[7 - j]; AY[j] = AY[7 - j]; VX[j] = VX[7 - j]; VY[j] = VY[7 - j]; w[j] = w[7 - j];
      AX[7 - j] = ax; AY[7 - j] = ay; VX[7 - j] = vx; VY[7 - j] = vy; w[7 - j] = wj;
    }
  }
}

// (rows q, q + 4 interleaved re/im)  <->  packed (X = re of both rows, Y = im of both rows)
__device__ __forceinline__ void pack_rows(p2 (&X)[8], p2 (&Y)[8], const float2 (&A0)[8], const float2 (&A1)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) { X[j] = x2::pk(A0[j].x, A1[j].x); Y[j] = x2::pk(A0[j].y, A1[j].y); }
}
__device__ __forceinline__ void unpack_rows(float2 (&A0)[8], float2 (&A1)[8], const p2 (&X)[8], const p2 (&Y)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 x = x2::unpk(X[j]), y = x2::unpk(Y[j]);
    A0[j] = make_float2(x.x, y.x);
    A1[j] = make_float2(x.y, y.y);
  }
}

__device__ __forceinline__ void load_row(float2 (&A)[8], const float2* src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = __ldg(s4 + i);
    A[2 * i] = make_float2(v.x, v.y);
    A[2 * i + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void lds_row(float2 (&A)[8], const unsigned char* src) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(src + 16 * i);
    A[2 * i] = make_float2(v.x, v.y);
    A[2 * i + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void sts_row(unsigned char* dst, const float2 (&A)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + 16 * i) = make_float4(A[2 * i].x, A[2 * i].y, A[2 * i + 1].x, A[2 * i + 1].y);
}
// element `idx` (0..3 or 4..7 selected by `hi`) of an 8-vector that every lane holds in registers
__device__ __forceinline__ float pick(const float (&w)[8], int q, bool hi) {
  const float a = hi ? w[4] : w[0], b = hi ? w[5] : w[1], c = hi ? w[6] : w[2], d = hi ? w[7] : w[3];
  return (q & 2) ? ((q & 1) ? d : c) : ((q & 1) ? b : a);
}

__global__ void __launch_bounds__(kWarps * 32, 4) k_canon8(long long L, const float2* __restrict__ ext,
                                                        float2* __restrict__ canon, float* __restrict__ lmbds,
                                                        float* __restrict__ colmax, float pinv_eps, int ncols) {
  __shared__ __align__(16) unsigned char smem[kWarps * kWarpBytes];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int eg = lane >> 2, q = lane & 3;                   // edge slot in the warp; this lane owns rows q and q + 4
  const int r0 = q, r1 = q + 4;
  unsigned char* tile0 = smem + wib * kWarpBytes + eg * kEdgeBytes;
  unsigned char* tile1 = tile0 + kMat;
  float* seig = reinterpret_cast<float*>(tile0 + 2 * kMat);            // eigenvalues of m_f (8) and m_b (8)
  float* ssig = seig + 16;
  int* scol = reinterpret_cast<int*>(ssig + 8);
  const long long groups = (L + kEdges - 1) / kEdges;
  const long long nwarps = (long long)gridDim.x * kWarps;
  float cm0 = 0.f, cm1 = 0.f;                               // running max of lambda[:, q] and lambda[:, q + 4]
  int n_sweeps = 0, n_jac = 0, n_ker = 0;
  for (long long g = (long long)blockIdx.x * kWarps + wib; g < groups; g += nwarps) {
    long long e = g * kEdges + eg;
    const bool live = e < L;
    e = live ? e : L - 1;
    float2 A0[8], A1[8], W0[8], W1[8];
    float sk[8];
    // m = 0: eigenvectors of m_f -> tile0, m = 1: eigenvectors of m_b -> tile1, m = 2: SVD of ker
#pragma unroll 1
    for (int m = 0; m < 3; ++m) {
      if (m < 2) {
        const float2* src = ext + (size_t)(e + (m ? L : 0)) * 64;
        load_row(A0, src + r0 * 8);
        load_row(A1, src + r1 * 8);
      } else {
        // ker[i][j] = sqrt(sf_i sb_j) sum_k conj(Vf[k][i]) conj(Vb[k][j]) with masked eigenvalues: this lane builds
        // rows i = q, q + 4 from columns q, q + 4 of Vf and whole rows of Vb in the tiles
        const float ef0 = seig[r0], ef1 = seig[r1];
        const float rsf0 = ef0 > pinv_eps ? sqrtf(ef0) : 0.f, rsf1 = ef1 > pinv_eps ? sqrtf(ef1) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { A0[j] = make_float2(0.f, 0.f); A1[j] = make_float2(0.f, 0.f); }
#pragma unroll 2
        for (int k = 0; k < 8; ++k) {
          const float2 vf0 = *reinterpret_cast<const float2*>(tile0 + k * kRow + r0 * 8);
          const float2 vf1 = *reinterpret_cast<const float2*>(tile0 + k * kRow + r1 * 8);
          float2 row[8];
          lds_row(row, tile1 + k * kRow);
#pragma unroll
          for (int j = 0; j < 8; ++j) {                     // conj(vf) conj(vb) = conj(vf vb)
            A0[j].x += vf0.x * row[j].x - vf0.y * row[j].y;
            A0[j].y -= vf0.x * row[j].y + vf0.y * row[j].x;
            A1[j].x += vf1.x * row[j].x - vf1.y * row[j].y;
            A1[j].y -= vf1.x * row[j].y + vf1.y * row[j].x;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float eb = seig[8 + j];
          const float sb = eb > pinv_eps ? sqrtf(eb) : 0.f;
          A0[j].x *= rsf0 * sb; A0[j].y *= rsf0 * sb;
          A1[j].x *= rsf1 * sb; A1[j].y *= rsf1 * sb;
        }
      }
      const int before = n_sweeps;
      {
        p2 AX[8], AY[8], VX[8], VY[8];
        pack_rows(AX, AY, A0, A1);
        jacobi8(AX, AY, VX, VY, sk, q, n_sweeps);
        unpack_rows(A0, A1, AX, AY);
        unpack_rows(W0, W1, VX, VY);
      }
      if (m == 2) n_ker += n_sweeps - before;
      if (m < 2) {
        unsigned char* tile = m ? tile1 : tile0;
        __syncwarp();
        sts_row(tile + r0 * kRow, W0);
        sts_row(tile + r1 * kRow, W1);
        seig[m * 8 + r0] = sqrtf(pick(sk, q, false));       // eigenvalue = singular value of the PSD message
        seig[m * 8 + r1] = sqrtf(pick(sk, q, true));
        __syncwarp();
      }
    }
    n_jac += 3;
    // this lane's rows of V_f L_f^-1/2 and V_b L_b^-1/2 (masked like pinv_raw, backends.py:719-727)
    float2 Vf0[8], Vf1[8], Vb0[8], Vb1[8];
    lds_row(Vf0, tile0 + r0 * kRow);
    lds_row(Vf1, tile0 + r1 * kRow);
    lds_row(Vb0, tile1 + r0 * kRow);
    lds_row(Vb1, tile1 + r1 * kRow);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ef = seig[j], eb = seig[8 + j];
      const float isf = (ef > pinv_eps && sqrtf(ef) > 1.1920929e-07f) ? rsqrtf(ef) : 0.f;
      const float isb = (eb > pinv_eps && sqrtf(eb) > 1.1920929e-07f) ? rsqrtf(eb) : 0.f;
      Vf0[j].x *= isf; Vf0[j].y *= isf; Vf1[j].x *= isf; Vf1[j].y *= isf;
      Vb0[j].x *= isb; Vb0[j].y *= isb; Vb1[j].x *= isb; Vb1[j].y *= isb;
    }
    // sort the singular values (descending, stable) and publish A = U S and W through the tiles
    const float mine0 = pick(sk, q, false), mine1 = pick(sk, q, true);
    int rank0 = 0, rank1 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rank0 += (sk[j] > mine0 || (sk[j] == mine0 && j < r0)) ? 1 : 0;
      rank1 += (sk[j] > mine1 || (sk[j] == mine1 && j < r1)) ? 1 : 0;
    }
    __syncwarp();
    ssig[rank0] = sqrtf(mine0); scol[rank0] = r0;
    ssig[rank1] = sqrtf(mine1); scol[rank1] = r1;
    sts_row(tile0 + r0 * kRow, A0);
    sts_row(tile0 + r1 * kRow, A1);
    sts_row(tile1 + r0 * kRow, W0);
    sts_row(tile1 + r1 * kRow, W1);
    __syncwarp();
    // lambda = masked S / |masked S|
    float nrm2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = ssig[j];
      const float sm = s > pinv_eps ? s : 0.f;
      nrm2 += sm * sm;
    }
    const float inrm = 1.f / sqrtf(nrm2);
    const float s0 = ssig[r0], s1 = ssig[r1];
    const float lam0 = (s0 > pinv_eps ? s0 : 0.f) * inrm, lam1 = (s1 > pinv_eps ? s1 : 0.f) * inrm;
    if (live) {
      lmbds[(size_t)e * 8 + r0] = lam0;
      lmbds[(size_t)e * 8 + r1] = lam1;
      cm0 = fmaxf(cm0, lam0);
      cm1 = fmaxf(cm1, lam1);
    }
    // C_f[r][c] = sum_i Vf[r][i] isf_i U[i][col_c],  U[i][col] = A[i][col] / S_col   (slot e + L)
    // C_b[r][c] = sum_j Vb[r][j] isb_j conj(W[j][col_c])                               (slot e)
    for (int c = 0; c < ncols; c += 2) {
      float2 cf0[2], cf1[2], cb0[2], cb1[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = c + h < 8 ? c + h : 7;
        const int col = scol[cc];
        const float s = ssig[cc];
        float2 f0 = make_float2(0.f, 0.f), f1 = f0, b0 = f0, b1 = f0;
        if (s > pinv_eps) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 u = *reinterpret_cast<const float2*>(tile0 + i * kRow + col * 8);
            const float2 w = *reinterpret_cast<const float2*>(tile1 + i * kRow + col * 8);
            f0.x += Vf0[i].x * u.x - Vf0[i].y * u.y; f0.y += Vf0[i].x * u.y + Vf0[i].y * u.x;
            f1.x += Vf1[i].x * u.x - Vf1[i].y * u.y; f1.y += Vf1[i].x * u.y + Vf1[i].y * u.x;
            b0.x += Vb0[i].x * w.x + Vb0[i].y * w.y; b0.y += Vb0[i].y * w.x - Vb0[i].x * w.y;     // times conj(w)
            b1.x += Vb1[i].x * w.x + Vb1[i].y * w.y; b1.y += Vb1[i].y * w.x - Vb1[i].x * w.y;
          }
          const float is = 1.f / s;
          f0.x *= is; f0.y *= is; f1.x *= is; f1.y *= is;
        }
        cf0[h] = f0; cf1[h] = f1; cb0[h] = b0; cb1[h] = b1;
      }
      if (live) {
        float2* cf = canon + (size_t)(e + L) * 64 + c;
        float2* cb = canon + (size_t)e * 64 + c;
        *reinterpret_cast<float4*>(cf + r0 * 8) = make_float4(cf0[0].x, cf0[0].y, cf0[1].x, cf0[1].y);
        *reinterpret_cast<float4*>(cf + r1 * 8) = make_float4(cf1[0].x, cf1[0].y, cf1[1].x, cf1[1].y);
        *reinterpret_cast<float4*>(cb + r0 * 8) = make_float4(cb0[0].x, cb0[0].y, cb0[1].x, cb0[1].y);
        *reinterpret_cast<float4*>(cb + r1 * 8) = make_float4(cb1[0].x, cb1[0].y, cb1[1].x, cb1[1].y);
      }
    }
    __syncwarp();
  }
  // column-wise max of lambda over all edges (truncate_lmbds, backends.py:297-299)
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, o));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, o));
  }
  if (lane < 4) {
    atomicMax(reinterpret_cast<unsigned int*>(colmax + lane), __float_as_uint(cm0));
    atomicMax(reinterpret_cast<unsigned int*>(colmax + lane + 4), __float_as_uint(cm1));
  }
  if (lane == 0) {
    atomicAdd(&g_stats[0], (unsigned long long)n_jac);
    atomicAdd(&g_stats[1], (unsigned long long)n_sweeps);
    atomicAdd(&g_stats[2], (unsigned long long)n_ker);
  }
}

}  // namespace canon8

void canon8_stats(unsigned long long* out3) {
  cudaMemcpyFromSymbol(out3, canon8::g_stats, sizeof(unsigned long long) * 3);
}

int launch_fast_canon8(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                       int ncols, cudaStream_t st) {
  using namespace canon8;
  if (L == 0) return 0;
  if (ncols < 1 || ncols > 8) return set_error("canonicalize: %d canonicalizer columns requested for n = 8", ncols);
  ncols = (ncols + 1) & ~1;
  const long long groups = (L + kEdges - 1) / kEdges;
  long long grid = (groups + kWarps - 1) / kWarps;
  const long long cap = 148LL * 16;
  if (grid > cap) grid = cap;
  k_canon8<<<(int)grid, kWarps * 32, 0, st>>>(L, (const float2*)ext, (float2*)canon, (float*)lmbds, (float*)colmax,
                                              (float)pinv_eps, ncols);
  return after_launch("canonicalize(n=8)");
}

}  // namespace bqa
