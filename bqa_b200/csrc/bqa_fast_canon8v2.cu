// bqa_fast_canon8v2.cu -- canonicalizers of bond dimension 4 (extended dimension n = 8) in complex64, second design.
//
// replaces _get_canonicalizers (src/bqa/state.py:171-200) for the headline shape: per undirected edge
//   m_f = V_f L_f V_f^H, m_b = V_b L_b V_b^H        masked eigendecompositions   (backends.py:483-490, 709-727)
//   ker = L_f^1/2 V_f^H conj(V_b) L_b^1/2            (state.py:186-187)
//   ker = U S W^H                                    masked SVD                   (state.py:189)
//   C_f = V_f L_f^-1/2 U (slot e + L),  C_b = V_b L_b^-1/2 conj(W) (slot e),  lambda = S / |S|   (state.py:196-200)
//
// What changed against bqa_fast_canon8.cu (three one-sided Jacobi SVDs per edge, V accumulated in all of them, four
// lanes per matrix) -- each point measured first in a complex64 numpy emulation on extended messages of real anneals
// (scratch/canon_proto*.py, numbers in DESIGN.md section 3.2):
//   * The two eigenproblems run on the CHOLESKY FACTOR.  A message is Hermitian positive semi-definite: m = G G^H with G
//     lower triangular (diagonal pre-sorted, non-positive pivots give zero columns), and one-sided Jacobi on the columns
//     of G gives G J = Q diag(sigma) with m = Q diag(sigma^2) Q^H (Veselic-Hari).  The eigenvectors are the normalised
//     columns -- NO accumulation of the rotations, half the rotation work -- and because the columns of G are already
//     nearly orthogonal and graded the sweep count drops from 5.8 to 3.9 per problem with a very narrow spread (the
//     slowest of 32 matrices: 4.0 instead of 6.9).  The emulated end-to-end Bloch error against the complex128
//     reference is 6 times SMALLER than with Jacobi on the message itself (7.8e-6 vs 4.9e-5 mean).
//   * lu = L^1/2 V^H is simply (G J)^H and ul = V L^-1/2 is (G J) diag(1 / sigma^2): no square roots, no normalisation.
//   * The SVD of ker rotates a STACKED matrix [ker ; conj(ul_b)]: the lower block arrives as conj(C_b) = conj(ul_b) W,
//     no accumulated W and no product afterwards; C_f = ul_f (ker W) diag(1 / S).
//   * ONE lane owns a whole 8 x 8 matrix (128 registers as packed row pairs), so a rotation of columns (p, q) needs no
//     shuffle, no reduce-scatter and no broadcast; the four pairs of a round are independent instruction streams.  A
//     warp works on 16 edges: in the eigen phase lane 2i takes m_f and lane 2i + 1 takes m_b of edge i, in the SVD phase
//     lane 2i holds ker ("leader": computes the rotations) and lane 2i + 1 the stacked block ("follower": receives the
//     four rotation scalars by shuffle).
// Results do not depend on which edges share a warp (a converged matrix is frozen: its later rotations are exact
// column exchanges), so single-GPU and partitioned runs stay bit-identical.
#include <cuda_runtime.h>

#include <cstdlib>

#include "bqa_core.cuh"
#include "bqa_f32x2.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace canon8v2 {

using x2::p2;

constexpr int kWarps = 8;                  // one persistent CTA per SM, 2 warps per sub-partition (the kernel lives on
                                           // ~250 registers: a whole 8 x 8 complex matrix per lane)
constexpr int kEdges = 16;                 // edges per warp iteration (2 lanes each)
constexpr int kMat = 528;                  // 8 rows x 64 bytes + 16: an odd number of 16-byte units, so the 8 lanes of a
                                           // quarter warp hit 8 different bank groups with 128-bit accesses
constexpr int kPad = 512;                  // offset of the 16 spare bytes of a matrix slot (row permutation of the owner)
constexpr int kPair = 3 * kMat;            // per edge: F (A_f sorted, lives to the epilogue) | B (input m_f, then A_b) | Q (input m_b)
constexpr int kWarpBytes = kEdges * kPair;
constexpr int kBars = kWarps * kWarpBytes; // one mbarrier per warp behind the matrix slots
constexpr int kSmem = kBars + kWarps * 8;  // 202 816 bytes
static_assert(kSmem <= 227 * 1024, "shared memory budget");

// statistics: [0] warp-level Jacobi runs (32 or 16 matrices at a time), [1] sweeps summed over them (a warp sweeps until
// its slowest matrix has converged), [2] the part of [1] spent on the SVD of ker
// [3] / [4]: sweeps summed over the single matrices until each one froze, eigen phase / SVD phase ([5] / [6]: matrices)
__device__ unsigned long long g_stats[7];
__device__ unsigned long long g_span[2] = {~0ull, 0ull};   // profiling aid: earliest start / latest end (%globaltimer, ns) of the last launches

// Multi-GPU, one owner per cut edge (bqa_b200_canonicalize_p2p): the owner stores the edge's canonicalizers and lambdas
// into the other endpoint's rank as well.  remote[e] = -1 or (peer << 27 | that rank's index of the edge).
struct Remote {
  const int* code;                          // per local edge, or null
  float2* canon[BQA_MAX_PEERS];             // peer-mapped bases of the ranks' canonicalizer arrays
  float* lmbds[BQA_MAX_PEERS];
  long long L[BQA_MAX_PEERS];               // local edge counts of the ranks (slot of C_f = index + L)
};

struct Mat {                               // column j, row pair k = (rows 2k, 2k + 1): X = real parts, Y = imaginary parts
  p2 X[8][4], Y[8][4];
};

__device__ __forceinline__ float rsqrt_nr(float x) {       // MUFU.RSQ + one Newton step (see bqa_fast_canon8.cu)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (1.5f - 0.5f * x * y * y);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem) : "memory");
}
// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) completed on an mbarrier -------------------------------------
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
// generic-proxy accesses of this thread to shared memory are ordered before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Rot {
  float c, s, phx, phy, dw;                // cos, sin, unimodular phase conj(g) / |g|, norm transfer t |g|
};
// rotation of one column pair from its inner product (gr, gi) and squared norms (same arithmetic as bqa_fast_canon8.cu)
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

// conj(a_p) . a_q over the 8 rows
template <int P, int Q>
__device__ __forceinline__ void gamma(const Mat& A, float& re, float& im) {
  p2 r = x2::mul2(A.X[P][0], A.X[Q][0]);
  p2 i = x2::mul2(A.X[P][0], A.Y[Q][0]);
  r = x2::fma2(A.Y[P][0], A.Y[Q][0], r);
  i = x2::fnma2(A.Y[P][0], A.X[Q][0], i);
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    r = x2::fma2(A.X[P][k], A.X[Q][k], r);
    i = x2::fma2(A.X[P][k], A.Y[Q][k], i);
    r = x2::fma2(A.Y[P][k], A.Y[Q][k], r);
    i = x2::fnma2(A.Y[P][k], A.X[Q][k], i);
  }
  re = x2::hsum(r);
  im = x2::hsum(i);
}

// a_q <- phase a_q, then (a_p, a_q) <- (s a_p + c a_q, c a_p - s a_q): rotation AND exchange of the two columns
template <int P, int Q>
__device__ __forceinline__ void rot_cols(Mat& A, const Rot& r) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const p2 qx = x2::fnma2s(r.phy, A.Y[Q][k], x2::mul2s(r.phx, A.X[Q][k]));
    const p2 qy = x2::fma2s(r.phy, A.X[Q][k], x2::mul2s(r.phx, A.Y[Q][k]));
    const p2 px = A.X[P][k], py = A.Y[P][k];
    A.X[Q][k] = x2::fnma2s(r.s, qx, x2::mul2s(r.c, px));
    A.Y[Q][k] = x2::fnma2s(r.s, qy, x2::mul2s(r.c, py));
    A.X[P][k] = x2::fma2s(r.c, qx, x2::mul2s(r.s, px));
    A.Y[P][k] = x2::fma2s(r.c, qy, x2::mul2s(r.s, py));
  }
}

// One round = NP (4 or 3) disjoint pairs of neighbouring columns.  Every lane derives the rotations from its own columns
// and then takes the four rotation scalars of lane `src_lane`: itself in the eigen phase, its leader (lane - 1) for the
// follower lanes of the SVD phase.  The four inner products, then the four parameter chains, then the rotations: the
// independent chains sit next to each other for the scheduler.  (ONE code body for both phases: with two instantiations
// warps in different phases evicted each other's 16 KB loop from the 32 KB instruction cache -- 14 % misses, "no
// instruction" the second largest stall in the first ncu capture.)
template <int NP, int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
__device__ __forceinline__ void round_step(Mat& A, float (&w)[8], float nul, float tol2, bool frozen, float& mxg2,
                                           float& mxs2, int src_lane) {
  float gr[4], gi[4];
  gamma<P0, Q0>(A, gr[0], gi[0]);
  gamma<P1, Q1>(A, gr[1], gi[1]);
  gamma<P2, Q2>(A, gr[2], gi[2]);
  if (NP == 4) gamma<P3, Q3>(A, gr[3], gi[3]);
  Rot r[4];
  r[0] = rot_params(w[P0], w[Q0], gr[0], gi[0], nul, tol2, frozen, mxg2, mxs2);
  r[1] = rot_params(w[P1], w[Q1], gr[1], gi[1], nul, tol2, frozen, mxg2, mxs2);
  r[2] = rot_params(w[P2], w[Q2], gr[2], gi[2], nul, tol2, frozen, mxg2, mxs2);
  if (NP == 4) r[3] = rot_params(w[P3], w[Q3], gr[3], gi[3], nul, tol2, frozen, mxg2, mxs2);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    r[k].c = __shfl_sync(0xffffffffu, r[k].c, src_lane);
    r[k].s = __shfl_sync(0xffffffffu, r[k].s, src_lane);
    r[k].phx = __shfl_sync(0xffffffffu, r[k].phx, src_lane);
    r[k].phy = __shfl_sync(0xffffffffu, r[k].phy, src_lane);
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

__device__ __forceinline__ void col_norms(const Mat& A, float (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p2 s = x2::mul2(A.X[j][0], A.X[j][0]);
    s = x2::fma2(A.Y[j][0], A.Y[j][0], s);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      s = x2::fma2(A.X[j][k], A.X[j][k], s);
      s = x2::fma2(A.Y[j][k], A.Y[j][k], s);
    }
    w[j] = x2::hsum(s);
  }
}

// One-sided Jacobi on the columns of A, odd-even (transposition) ordering with column exchange as in
// bqa_fast_canon8.cu: a sweep is 4 x [pairs (0,1) (2,3) (4,5) (6,7) | pairs (1,2) (3,4) (5,6)] by position; after a sweep
// the column order is reversed, an odd number of sweeps is undone at the end.  On exit the columns are orthogonal and
// w[j] is the squared norm of column j.  paired: lanes 2i (leader) and 2i + 1 (follower, see pair_step).
__device__ __forceinline__ void jacobi8(Mat& A, float (&w)[8], int& sweeps, bool paired, int& own_sweeps, float conv) {
  const int lane = threadIdx.x & 31;
  const int src_lane = paired ? (lane & ~1) : lane;
  const float eps = 1.1920929e-07f;
  const float tol = eps * 2.f * 2.8284271f;                 // eps * 2 * sqrt(n), like the generic kernel
  const float tol2 = tol * tol;
  col_norms(A, w);
  const float fro2 = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
  const float nul = eps * eps * fro2;                       // columns below eps |A|_F are numerically zero
  bool frozen = false;
  int done = 0;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    float mxg2 = 0.f, mxs2 = 0.f;
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
      round_step<4, 0, 1, 2, 3, 4, 5, 6, 7>(A, w, nul, tol2, frozen, mxg2, mxs2, src_lane);
      round_step<3, 1, 2, 3, 4, 5, 6, 0, 0>(A, w, nul, tol2, frozen, mxg2, mxs2, src_lane);
    }
    col_norms(A, w);                                        // exact norms once per sweep
    ++done;
    own_sweeps += frozen ? 0 : 1;
    // LAPACK xGESVJ's quadratic-convergence test: the next sweep's rotations would be below the tolerance
    const bool fin = frozen || conv * mxg2 * mxs2 < tol2;
    frozen = __shfl_sync(0xffffffffu, fin ? 1 : 0, src_lane) != 0;
    if (!__any_sync(0xffffffffu, !frozen)) break;
  }
  sweeps += done;
  if (done & 1) {                                           // warp-uniform: restore the original column order
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
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

// rank of every entry in the descending, stable order of v[0..7]
__device__ __forceinline__ void ranks_desc(const float (&v)[8], int (&rank)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += (v[i] > v[j] || (v[i] == v[j] && i < j)) ? 1 : 0;
    rank[j] = r;
  }
}

// ---- phase 1: eigendecomposition of one Hermitian PSD message through its Cholesky factor --------------------------
// eig_load: the 8 x 8 message in `src` (row-major, 64-byte rows) -> columns of its Cholesky factor G (m = G G^H) in A.
// Rows / columns are visited in the order of decreasing diagonal (a static stand-in for diagonal pivoting); the
// permutation perm[pos] = original index is kept in the spare bytes of the owner's input slot.
__device__ __forceinline__ void eig_load(const unsigned char* src, unsigned char* pad, Mat& A) {
  float d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = *reinterpret_cast<const float*>(src + i * 64 + i * 8);
  {
    int rk[8];
    ranks_desc(d, rk);
#pragma unroll
    for (int i = 0; i < 8; ++i) pad[rk[i]] = (unsigned char)i;
  }
  const uint2 pw = *reinterpret_cast<const uint2*>(pad);
  int roff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) roff[i] = (int)(((i < 4 ? pw.x : pw.y) >> (8 * (i & 3))) & 0xffu);
  // lower triangle of the permuted matrix and its Cholesky factor, in place (right-looking, fully unrolled)
  float2 G[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) G[i][j] = *reinterpret_cast<const float2*>(src + roff[i] * 64 + roff[j] * 8);
  __syncwarp();                                             // every lane holds its input: the slots may be overwritten
  const float thr = G[0][0].x * 0x1p-22f;                   // pivots at the rounding level of the largest one: zero column
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
      for (int i = j; i < 8; ++i) {                          // G[i][j] -= G[i][k] conj(G[j][k])
        const float2 a = G[i][k];
        G[i][j].x -= a.x * b.x + a.y * b.y;
        G[i][j].y -= a.y * b.x - a.x * b.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i0 = 2 * k, i1 = 2 * k + 1;
      A.X[j][k] = x2::pk(i0 >= j ? G[i0][j].x : 0.f, i1 >= j ? G[i1][j].x : 0.f);
      A.Y[j][k] = x2::pk(i0 >= j ? G[i0][j].y : 0.f, i1 >= j ? G[i1][j].y : 0.f);
    }
}

// eig_publish: after the Jacobi sweeps the columns of A are Q diag(sigma), w = sigma^2 = the eigenvalues.
// dst[row][c] = A[row][column of rank c]: rows back in the original order, columns by descending eigenvalue, columns
// with eigenvalue <= pinv_eps (or below the pinv_raw cut) zeroed: dst^H is lu, dst diag(1 / |column|^2) is ul of
// decompose_iden_using_msgs (backends.py:483-490)
__device__ __forceinline__ void eig_publish(const Mat& A, const float (&w)[8], const unsigned char* pad, unsigned char* dst,
                                            float pinv_eps) {
  int rk[8];
  ranks_desc(w, rk);
  const uint2 pw2 = *reinterpret_cast<const uint2*>(pad);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool keep = w[j] > pinv_eps && w[j] > 1.4210855e-14f;    // batched_svd mask; pinv_raw: sqrt(lambda) > eps
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xr = x2::unpk(A.X[j][k]), yi = x2::unpk(A.Y[j][k]);
      const int r0 = (int)(((k < 2 ? pw2.x : pw2.y) >> (8 * ((2 * k) & 3))) & 0xffu);
      const int r1 = (int)(((k < 2 ? pw2.x : pw2.y) >> (8 * ((2 * k + 1) & 3))) & 0xffu);
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

// 1 / |column|^2 of a published matrix (0 for the masked, zeroed columns)
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

// n edges are processed: order[0 .. n) (or 0 .. n - 1 without `order`); L = local edge count (slot of m_b / C_f = e + L).
// REMOTE: the multi-GPU instantiation that also stores into the peers (compiled separately: with the peer stores in the
// one kernel the single-GPU launch measured 0.94 ms instead of 0.85 ms)
template <bool REMOTE>
__global__ void __launch_bounds__(kWarps * 32, 1) k_canon8v2(long long n, long long L, const float2* __restrict__ ext,
                                                             float2* __restrict__ canon, float* __restrict__ lmbds,
                                                             float* __restrict__ colmax, float pinv_eps, int ncols, int nphases,
                                                             const int* __restrict__ order, unsigned char* __restrict__ cost, float conv,
                                                             const __grid_constant__ Remote rem) {
  extern __shared__ __align__(16) unsigned char smem[];
  // Bookkeeping that lives across the Jacobi loops sits in shared memory, not in registers (the kernel runs at the 255
  // register limit: every long-lived scalar became a spill inside the sweep loop): per warp 8 column maxima + 6 counters
  __shared__ unsigned s_cm[kWarps][8];
  __shared__ int s_cnt[kWarps][8];          // sweeps, ker sweeps, Jacobi runs, own eig sweeps, own ker sweeps, live edges
  __shared__ unsigned long long s_t0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(s_t0));
  if ((threadIdx.x & 31) < 8) { s_cm[threadIdx.x >> 5][threadIdx.x & 31] = 0u; s_cnt[threadIdx.x >> 5][threadIdx.x & 31] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = lane >> 1;
  const bool leader = (lane & 1) == 0;
  unsigned char* wbase = smem + wib * kWarpBytes;
  unsigned char* F = wbase + pair * kPair;                  // A_f (published), needed until the epilogue
  unsigned char* Bm = F + kMat;                             // input m_f, then A_b (published)
  unsigned char* Qm = Bm + kMat;                            // input m_b
  const int groups = (int)((n + kEdges - 1) / kEdges);
  const int nwarps = (int)gridDim.x * kWarps;
  const unsigned char* gext = reinterpret_cast<const unsigned char*>(ext);

  // The 32 input matrices of a group (m_f of its 16 edges, then m_b) arrive by TMA: every lane issues ONE 512-byte bulk
  // copy (lane t < 16: m_f of edge t -> slot B of pair t; lane t >= 16: m_b of edge t - 16 -> slot Q) and the warp's
  // mbarrier completes when all 16 KB have landed -- instead of 32 16-byte LDGSTS per lane.
  const unsigned bar = smem_u32(smem + kBars + wib * 8);
  if (lane == 0) mbar_init(bar, 1);
  fence_proxy_async();
  __syncwarp();
  unsigned bar_parity = 0;
  auto prefetch = [&](int g) {
    long long e = (long long)g * kEdges + (lane & 15);
    e = e < n ? e : n - 1;
    if (order) e = __ldg(order + e);
    const unsigned dst = smem_u32(wbase) + (lane & 15) * kPair + (lane < 16 ? kMat : 2 * kMat);
    fence_proxy_async();                                    // this lane's earlier loads / stores of the slots come first
    __syncwarp();
    if (lane == 0) mbar_expect_tx(bar, 32 * 512);
    __syncwarp();
    bulk_g2s(dst, gext + (size_t)(lane < 16 ? e : e + L) * 512, 512, bar);
  };
  int g = (int)blockIdx.x * kWarps + wib;
  if (g < groups) prefetch(g);
#pragma unroll 1
  for (; g < groups; g += nwarps) {
    int it_eig = 0, it_ker = 0;
    mbar_wait(bar, bar_parity);
    bar_parity ^= 1;
    // Two Jacobi runs per iteration through ONE copy of the sweep code (a rolled loop over the phases):
    //   phase 0: lane 2i decomposes m_f (slot B -> F), lane 2i + 1 decomposes m_b (slot Q -> B)
    //   phase 1: lane 2i holds ker = A_f^H conj(A_b), lane 2i + 1 the stacked block conj(ul_b) = conj(A_b) / |col|^2
    Mat A;
    float w[8];
#pragma unroll 1
    for (int phase = 0; phase < nphases; ++phase) {
      if (phase == 0) {
        // (the follower publishes into B, which the leader reads its input from: eig_load holds the whole input in
        // registers and passes a __syncwarp before anything is stored)
        eig_load(leader ? Bm : Qm, (leader ? Bm : Qm) + kPad, A);
      } else if (leader) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) { A.X[j][k] = x2::pk(0.f, 0.f); A.Y[j][k] = x2::pk(0.f, 0.f); }
#pragma unroll 2
        for (int r = 0; r < 8; ++r) {
          float2 f[8], b[8];
          lds_row(f, F + r * 64);
          lds_row(b, Bm + r * 64);
          p2 FX[4], FY[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { FX[k] = x2::pk(f[2 * k].x, f[2 * k + 1].x); FY[k] = x2::pk(f[2 * k].y, f[2 * k + 1].y); }
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {                    // ker[i][j] = conj(sum_r A_f[r][i] A_b[r][j])
              A.X[j][k] = x2::fma2s(b[j].x, FX[k], A.X[j][k]);
              A.X[j][k] = x2::fnma2s(b[j].y, FY[k], A.X[j][k]);
              A.Y[j][k] = x2::fnma2s(b[j].y, FX[k], A.Y[j][k]);
              A.Y[j][k] = x2::fnma2s(b[j].x, FY[k], A.Y[j][k]);
            }
        }
      } else {
        float inv[8];
        inv_col_norms(Bm, inv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 r0[8], r1[8];
          lds_row(r0, Bm + (2 * k) * 64);
          lds_row(r1, Bm + (2 * k + 1) * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            A.X[j][k] = x2::pk(r0[j].x * inv[j], r1[j].x * inv[j]);
            A.Y[j][k] = x2::pk(-r0[j].y * inv[j], -r1[j].y * inv[j]);
          }
        }
      }
      if (phase == 1) {
        __syncwarp();
        // slots B and Q are free: the next group's inputs stream in behind the SVD phase
        if (g + nwarps < groups) prefetch(g + nwarps);
      }
      int own = 0, ran = 0;
      jacobi8(A, w, ran, phase == 1, own, conv);
      if (phase == 0) it_eig = max(own, __shfl_xor_sync(0xffffffffu, own, 1)); else it_ker = own;
      if (lane == 0) { s_cnt[wib][0] += ran; s_cnt[wib][1] += phase == 1 ? ran : 0; s_cnt[wib][2] += 1; }
      atomicAdd(&s_cnt[wib][phase == 0 ? 3 : 4], (phase == 0 || leader) ? own : 0);
      if (phase == 0) {
        eig_publish(A, w, (leader ? Bm : Qm) + kPad, leader ? F : Bm, pinv_eps);
        __syncwarp();
      }
    }
    // the edge of this lane pair (re-derived here: nothing but the matrix lives across the sweeps)
    long long e = (long long)g * kEdges + pair;
    const bool live = e < n;
    e = live ? e : n - 1;
    if (order) e = __ldg(order + e);                        // the edges this rank owns / edges grouped by cost
    // a cut edge this rank owns: the other endpoint's rank gets the same results
    const int rcode = (REMOTE && rem.code && live) ? __ldg(rem.code + e) : -1;
    float2* far_canon = nullptr;
    float* far_lmbds = nullptr;
    long long far_e = 0, far_L = 0;
    if (rcode >= 0) {
      const int q = rcode >> 27;
      far_e = rcode & ((1 << 27) - 1);
      far_L = rem.L[q];
      far_canon = rem.canon[q];
      far_lmbds = rem.lmbds[q];
    }
    // ---- epilogue.  Leader: w = S^2 per column
    int rk[8];
    ranks_desc(w, rk);
    unsigned packed = 0;
    float sig[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sig[j] = sqrtf(w[j]);
      packed |= (unsigned)rk[j] << (3 * j);
      packed |= (sig[j] > pinv_eps ? 1u : 0u) << (24 + j);
    }
    packed = __shfl_sync(0xffffffffu, packed, lane & ~1);     // ranks and masks of the leader
    if (leader) {
      float nrm2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) nrm2 += sig[j] > pinv_eps ? w[j] : 0.f;
      const float inrm = 1.f / sqrtf(nrm2);
      float lam[8], sorted[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lam[j] = sig[j] > pinv_eps ? sig[j] * inrm : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v = rk[j] == c ? lam[j] : v;
        sorted[c] = v;
      }
      // what this edge cost: the key bqa_b200_sort_edges_by_cost groups the edges by, so that the edges of a warp need
      // about the same number of sweeps next time (a warp sweeps until its slowest matrix has converged)
      if (live && cost) cost[e] = (unsigned char)(min(it_eig, 15) * 16 + min(it_ker, 15));
      if (live) {
        float4* lo = reinterpret_cast<float4*>(lmbds + (size_t)e * 8);
        lo[0] = make_float4(sorted[0], sorted[1], sorted[2], sorted[3]);
        lo[1] = make_float4(sorted[4], sorted[5], sorted[6], sorted[7]);
        if (far_lmbds) {
          float4* fo = reinterpret_cast<float4*>(far_lmbds + (size_t)far_e * 8);
          fo[0] = lo[0];
          fo[1] = lo[1];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicMax(&s_cm[wib][c], __float_as_uint(sorted[c]));
      }
      // G[i][j] = (ker W)[i][j] / (S_j |A_f col i|^2) for the kept columns;  C_f[r][j] = sum_i A_f[r][i] G[i][j]
      float invf[8];
      inv_col_norms(F, invf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float is = sig[j] > pinv_eps ? 1.f / sig[j] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const p2 sc = x2::pk(invf[2 * k] * is, invf[2 * k + 1] * is);
          A.X[j][k] = x2::mul2(A.X[j][k], sc);
          A.Y[j][k] = x2::mul2(A.Y[j][k], sc);
        }
      }
      float2* cf = canon + (size_t)(e + L) * 64;
      float2* far_cf = far_canon ? far_canon + (size_t)(far_e + far_L) * 64 : nullptr;
#pragma unroll 2
      for (int r = 0; r < 8; ++r) {
        float2 f[8];
        lds_row(f, F + r * 64);
        p2 FX[4], FY[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { FX[k] = x2::pk(f[2 * k].x, f[2 * k + 1].x); FY[k] = x2::pk(f[2 * k].y, f[2 * k + 1].y); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          p2 re = x2::mul2(FX[0], A.X[j][0]);
          p2 im = x2::mul2(FX[0], A.Y[j][0]);
          re = x2::fnma2(FY[0], A.Y[j][0], re);
          im = x2::fma2(FY[0], A.X[j][0], im);
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            re = x2::fma2(FX[k], A.X[j][k], re);
            im = x2::fma2(FX[k], A.Y[j][k], im);
            re = x2::fnma2(FY[k], A.Y[j][k], re);
            im = x2::fma2(FY[k], A.X[j][k], im);
          }
          if (live && rk[j] < ncols) {
            const float2 v = make_float2(x2::hsum(re), x2::hsum(im));
            cf[r * 8 + rk[j]] = v;
            if (far_cf) far_cf[r * 8 + rk[j]] = v;
          }
        }
      }
    } else {
      // follower: column j holds conj(C_b)[:, j]; slot e, column position = the leader's rank of j
      float2* cb = canon + (size_t)e * 64;
      float2* far_cb = far_canon ? far_canon + (size_t)far_e * 64 : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (int)((packed >> (3 * j)) & 7u);
        const bool keep = (packed >> (24 + j)) & 1u;
        if (live && c < ncols) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 xr = x2::unpk(A.X[j][k]), yi = x2::unpk(A.Y[j][k]);
            const float2 v0 = keep ? make_float2(xr.x, -yi.x) : make_float2(0.f, 0.f);
            const float2 v1 = keep ? make_float2(xr.y, -yi.y) : make_float2(0.f, 0.f);
            cb[(2 * k) * 8 + c] = v0;
            cb[(2 * k + 1) * 8 + c] = v1;
            if (far_cb) { far_cb[(2 * k) * 8 + c] = v0; far_cb[(2 * k + 1) * 8 + c] = v1; }
          }
        }
      }
    }
    if (leader && live) atomicAdd(&s_cnt[wib][5], 1);
    __syncwarp();                                           // F is rewritten by the next iteration's phase 1
  }
  // column-wise max of lambda over all edges (truncate_lmbds, backends.py:297-299)
  __syncwarp();
  if (lane < 8) atomicMax(reinterpret_cast<unsigned int*>(colmax + lane), s_cm[wib][lane]);
  if (lane == 0) {
    atomicAdd(&g_stats[0], (unsigned long long)s_cnt[wib][2]);
    atomicAdd(&g_stats[1], (unsigned long long)s_cnt[wib][0]);
    atomicAdd(&g_stats[2], (unsigned long long)s_cnt[wib][1]);
    atomicAdd(&g_stats[3], (unsigned long long)s_cnt[wib][3]);
    atomicAdd(&g_stats[4], (unsigned long long)s_cnt[wib][4]);
    atomicAdd(&g_stats[5], 2ull * (unsigned long long)s_cnt[wib][5]);
    atomicAdd(&g_stats[6], (unsigned long long)s_cnt[wib][5]);
  }
  if (threadIdx.x == 0) {
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    atomicMin(&g_span[0], s_t0);
    atomicMax(&g_span[1], t_end);
  }
}

}  // namespace canon8v2

void canon8v2_stats(unsigned long long* out3) {
  cudaMemcpyFromSymbol(out3, canon8v2::g_stats, sizeof(unsigned long long) * 3);
}
void canon8v2_stats_detail(unsigned long long* out7) {
  cudaMemcpyFromSymbol(out7, canon8v2::g_stats, sizeof(unsigned long long) * 7);
}
// (start, end) of the launches since the last call, then reset
void canon8v2_span(unsigned long long* out2) {
  cudaMemcpyFromSymbol(out2, canon8v2::g_span, sizeof(unsigned long long) * 2);
  const unsigned long long init[2] = {~0ull, 0ull};
  cudaMemcpyToSymbol(canon8v2::g_span, init, sizeof(init));
}

// ---- counting sort of the edges by cost, descending (one block; keys are bytes; warp-aggregated shared-memory atomics)
__global__ void __launch_bounds__(1024) k_sort_by_cost(long long L, const unsigned char* __restrict__ cost, int* __restrict__ order) {
  __shared__ unsigned hist[256];
  const int lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  const long long rounds = (L + blockDim.x - 1) / blockDim.x;
  for (long long rnd = 0; rnd < rounds; ++rnd) {
    const long long i = rnd * blockDim.x + threadIdx.x;
    const unsigned key = i < L ? cost[i] : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[key], __popc(peers));
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                   // start of every key's run, largest key first
    unsigned run = 0;
    for (int k = 255; k >= 0; --k) { const unsigned c = hist[k]; hist[k] = run; run += c; }
  }
  __syncthreads();
  for (long long rnd = 0; rnd < rounds; ++rnd) {
    const long long i = rnd * blockDim.x + threadIdx.x;
    const unsigned key = i < L ? cost[i] : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (key < 256u && lane == leader) base = atomicAdd(&hist[key], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (key < 256u) order[base + __popc(peers & ((1u << lane) - 1u))] = (int)i;
  }
}

int launch_sort_edges_by_cost(long long L, const void* cost, int32_t* order, cudaStream_t st) {
  if (L <= 0) return 0;
  if (L >= (1LL << 31)) return set_error("sort_edges_by_cost: %lld edges exceed the 32-bit index range", L);
  k_sort_by_cost<<<1, 1024, 0, st>>>(L, (const unsigned char*)cost, order);
  return after_launch("sort_edges_by_cost");
}

int launch_fast_canon8v2(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                         int ncols, const int32_t* order, void* cost, cudaStream_t st, long long n_edges,
                         const int32_t* remote, void* const* peer_canon, void* const* peer_lmbds, const long long* peer_L) {
  using namespace canon8v2;
  const long long n = n_edges >= 0 ? n_edges : L;
  if (n == 0) return 0;
  if (n > L || (n < L && !order)) return set_error("canonicalize: %lld of %lld edges without an edge list", n, L);
  Remote rem{};
  rem.code = (remote && peer_canon && peer_lmbds && peer_L) ? remote : nullptr;
  for (int q = 0; q < BQA_MAX_PEERS; ++q) {
    rem.canon[q] = rem.code ? (float2*)peer_canon[q] : nullptr;
    rem.lmbds[q] = rem.code ? (float*)peer_lmbds[q] : nullptr;
    rem.L[q] = rem.code ? peer_L[q] : 0;
  }
  if (ncols < 1 || ncols > 8) return set_error("canonicalize: %d canonicalizer columns requested for n = 8", ncols);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_canon8v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_canon8v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(k_canon8v2): %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const long long groups = (n + kEdges - 1) / kEdges;
  long long grid = (groups + kWarps - 1) / kWarps;
  if (grid > sms) grid = sms;
  // n^2 of LAPACK xGESVJ's estimate "next sweep's largest cosine ~ n max|cos| max|sin|" (experiments: BQA_B200_CANON_CONV)
  static const float conv = [] { const char* e = getenv("BQA_B200_CANON_CONV"); return e ? (float)atof(e) : 64.f; }();
  if (rem.code)
    k_canon8v2<true><<<(int)grid, kWarps * 32, kSmem, st>>>(n, L, (const float2*)ext, (float2*)canon, (float*)lmbds,
                                                           (float*)colmax, (float)pinv_eps, ncols, 2, order,
                                                           (unsigned char*)cost, conv, rem);
  else
    k_canon8v2<false><<<(int)grid, kWarps * 32, kSmem, st>>>(n, L, (const float2*)ext, (float2*)canon, (float*)lmbds,
                                                            (float*)colmax, (float)pinv_eps, ncols, 2, order,
                                                            (unsigned char*)cost, conv, rem);
  return after_launch("canonicalize(n=8, v2)");
}

}  // namespace bqa
