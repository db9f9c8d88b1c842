// bqa_fast_canon8.cu -- canonicalizers of bond dimension 4 (extended dimension n = 8) in complex64.
//
// replaces _get_canonicalizers (src/bqa/state.py:171-200) for the headline shape: per undirected edge
//   m_f = V_f L_f V_f^H, m_b = V_b L_b V_b^H        masked eigendecompositions   (backends.py:483-490, 709-727)
//   ker = L_f^1/2 V_f^H conj(V_b) L_b^1/2            (state.py:186-187)
//   ker = U S W^H                                    masked SVD                   (state.py:189)
//   C_f = V_f L_f^-1/2 U (slot e + L),  C_b = V_b L_b^-1/2 conj(W) (slot e),  lambda = S / |S|   (state.py:196-200)
//
// Eight lanes own an edge: lane r holds ROW r of the working matrix and of the accumulated rotations, so a
// one-sided (Hestenes) Jacobi rotation of columns (p, q) is thread-local once the column inner product has
// been all-reduced over the 8 lanes with xor shuffles.  Rotations are scheduled round-robin (7 rounds of 4
// disjoint pairs per sweep): the four inner products of a round are reduced together, which gives every
// lane four independent dependency chains.  Column norms are carried along (alpha' = alpha - t |g|,
// beta' = beta + t |g|) and recomputed exactly once per sweep.  A warp holds 4 edges; matrices move between
// the "row per lane" and "column per lane" views through a padded shared-memory tile.
#include <cuda_runtime.h>

#include "bqa_core.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace canon8 {

constexpr int kWarps = 4;
constexpr int kRow = 80;                   // 64-byte row of 8 complex + 16 bytes of padding
constexpr int kMat = 8 * kRow;             // one 8 x 8 complex tile
constexpr int kEdgeBytes = 2 * kMat + 128; // two tiles + eigenvalues of m_f, m_b (16 floats) + sorted sigma (8) + permutation (8 ints)
constexpr int kWarpBytes = 4 * kEdgeBytes;

// statistics: [0] Jacobi problems solved (per warp: 4 matrices at a time), [1] sweeps summed over them,
// [2] of which spent on the SVD of ker (the other two problems per edge are the message eigendecompositions)
__device__ unsigned long long g_stats[3];

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float red8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// squared column norms of the 8 x 8 matrix whose row lives in this lane
__device__ __forceinline__ void col_norms(const float2 (&A)[8], float (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = red8(A[j].x * A[j].x + A[j].y * A[j].y);
}

// rsqrt with one Newton step (MUFU.RSQ is good to ~2 ulp; rotations must stay orthonormal to rounding)
__device__ __forceinline__ float rsqrt_nr(float x) {
  const float y = rsqrtf(x);
  return y * (1.5f - 0.5f * x * y * y);
}

// Rotation parameters of one column pair, computed from the all-reduced inner product (gr, gi) and the column norms.
// Branch-free: an inactive pair (converged, or a numerically-zero column) gets the identity through selects.
struct Rot {
  float c, s, phx, phy, dw;      // cos, sin, unimodular phase conj(g)/|g|, norm transfer t |g|
};
// `rotated` is raised when the pair needed a rotation.  The sweep loop runs until a whole sweep needs none, for every
// edge of the warp: extra sweeps on an already converged edge are exact identities, so an edge's result does not
// depend on which other edges share its warp (single-GPU and partitioned runs stay bit-identical).
__device__ __forceinline__ Rot rot_params(float al, float be, float gr, float gi, float nul, float tol2, bool& rotated) {
  const float g2 = gr * gr + gi * gi;
  const bool act = !(al <= nul || be <= nul || g2 <= tol2 * al * be);
  rotated |= act;
  const float ig = rsqrt_nr(g2);                  // 1 / |g|   (inf / nan when inactive: discarded below)
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
  return r;
}

template <int P, int Q>
__device__ __forceinline__ void apply_rot(float2 (&A)[8], float2 (&V)[8], float (&w)[8], const Rot& r) {
  const float2 ph = make_float2(r.phx, r.phy);
  float2 ap = A[P], aq = cmulf(ph, A[Q]);
  A[P] = make_float2(r.c * ap.x - r.s * aq.x, r.c * ap.y - r.s * aq.y);
  A[Q] = make_float2(r.s * ap.x + r.c * aq.x, r.s * ap.y + r.c * aq.y);
  ap = V[P]; aq = cmulf(ph, V[Q]);
  V[P] = make_float2(r.c * ap.x - r.s * aq.x, r.c * ap.y - r.s * aq.y);
  V[Q] = make_float2(r.s * ap.x + r.c * aq.x, r.s * ap.y + r.c * aq.y);
  w[P] -= r.dw;
  w[Q] += r.dw;
}

// One round = four disjoint pairs (P0,Q0) .. (P3,Q3).  The 8 partial inner-product components (re, im of 4 pairs)
// are reduce-scattered over the 8 lanes of the edge (7 shuffles; lane r ends up with component r), the two lanes
// of pair k = r / 2 swap components and compute THAT pair's rotation only, and the four parameter sets are then
// broadcast (5 shuffles each) -- instead of every lane all-reducing 8 values and deriving all four rotations.
template <int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
__device__ __forceinline__ void jacobi_round(float2 (&A)[8], float2 (&V)[8], float (&w)[8], float nul, float tol2,
                                             bool& rotated, int r) {
  float g[8];
  g[0] = A[P0].x * A[Q0].x + A[P0].y * A[Q0].y; g[1] = A[P0].x * A[Q0].y - A[P0].y * A[Q0].x;
  g[2] = A[P1].x * A[Q1].x + A[P1].y * A[Q1].y; g[3] = A[P1].x * A[Q1].y - A[P1].y * A[Q1].x;
  g[4] = A[P2].x * A[Q2].x + A[P2].y * A[Q2].y; g[5] = A[P2].x * A[Q2].y - A[P2].y * A[Q2].x;
  g[6] = A[P3].x * A[Q3].x + A[P3].y * A[Q3].y; g[7] = A[P3].x * A[Q3].y - A[P3].y * A[Q3].x;
  const bool b2 = r & 4, b1 = r & 2, b0 = r & 1;
  float h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b2 ? g[i] : g[4 + i], keep = b2 ? g[4 + i] : g[i];
    h[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float k2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b1 ? h[i] : h[2 + i], keep = b1 ? h[2 + i] : h[i];
    k2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float mine = (b0 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? k2[0] : k2[1], 1);   // component r
  const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
  const float gr = b0 ? other : mine, gi = b0 ? mine : other;
  // column norms of this lane's pair k = r / 2
  const float al = b2 ? (b1 ? w[P3] : w[P2]) : (b1 ? w[P1] : w[P0]);
  const float be = b2 ? (b1 ? w[Q3] : w[Q2]) : (b1 ? w[Q1] : w[Q0]);
  const Rot mineR = rot_params(al, be, gr, gi, nul, tol2, rotated);
  const int base = (threadIdx.x & 31) & ~7;
  Rot R[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    R[k].c = __shfl_sync(0xffffffffu, mineR.c, base + 2 * k);
    R[k].s = __shfl_sync(0xffffffffu, mineR.s, base + 2 * k);
    R[k].phx = __shfl_sync(0xffffffffu, mineR.phx, base + 2 * k);
    R[k].phy = __shfl_sync(0xffffffffu, mineR.phy, base + 2 * k);
    R[k].dw = __shfl_sync(0xffffffffu, mineR.dw, base + 2 * k);
  }
  apply_rot<P0, Q0>(A, V, w, R[0]);
  apply_rot<P1, Q1>(A, V, w, R[1]);
  apply_rot<P2, Q2>(A, V, w, R[2]);
  apply_rot<P3, Q3>(A, V, w, R[3]);
}

// one-sided Jacobi SVD: on exit A = U diag(sigma) (row of this lane), V = right singular vectors (row of this
// lane), w = sigma^2 per column (all lanes).  Code size matters here (the instruction cache is 32 KB and the
// warps of an SM sit at different points of the kernel): a sweep is ONE round body executed 7 times, with the
// columns 1..7 rotated through the registers between rounds (circle method: pairs (0,7) (1,6) (2,5) (3,4) by
// position); after 7 rounds every pair has met once and the columns are back in place.
__device__ __forceinline__ void jacobi8(float2 (&A)[8], float2 (&V)[8], float (&w)[8], int r, int& sweeps) {
#pragma unroll
  for (int j = 0; j < 8; ++j) V[j] = make_float2(j == r ? 1.f : 0.f, 0.f);
  const float eps = 1.1920929e-07f;
  const float tol = eps * 2.f * 2.8284271f;                 // eps * 2 * sqrt(n), like the generic kernel
  const float tol2 = tol * tol;
  col_norms(A, w);
  const float fro2 = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
  const float nul = eps * eps * fro2;                       // columns below eps |A|_F are numerically zero
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll 1
    for (int round = 0; round < 7; ++round) {
      jacobi_round<0, 7, 1, 6, 2, 5, 3, 4>(A, V, w, nul, tol2, rotated, r);
      const float2 a7 = A[7], v7 = V[7];
      const float w7 = w[7];
#pragma unroll
      for (int j = 7; j > 1; --j) { A[j] = A[j - 1]; V[j] = V[j - 1]; w[j] = w[j - 1]; }
      A[1] = a7; V[1] = v7; w[1] = w7;
    }
    col_norms(A, w);                                        // exact norms once per sweep
    ++sweeps;
    if (!__any_sync(0xffffffffu, rotated)) break;
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
__device__ __forceinline__ void sts_row(unsigned char* dst, const float2 (&A)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + 16 * i) = make_float4(A[2 * i].x, A[2 * i].y, A[2 * i + 1].x, A[2 * i + 1].y);
}

__global__ void __launch_bounds__(kWarps * 32) k_canon8(long long L, const float2* __restrict__ ext,
                                                        float2* __restrict__ canon, float* __restrict__ lmbds,
                                                        float* __restrict__ colmax, float pinv_eps, int ncols) {
  __shared__ __align__(16) unsigned char smem[kWarps * kWarpBytes];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int eg = lane >> 3, r = lane & 7;                   // edge slot in the warp, row owned by this lane
  unsigned char* tile0 = smem + wib * kWarpBytes + eg * kEdgeBytes;
  unsigned char* tile1 = tile0 + kMat;
  float* seig = reinterpret_cast<float*>(tile0 + 2 * kMat);            // eigenvalues of m_f (8) and m_b (8)
  float* ssig = seig + 16;
  int* scol = reinterpret_cast<int*>(ssig + 8);
  const long long groups = (L + 3) >> 2;
  const long long nwarps = (long long)gridDim.x * kWarps;
  float cm = 0.f;                                           // running max of lambda[:, r] over this lane's edges
  int n_sweeps = 0, n_jac = 0, n_ker = 0;
  for (long long g = (long long)blockIdx.x * kWarps + wib; g < groups; g += nwarps) {
    long long e = g * 4 + eg;
    const bool live = e < L;
    e = live ? e : L - 1;
    float2 A[8], W[8];
    float sk[8];
    // m = 0: eigenvectors of m_f -> tile0, m = 1: eigenvectors of m_b -> tile1, m = 2: SVD of ker
#pragma unroll 1
    for (int m = 0; m < 3; ++m) {
      if (m < 2) {
        load_row(A, ext + (size_t)(e + (m ? L : 0)) * 64 + r * 8);
      } else {
        // ker[i][j] = sqrt(sf_i sb_j) sum_k conj(Vf[k][i]) conj(Vb[k][j]) with masked eigenvalues: lane i takes
        // column i of Vf and whole rows of Vb from the tiles
        const float ef = seig[r];
        const float rsf = ef > pinv_eps ? sqrtf(ef) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) A[j] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int k = 0; k < 8; ++k) {
          const float2 vf = *reinterpret_cast<const float2*>(tile0 + k * kRow + r * 8);
          float2 row[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(tile1 + k * kRow + 16 * i);
            row[2 * i] = make_float2(v.x, v.y);
            row[2 * i + 1] = make_float2(v.z, v.w);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {                     // conj(vf) conj(vb) = conj(vf vb)
            A[j].x += vf.x * row[j].x - vf.y * row[j].y;
            A[j].y -= vf.x * row[j].y + vf.y * row[j].x;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float eb = seig[8 + j];
          const float sc = rsf * (eb > pinv_eps ? sqrtf(eb) : 0.f);
          A[j].x *= sc; A[j].y *= sc;
        }
      }
      const int before = n_sweeps;
      jacobi8(A, W, sk, r, n_sweeps);
      if (m == 2) n_ker += n_sweeps - before;
      if (m < 2) {
        float mine = sk[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) mine = (r == j) ? sk[j] : mine;
        unsigned char* tile = m ? tile1 : tile0;
        __syncwarp();
        sts_row(tile + r * kRow, W);
        seig[m * 8 + r] = sqrtf(mine);                      // eigenvalue = singular value of the PSD message
        __syncwarp();
      }
    }
    n_jac += 3;
    // this lane's rows of V_f L_f^-1/2 and V_b L_b^-1/2 (masked like pinv_raw, backends.py:719-727)
    float2 Vf[8], Vb[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(tile0 + r * kRow + 16 * i);
      const float4 b = *reinterpret_cast<const float4*>(tile1 + r * kRow + 16 * i);
      Vf[2 * i] = make_float2(a.x, a.y); Vf[2 * i + 1] = make_float2(a.z, a.w);
      Vb[2 * i] = make_float2(b.x, b.y); Vb[2 * i + 1] = make_float2(b.z, b.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ef = seig[j], eb = seig[8 + j];
      const float isf = (ef > pinv_eps && sqrtf(ef) > 1.1920929e-07f) ? rsqrtf(ef) : 0.f;
      const float isb = (eb > pinv_eps && sqrtf(eb) > 1.1920929e-07f) ? rsqrtf(eb) : 0.f;
      Vf[j].x *= isf; Vf[j].y *= isf;
      Vb[j].x *= isb; Vb[j].y *= isb;
    }
    // sort the singular values (descending, stable) and publish A = U S and W through the tiles
    float mine = sk[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) mine = (r == j) ? sk[j] : mine;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) rank += (sk[j] > mine || (sk[j] == mine && j < r)) ? 1 : 0;
    __syncwarp();
    ssig[rank] = sqrtf(mine);
    scol[rank] = r;
    sts_row(tile0 + r * kRow, A);
    sts_row(tile1 + r * kRow, W);
    __syncwarp();
    // lambda = masked S / |masked S|
    float nrm2 = 0.f, my_s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = ssig[j];
      const float sm = s > pinv_eps ? s : 0.f;
      nrm2 += sm * sm;
      my_s = (r == j) ? sm : my_s;
    }
    const float lam = my_s / sqrtf(nrm2);
    if (live) {
      lmbds[(size_t)e * 8 + r] = lam;
      cm = fmaxf(cm, lam);
    }
    // C_f[r][c] = sum_i Vf[r][i] isf_i U[i][col_c],  U[i][col] = A[i][col] / S_col   (slot e + L)
    // C_b[r][c] = sum_j Vb[r][j] isb_j conj(W[j][col_c])                               (slot e)
    for (int c = 0; c < ncols; c += 2) {
      float2 cf[2], cb[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = c + h < 8 ? c + h : 7;
        const int col = scol[cc];
        const float s = ssig[cc];
        float2 f = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        if (s > pinv_eps) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 u = *reinterpret_cast<const float2*>(tile0 + i * kRow + col * 8);
            const float2 w = *reinterpret_cast<const float2*>(tile1 + i * kRow + col * 8);
            f.x += Vf[i].x * u.x - Vf[i].y * u.y; f.y += Vf[i].x * u.y + Vf[i].y * u.x;
            b.x += Vb[i].x * w.x + Vb[i].y * w.y; b.y += Vb[i].y * w.x - Vb[i].x * w.y;     // times conj(w)
          }
          const float is = 1.f / s;
          f.x *= is; f.y *= is;
        }
        cf[h] = f; cb[h] = b;
      }
      if (live) {
        *reinterpret_cast<float4*>(canon + (size_t)(e + L) * 64 + r * 8 + c) = make_float4(cf[0].x, cf[0].y, cf[1].x, cf[1].y);
        *reinterpret_cast<float4*>(canon + (size_t)e * 64 + r * 8 + c) = make_float4(cb[0].x, cb[0].y, cb[1].x, cb[1].y);
      }
    }
    __syncwarp();
  }
  // column-wise max of lambda over all edges (truncate_lmbds, backends.py:297-299)
  cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
  cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
  if (lane < 8) atomicMax(reinterpret_cast<unsigned int*>(colmax + lane), __float_as_uint(cm));
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
  const long long groups = (L + 3) / 4;
  long long grid = (groups + kWarps - 1) / kWarps;
  const long long cap = 148LL * 16;
  if (grid > cap) grid = cap;
  k_canon8<<<(int)grid, kWarps * 32, 0, st>>>(L, (const float2*)ext, (float2*)canon, (float*)lmbds, (float*)colmax,
                                              (float)pinv_eps, ncols);
  return after_launch("canonicalize(n=8)");
}

}  // namespace bqa
