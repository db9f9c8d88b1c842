// bqa_generic.cuh -- shape-generic kernels (any degree <= 8, any bond dimension <= 16, c64 / c128).
// One warp per node (or per undirected edge); scratch for the contracted tensors lives in a
// per-warp slab of a global workspace (L1/L2 resident), the small Jacobi matrices in shared memory.
// The headline shapes are served by the specialised kernels in bqa_fast_*.cu; these kernels are the
// functional baseline they are validated against and the path for every other (degree, D).
#pragma once
#include <cuda_runtime.h>

#include "bqa_core.cuh"
#include "bqa_launch.cuh"

namespace bqa {

// ---- atomic max on non-negative reals through their bit patterns -------------------------------
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
template <typename R> __device__ __forceinline__ R warp_max(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// convergence test shared by every sweep kernel: true => this launch must do nothing
template <typename R>
__device__ __forceinline__ bool bp_already_converged(int it, const R* resid, int32_t* status, R bp_eps) {
  if (it == 0) return false;
  if (*((volatile int32_t*)status) != 0) return true;
  const R num = resid[2 * (it - 1)], den = resid[2 * (it - 1) + 1];
  if (msqrt(num / den) < bp_eps) {
    if (threadIdx.x == 0) { status[1] = it; __threadfence(); status[0] = 1; }
    return true;
  }
  return false;
}

template <typename R>
struct NodeArgs {
  int d, D, Dn;
  long long B;
  const cx<R>* T;
  cx<R>* Tout;
  const cx<R>* msgs_cur;
  cx<R>* msgs_out;          // msgs_nxt / ext / re-initialised messages
  const int32_t* in_pos;
  const int32_t* out_pos;
  const int32_t* lmbd_pos;
  const int32_t* node_ids;
  const R* node_ampls;
  const R* edge_ampls;
  const cx<R>* canon;
  const R* lmbds;
  R* bloch;
  R damping, bp_eps, ztime, xtime;
  int write_undamped, it;
  R* resid;
  int32_t* status;
  cx<R>* ws;
  size_t ws_per_warp;       // elements
  // multi-GPU over peer memory: remote_pos (d, B), -1 or (peer << 27 | slot in that peer's array); peers[q] = base
  // of rank q's destination array mapped into this process
  const int32_t* remote_pos;
  unsigned char* peers[BQA_MAX_PEERS];
};

// ---- BP sweep / extended messages ---------------------------------------------------------------
template <typename R, bool EXT>
__global__ void __launch_bounds__(128) k_node_msgs(NodeArgs<R> a) {
  if (!EXT && bp_already_converged<R>(a.it, a.resid, a.status, a.bp_eps)) return;
  GroupWarp g;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int d = a.d, D = a.D, DD = D * D;
  const int W = 2 * ipow(D, d);
  cx<R>* P = a.ws + (size_t)warp * a.ws_per_warp;
  cx<R>* E = P + W;
  cx<R>* gram = E + W;
  R mnum = 0, mden = 0;
  for (long long node = warp; node < a.B; node += nwarps) {
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = a.msgs_cur + (size_t)a.in_pos[(size_t)j * a.B + node] * DD;
    node_gram<R>(g, d, D, a.T + (size_t)node * W, mp, P, E, gram);
    for (int k = 0; k < d; ++k) {
      const cx<R>* g0 = gram + (size_t)k * 2 * DD;
      const cx<R>* g1 = g0 + DD;
      const size_t slot = (size_t)a.out_pos[(size_t)k * a.B + node];
      const int rp = a.remote_pos ? a.remote_pos[(size_t)k * a.B + node] : -1;
      cx<R>* far = rp < 0 ? nullptr
                          : reinterpret_cast<cx<R>*>(a.peers[rp >> 27]) + (size_t)(rp & ((1 << 27) - 1)) * (EXT ? 4 * DD : DD);
      if (!EXT)
        emit_bp_msg<R>(g, D, g0, g1, a.msgs_cur + slot * DD, a.msgs_out + slot * DD, a.damping, a.write_undamped,
                       mnum, mden, far);
      else
        emit_ext_msg<R>(g, D, g0, g1, a.edge_ampls[(size_t)k * a.B + node] * a.ztime, a.msgs_out + slot * 4 * DD, far);
    }
    g.sync();
  }
  if (!EXT) {
    mnum = warp_max(mnum);
    mden = warp_max(mden);
    if (lane == 0) {
      atomic_max_nonneg(a.resid + 2 * a.it, mnum);
      atomic_max_nonneg(a.resid + 2 * a.it + 1, mden);
    }
  }
}

// ---- canonicalizers ------------------------------------------------------------------------------
// One-sided Jacobi SVD of an N x N complex matrix (N = 16, 32) by one warp with ROUND-ROBIN ordering: a sweep is N - 1
// rounds of N / 2 disjoint column pairs, and all pairs of a round are handled together --
//   1. every lane forms the four partial sums (|a_p|^2, |a_q|^2, <a_p, a_q>) of each pair from its row,
//   2. one butterfly reduce-scatter over the lanes leaves the four totals of pair i on the lanes with
//      (lane mod N/2) == i  (36 shuffles for N = 16, 64 for N = 32, against 20 per pair one after the other),
//   3. those lanes derive the rotation of their pair concurrently and publish it through shared memory,
//   4. every lane applies the N / 2 rotations to its row of A and of V (N = 16: lanes 16..31 take the rows of V).
// Same rotation formulas, thresholds and stopping rule as the serial cyclic routine jacobi_svd (bqa_core.cuh), which
// handles one pair at a time with a warp-wide reduction and a dependent parameter chain each (r2 capture of the n = 16
// kernel: issue slots 34 % used, the warp waits on its own shuffle / square-root latencies).  The order of the rotations
// differs, so results agree with the serial routine to rounding, not bit for bit.
// one step of the butterfly reduce-scatter: the lanes whose bit MASK is set keep the upper half of the CNT values
template <typename R, int NV, int CNT, int MASK>
__device__ __forceinline__ void reduce_scatter_step(R (&v)[NV], int lane) {
  if constexpr (MASK >= 1) {
    const bool up = (lane & MASK) != 0;
#pragma unroll
    for (int j = 0; j < CNT / 2; ++j) {
      const R send = up ? v[j] : v[j + CNT / 2];
      const R keep = up ? v[j + CNT / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
    reduce_scatter_step<R, NV, CNT / 2, MASK / 2>(v, lane);
  }
}

template <typename R, int N>
__device__ void jacobi_svd_round_robin(int ld, cx<R>* A, cx<R>* V, R* sigma, int* order, R* prm) {
  constexpr int H = N / 2, M1 = N - 1, NV = 4 * H;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < N * N; i += 32) V[(i / N) * ld + i % N] = mk<R>((i / N == i % N) ? R(1) : R(0), R(0));
  const R tol = num_traits<R>::eps() * R(2) * msqrt((R)N);
  R fro2 = 0;
  if (lane < N)
    for (int j = 0; j < N; ++j) fro2 += norm2(A[lane * ld + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fro2 += __shfl_xor_sync(full, fro2, o);
  const R nul = num_traits<R>::eps() * num_traits<R>::eps() * fro2;
  __syncwarp();
  // pair i of round t (circle method: column N - 1 stays, the others rotate)
  auto pair_of = [&](int t, int i, int& p, int& q) {
    int a = t + i, b = t + M1 - i;
    if (a >= M1) a -= M1;
    if (b >= M1) b -= M1;
    if (i == 0) a = M1;
    p = a < b ? a : b;
    q = a < b ? b : a;
  };
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    for (int t = 0; t < M1; ++t) {
      R v[NV];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        int p, q;
        pair_of(t, i, p, q);
        cx<R> ap = mk<R>(0, 0), aq = ap;
        if (lane < N) { ap = A[lane * ld + p]; aq = A[lane * ld + q]; }
        v[4 * i] = norm2(ap);
        v[4 * i + 1] = norm2(aq);
        v[4 * i + 2] = ap.re * aq.re + ap.im * aq.im;        // conj(ap) * aq
        v[4 * i + 3] = ap.re * aq.im - ap.im * aq.re;
      }
      // reduce-scatter over the pair index (lane bits log2(H) - 1 .. 0), all-reduce over the remaining lane bits
      reduce_scatter_step<R, NV, NV, H / 2>(v, lane);
#pragma unroll
      for (int mask = H; mask < 32; mask *= 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(full, v[j], mask);
      }
      const R al = v[0], be = v[1], gr = v[2], gi = v[3];
      const R g2 = gr * gr + gi * gi;
      const bool rot = !(al <= nul || be <= nul || g2 <= tol * tol * al * be);
      R c = R(1), sn = R(0), phr = R(1), phi = R(0);
      if (rot) {
        const R ag = msqrt(g2);
        const R zeta = (be - al) / (R(2) * ag);
        const R tt = ((zeta >= R(0)) ? R(1) : R(-1)) / (mabs(zeta) + msqrt(R(1) + zeta * zeta));
        c = R(1) / msqrt(R(1) + tt * tt);
        sn = c * tt;
        phr = gr / ag;                                   // e^{-i arg(gamma)}
        phi = -gi / ag;
      }
      if (lane < H) { prm[4 * lane] = c; prm[4 * lane + 1] = sn; prm[4 * lane + 2] = phr; prm[4 * lane + 3] = phi; }
      rotated = rotated || __any_sync(full, rot);
      __syncwarp();
      // rows: N = 16: lanes 0..15 own the rows of A, lanes 16..31 the rows of V; N = 32: every lane one row of each
#pragma unroll
      for (int pass = 0; pass < (N == 32 ? 2 : 1); ++pass) {
        cx<R>* row = (N == 32 ? (pass == 0 ? A : V) : (lane < 16 ? A : V)) + (lane & (N - 1)) * ld;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          int p, q;
          pair_of(t, i, p, q);
          const R ci = prm[4 * i], si = prm[4 * i + 1];
          const cx<R> ph = mk<R>(prm[4 * i + 2], prm[4 * i + 3]);
          if (si != R(0)) {                              // identity: the pair was skipped
            const cx<R> ap = row[p], aq = ph * row[q];
            row[p] = ci * ap - si * aq;
            row[q] = si * ap + ci * aq;
          }
        }
      }
      __syncwarp();
    }
    if (!rotated) break;
  }
  for (int j = 0; j < N; ++j) {                          // singular values = column norms (as in jacobi_svd)
    R a = 0;
    if (lane < N) a = norm2(A[lane * ld + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(full, a, o);
    if (lane == 0) sigma[j] = msqrt(a);
  }
  __syncwarp();
  if (lane == 0) {
    for (int j = 0; j < N; ++j) order[j] = j;
    for (int i = 1; i < N; ++i) {                        // stable insertion sort, descending
      const int oi = order[i];
      int j = i - 1;
      while (j >= 0 && sigma[order[j]] < sigma[oi]) { order[j + 1] = order[j]; --j; }
      order[j + 1] = oi;
    }
  }
  __syncwarp();
}

struct RoundRobinJacobi {
  template <typename R, typename G>
  static __device__ void run(G g, int n, int ld, cx<R>* A, cx<R>* V, R* sigma, int* order, R* prm) {
    if (n == 16) jacobi_svd_round_robin<R, 16>(ld, A, V, sigma, order, prm);
    else if (n == 32) jacobi_svd_round_robin<R, 32>(ld, A, V, sigma, order, prm);
    else jacobi_svd<R>(g, n, ld, A, V, sigma, order);
  }
};

template <typename R>
__host__ __device__ inline size_t canon_warp_bytes(int n) {        // matrices | 3 n reals | 3 n ints | 2 n reals (rotations)
  const size_t b = edge_scratch_elems<R>(n) * sizeof(cx<R>) + 3 * n * sizeof(R) + 3 * n * sizeof(int) + 2 * n * sizeof(R);
  return (b + 15) / 16 * 16;
}

// MINB = 2 (n = 16 in complex64): the round-robin routine took 251 registers, one CTA of 8 warps per SM; capped at 128 two
// CTAs fit (the n = 32 branch, dead at that size, is what spills)
template <typename R, bool PAR, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) k_canonicalize(int D, long long L, const cx<R>* ext, cx<R>* canon,
                                                      R* lmbds, R* colmax, R pinv_eps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GroupWarp g;
  const int n = 2 * D, nn = n * n;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem_raw + (size_t)wib * canon_warp_bytes<R>(n);
  cx<R>* scratch = reinterpret_cast<cx<R>*>(base);
  R* rs = reinterpret_cast<R*>(scratch + edge_scratch_elems<R>(n));
  int* is = reinterpret_cast<int*>(rs + 3 * n);
  R* prm = reinterpret_cast<R*>(is + 3 * n);
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  R cm = 0;
  for (long long e = warp; e < L; e += nwarps) {
    if constexpr (PAR)
      edge_canonicalize<R, GroupWarp, RoundRobinJacobi>(g, n, ext + (size_t)e * nn, ext + (size_t)(e + L) * nn, pinv_eps,
                                                        scratch, rs, is, canon + (size_t)e * nn,
                                                        canon + (size_t)(e + L) * nn, lmbds + (size_t)e * n, prm);
    else
      edge_canonicalize<R>(g, n, ext + (size_t)e * nn, ext + (size_t)(e + L) * nn, pinv_eps, scratch, rs, is,
                           canon + (size_t)e * nn, canon + (size_t)(e + L) * nn, lmbds + (size_t)e * n);
    g.sync();
    if (lane < n) cm = max(cm, lmbds[(size_t)e * n + lane]);
  }
  if (lane < n) atomic_max_nonneg(colmax + lane, cm);
}

// ---- simple update application + Rz/Rx + symmetric gauge ---------------------------------------
template <typename R>
__global__ void __launch_bounds__(128) k_apply_update(NodeArgs<R> a) {
  GroupWarp g;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int d = a.d, D = a.D, Dn = a.Dn, n = 2 * D;
  const int Win = 2 * ipow(D, d), Wout = 2 * ipow(Dn, d);
  const int Wmax = 2 * ipow(D > Dn ? D : Dn, d);
  cx<R>* bufA = a.ws + (size_t)warp * a.ws_per_warp;
  cx<R>* bufB = bufA + Wmax;
  cx<R>* wbuf = bufB + Wmax;
  for (long long node = warp; node < a.B; node += nwarps) {
    const cx<R>* cp[BQA_MAX_DEGREE];
    const R* lp[BQA_MAX_DEGREE];
    R th[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) {
      cp[j] = a.canon + (size_t)a.in_pos[(size_t)j * a.B + node] * n * n;
      lp[j] = a.lmbds + (size_t)a.lmbd_pos[(size_t)j * a.B + node] * n;
      th[j] = a.edge_ampls[(size_t)j * a.B + node] * a.ztime;
    }
    node_apply_update<R>(g, d, D, Dn, a.T + (size_t)node * Win, cp, th, lp, a.node_ampls[node] * a.ztime,
                         a.xtime, bufA, bufB, wbuf, a.Tout + (size_t)node * Wout);
    // messages of the symmetric gauge: diag(lambda) / trace at this node's outgoing slots
    for (int j = 0; j < d; ++j)
      emit_gauge_msg<R>(g, Dn, lp[j], a.msgs_out + (size_t)a.out_pos[(size_t)j * a.B + node] * Dn * Dn);
  }
}

// ---- marginals -------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(128) k_density(NodeArgs<R> a) {
  GroupWarp g;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int d = a.d, D = a.D, DD = D * D;
  const int W = 2 * ipow(D, d);
  cx<R>* E = a.ws + (size_t)warp * a.ws_per_warp;
  for (long long node = warp; node < a.B; node += nwarps) {
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = a.msgs_cur + (size_t)a.in_pos[(size_t)j * a.B + node] * DD;
    node_density<R>(g, d, D, a.T + (size_t)node * W, mp, E, a.bloch + (size_t)a.node_ids[node] * 4);
  }
}

// ---- sampling helpers ------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(1024) k_argmax_unmeasured(long long N, const R* bloch, const int32_t* outcomes,
                                                            int32_t* result, R* result_p0) {
  __shared__ R skey[32];
  __shared__ long long sidx[32];
  __shared__ int scnt[32];
  R best = R(-1);
  long long bidx = -1;
  int cnt = 0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    if (outcomes[i] != 0) continue;
    ++cnt;
    const R key = fabs(R(2) * bloch[4 * i + 3] - R(1));
    if (key > best) { best = key; bidx = i; }           // strided ascending scan keeps the smallest index on ties
  }
  auto better = [](R k1, long long i1, R k2, long long i2) {
    return (i2 < 0) ? true : ((i1 < 0) ? false : (k1 > k2 || (k1 == k2 && i1 < i2)));
  };
  for (int o = 16; o > 0; o >>= 1) {
    const R ok = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (!better(best, bidx, ok, oi)) { best = ok; bidx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { skey[threadIdx.x >> 5] = best; sidx[threadIdx.x >> 5] = bidx; scnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? skey[threadIdx.x] : R(-1);
    bidx = threadIdx.x < nw ? sidx[threadIdx.x] : -1;
    cnt = threadIdx.x < nw ? scnt[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) {
      const R ok = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (!better(best, bidx, ok, oi)) { best = ok; bidx = oi; }
    }
    if (threadIdx.x == 0) {
      result[0] = (int32_t)bidx;
      result[1] = cnt;
      result_p0[0] = bidx >= 0 ? bloch[4 * bidx + 3] : R(0);
    }
  }
}

template <typename R>
__device__ void project_node_warp(cx<R>* t, int half, int bit) {
  node_project<R>(GroupWarp(), t, half, bit);
}

template <typename R>
__global__ void k_project_node(int half, cx<R>* T, long long pos, int bit) {
  project_node_warp<R>(T + (size_t)pos * 2 * half, half, bit);
}

template <typename R>
__global__ void __launch_bounds__(128) k_threshold_project(int half, long long B, cx<R>* T, const int32_t* node_ids,
                                                           const R* bloch, int32_t* outcomes, R thr, int32_t* n_proj) {
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long node = warp; node < B; node += nwarps) {
    const int32_t id = node_ids[node];
    if (outcomes[id] != 0) continue;
    const R p0 = bloch[4 * (size_t)id + 3];
    int bit = -1;
    if (p0 > thr) bit = 0;
    else if (p0 < R(1) - thr) bit = 1;
    if (bit < 0) continue;
    project_node_warp<R>(T + (size_t)node * 2 * half, half, bit);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { outcomes[id] = 1 - 2 * bit; atomicAdd(n_proj, 1); }
  }
}

// ---- launch helpers --------------------------------------------------------------------------------
static inline int node_grid(long long B, int warps_per_block) {
  long long blocks = (B + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)BQA_GENERIC_MAX_WARPS / warps_per_block;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename R>
static int check_ws(int d, int D, int Dn, size_t ws_bytes) {
  const size_t need = generic_ws_elems_per_warp(d, D, Dn) * sizeof(cx<R>) * BQA_GENERIC_MAX_WARPS;
  if (ws_bytes < need) return set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  return 0;
}

template <typename R>
int launch_node_msgs(bool ext, int d, int D, long long B, const void* T, const void* msgs_cur, void* msgs_out,
                     const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls, double ztime,
                     double damping, int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                     void* ws, size_t ws_bytes, const int32_t* remote_pos, void* const* peers, cudaStream_t st) {
  if (B == 0) return 0;
  if (int rc = check_ws<R>(d, D, D, ws_bytes)) return rc;
  NodeArgs<R> a{};
  a.d = d; a.D = D; a.Dn = D; a.B = B;
  a.T = (const cx<R>*)T; a.msgs_cur = (const cx<R>*)msgs_cur; a.msgs_out = (cx<R>*)msgs_out;
  a.in_pos = in_pos; a.out_pos = out_pos; a.edge_ampls = (const R*)edge_ampls;
  a.ztime = (R)ztime; a.damping = (R)damping; a.write_undamped = write_undamped; a.bp_eps = (R)bp_eps;
  a.it = it; a.resid = (R*)resid; a.status = status;
  a.ws = (cx<R>*)ws; a.ws_per_warp = generic_ws_elems_per_warp(d, D, D);
  a.remote_pos = peers ? remote_pos : nullptr;
  for (int q = 0; q < BQA_MAX_PEERS; ++q) a.peers[q] = peers ? (unsigned char*)peers[q] : nullptr;
  const int grid = node_grid(B, 4);
  if (ext) k_node_msgs<R, true><<<grid, 128, 0, st>>>(a);
  else k_node_msgs<R, false><<<grid, 128, 0, st>>>(a);
  return after_launch(ext ? "ext_msgs(generic)" : "bp_sweep(generic)");
}

template <typename R>
int launch_canonicalize(int D, long long L, const void* ext, void* canon, void* lmbds, void* colmax,
                        double pinv_eps, cudaStream_t st, bool round_robin) {
  if (L == 0) return 0;
  const int n = 2 * D;
  const size_t per_warp = canon_warp_bytes<R>(n);
  int wpb = (int)((size_t)200 * 1024 / per_warp);
  if (wpb > 8) wpb = 8;
  if (wpb < 1) return set_error("canonicalize: bond dimension %d needs %zu bytes of shared memory per edge", D, per_warp);
  const size_t smem = per_warp * wpb;
  // n = 16, 32: all column pairs of a round together (jacobi_svd_round_robin); smaller n: the serial routine
  const bool par = round_robin && (n == 16 || n == 32);
  const void* fn = par ? (const void*)k_canonicalize<R, true> : (const void*)k_canonicalize<R, false>;
  if (par && n == 16 && sizeof(R) == 4) fn = (const void*)k_canonicalize<R, true, 2>;
  if (smem > 48 * 1024) {                      // per device and per size: set on every call (microseconds per step)
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(canonicalize): %s", cudaGetErrorString(e));
  }
  long long blocks = (L + wpb - 1) / wpb;
  const long long cap = (long long)148 * 8;
  if (blocks > cap) blocks = cap;
  const cx<R>* e_ = (const cx<R>*)ext;
  cx<R>* c_ = (cx<R>*)canon;
  R* l_ = (R*)lmbds;
  R* m_ = (R*)colmax;
  R eps = (R)pinv_eps;
  void* params[] = {&D, &L, &e_, &c_, &l_, &m_, &eps};
  cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(wpb * 32), params, smem, st);
  if (e != cudaSuccess) return set_error("cudaLaunchKernel(canonicalize): %s", cudaGetErrorString(e));
  return after_launch("canonicalize(generic)");
}

template <typename R>
int launch_apply_update(int d, int D, int Dn, long long B, const void* T_in, void* T_out, const void* canon,
                        const void* lmbds, void* msgs_out, const int32_t* in_pos, const int32_t* out_pos,
                        const int32_t* lmbd_pos, const void* node_ampls, const void* edge_ampls, double ztime,
                        double xtime, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B == 0) return 0;
  if (int rc = check_ws<R>(d, D, Dn, ws_bytes)) return rc;
  NodeArgs<R> a{};
  a.d = d; a.D = D; a.Dn = Dn; a.B = B;
  a.T = (const cx<R>*)T_in; a.Tout = (cx<R>*)T_out; a.canon = (const cx<R>*)canon; a.lmbds = (const R*)lmbds;
  a.msgs_out = (cx<R>*)msgs_out; a.in_pos = in_pos; a.out_pos = out_pos; a.lmbd_pos = lmbd_pos;
  a.node_ampls = (const R*)node_ampls; a.edge_ampls = (const R*)edge_ampls;
  a.ztime = (R)ztime; a.xtime = (R)xtime;
  a.ws = (cx<R>*)ws; a.ws_per_warp = generic_ws_elems_per_warp(d, D, Dn);
  k_apply_update<R><<<node_grid(B, 4), 128, 0, st>>>(a);
  return after_launch("apply_update(generic)");
}

template <typename R>
int launch_density(int d, int D, long long B, const void* T, const void* msgs, const int32_t* in_pos,
                   const int32_t* node_ids, void* bloch, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B == 0) return 0;
  if (int rc = check_ws<R>(d, D, D, ws_bytes)) return rc;
  NodeArgs<R> a{};
  a.d = d; a.D = D; a.Dn = D; a.B = B;
  a.T = (const cx<R>*)T; a.msgs_cur = (const cx<R>*)msgs; a.in_pos = in_pos; a.node_ids = node_ids;
  a.bloch = (R*)bloch; a.ws = (cx<R>*)ws; a.ws_per_warp = generic_ws_elems_per_warp(d, D, D);
  k_density<R><<<node_grid(B, 4), 128, 0, st>>>(a);
  return after_launch("density(generic)");
}

template <typename R>
int launch_argmax(long long N, const void* bloch, const int32_t* outcomes, int32_t* result, void* result_p0,
                  cudaStream_t st) {
  k_argmax_unmeasured<R><<<1, 1024, 0, st>>>(N, (const R*)bloch, outcomes, result, (R*)result_p0);
  return after_launch("argmax_unmeasured");
}

template <typename R>
int launch_project(int d, int D, void* T, long long pos, int bit, cudaStream_t st) {
  int half = 1;
  for (int i = 0; i < d; ++i) half *= D;
  k_project_node<R><<<1, 32, 0, st>>>(half, (cx<R>*)T, pos, bit);
  return after_launch("project_node");
}

template <typename R>
int launch_threshold(int d, int D, long long B, void* T, const int32_t* node_ids, const void* bloch,
                     int32_t* outcomes, double thr, int32_t* n_proj, cudaStream_t st) {
  if (B == 0) return 0;
  int half = 1;
  for (int i = 0; i < d; ++i) half *= D;
  k_threshold_project<R><<<node_grid(B, 4), 128, 0, st>>>(half, B, (cx<R>*)T, node_ids, (const R*)bloch, outcomes,
                                                           (R)thr, n_proj);
  return after_launch("threshold_project");
}

#define BQA_INSTANTIATE(R)                                                                                         \
  template int launch_node_msgs<R>(bool, int, int, long long, const void*, const void*, void*, const int32_t*,     \
                                   const int32_t*, const void*, double, double, int, double, int, void*, int32_t*, \
                                   void*, size_t, const int32_t*, void* const*, cudaStream_t);                                                   \
  template int launch_canonicalize<R>(int, long long, const void*, void*, void*, void*, double, cudaStream_t, bool); \
  template int launch_apply_update<R>(int, int, int, long long, const void*, void*, const void*, const void*,      \
                                      void*, const int32_t*, const int32_t*, const int32_t*, const void*,          \
                                      const void*, double, double, void*, size_t, cudaStream_t);                   \
  template int launch_density<R>(int, int, long long, const void*, const void*, const int32_t*, const int32_t*,    \
                                 void*, void*, size_t, cudaStream_t);                                              \
  template int launch_argmax<R>(long long, const void*, const int32_t*, int32_t*, void*, cudaStream_t);            \
  template int launch_project<R>(int, int, void*, long long, int, cudaStream_t);                                   \
  template int launch_threshold<R>(int, int, long long, void*, const int32_t*, const void*, int32_t*, double,      \
                                   int32_t*, cudaStream_t);

}  // namespace bqa
