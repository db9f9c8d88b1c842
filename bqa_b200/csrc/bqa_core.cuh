// bqa_core.cuh -- per-node / per-edge math of the BP annealing path, written once as
// group-cooperative routines:
//   * on the GPU a "group" is a warp (generic kernels) -- lanes stride over the work, `sync()` is
//     __syncwarp and reductions are shuffles;
//   * under a plain C++ compiler (tests/hostemu, TEST ONLY) a group is one serial thread, which lets
//     the CPU test-suite check this exact source against the oracle without a GPU.
// The product library never runs these routines on the host (the C ABI only launches kernels).
//
// Reference semantics (file:line are relative to the bqa checkout):
//   node_gram ............ Tensor._apply_msgs_but_one/_compute_msg/pass_msgs  src/bqa/backends.py:381-408
//   ext_msg_entry ........ _apply_conditional_z_gate_to_single_axis           src/bqa/backends.py:519-526
//   jacobi_svd + edge_canonicalize .. _get_canonicalizers/decompose_iden_using_msgs/batched_svd/pinv_raw
//                                     src/bqa/state.py:171-200, src/bqa/backends.py:483-490,709-727
//   node_apply_update .... apply_canonicalizers_with_extensions + apply_z_gates + apply_x_gates +
//                          mul_by_lmbds  src/bqa/backends.py:416-432,506-510,450-462; src/bqa/state.py:219-227
//   node_density ......... get_density_matrices src/bqa/backends.py:440-448, utils.py:23-27
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BQA_HD __host__ __device__ __forceinline__
#define BQA_HDN __host__ __device__
#else
#define BQA_HD inline
#define BQA_HDN
#endif

#define BQA_MAX_D 16          // largest supported bond dimension (ext. dimension 2D <= 32)
#define BQA_MAX_N (2 * BQA_MAX_D)
#define BQA_MAX_DEGREE 8

namespace bqa {

// ------------------------------------------------------------------------------------------------
// complex numbers (interleaved re, im -- binary compatible with numpy / torch complex64/128)
// ------------------------------------------------------------------------------------------------
template <typename R>
struct cx {
  R re, im;
};

template <typename R> BQA_HD cx<R> mk(R a, R b) { cx<R> r; r.re = a; r.im = b; return r; }
template <typename R> BQA_HD cx<R> operator+(cx<R> a, cx<R> b) { return mk<R>(a.re + b.re, a.im + b.im); }
template <typename R> BQA_HD cx<R> operator-(cx<R> a, cx<R> b) { return mk<R>(a.re - b.re, a.im - b.im); }
template <typename R> BQA_HD cx<R> operator*(cx<R> a, cx<R> b) {
  return mk<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename R> BQA_HD cx<R> operator*(R s, cx<R> a) { return mk<R>(s * a.re, s * a.im); }
template <typename R> BQA_HD cx<R> conj(cx<R> a) { return mk<R>(a.re, -a.im); }
template <typename R> BQA_HD R norm2(cx<R> a) { return a.re * a.re + a.im * a.im; }
// acc += a * b
template <typename R> BQA_HD void cmac(cx<R>& acc, cx<R> a, cx<R> b) {
  acc.re += a.re * b.re; acc.re -= a.im * b.im;
  acc.im += a.re * b.im; acc.im += a.im * b.re;
}
// acc += conj(a) * b
template <typename R> BQA_HD void cmacc(cx<R>& acc, cx<R> a, cx<R> b) {
  acc.re += a.re * b.re; acc.re += a.im * b.im;
  acc.im += a.re * b.im; acc.im -= a.im * b.re;
}
template <typename R> BQA_HD cx<R> cinv(cx<R> a) {
  R d = R(1) / norm2(a);
  return mk<R>(a.re * d, -a.im * d);
}

// precision-exact math wrappers (the global ::sqrt(float) resolves to double on the host)
BQA_HD float msqrt(float x) { return sqrtf(x); }
BQA_HD double msqrt(double x) { return sqrt(x); }
BQA_HD float mabs(float x) { return fabsf(x); }
BQA_HD double mabs(double x) { return fabs(x); }
BQA_HD float mcos(float x) { return cosf(x); }
BQA_HD double mcos(double x) { return cos(x); }
BQA_HD float msin(float x) { return sinf(x); }
BQA_HD double msin(double x) { return sin(x); }

template <typename R> struct num_traits;
template <> struct num_traits<float> { static BQA_HD float eps() { return 1.1920929e-07f; } };
template <> struct num_traits<double> { static BQA_HD double eps() { return 2.220446049250313e-16; } };

// ------------------------------------------------------------------------------------------------
// groups
// ------------------------------------------------------------------------------------------------
struct GroupSerial {
  BQA_HD int rank() const { return 0; }
  BQA_HD int size() const { return 1; }
  BQA_HD void sync() const {}
  template <typename R> BQA_HD R sum(R v) const { return v; }
};

#if defined(__CUDACC__)
struct GroupWarp {
  __device__ __forceinline__ int rank() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  template <typename R> __device__ __forceinline__ R sum(R v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};
#endif

BQA_HD int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

// ------------------------------------------------------------------------------------------------
// mode products
// ------------------------------------------------------------------------------------------------
// buf viewed as [outer][D][inner]; every fibre v (fixed outer, inner) becomes m . v with m[a][b] row-major.
// One fibre per lane => safe in place.
// DT > 0: the bond dimension as a compile-time constant (same operations in the same order, loops unrolled, the fibre
// in registers instead of a dynamically indexed local array); DT = 0: any D <= BQA_MAX_D.
template <typename R, typename G, int DT = 0>
BQA_HDN void mode_product_inplace(G g, cx<R>* buf, int outer, int D_, int inner, const cx<R>* m) {
  const int D = DT ? DT : D_;
  const int nf = outer * inner;
  for (int f = g.rank(); f < nf; f += g.size()) {
    const int o = f / inner, i = f - o * inner;
    cx<R>* p = buf + (size_t)o * D * inner + i;
    cx<R> v[DT ? DT : BQA_MAX_D];
#pragma unroll
    for (int b = 0; b < (DT ? DT : D); ++b) v[b] = p[(size_t)b * inner];
#pragma unroll
    for (int a = 0; a < (DT ? DT : D); ++a) {
      cx<R> acc = mk<R>(0, 0);
#pragma unroll
      for (int b = 0; b < (DT ? DT : D); ++b) cmac(acc, m[a * D + b], v[b]);
      p[(size_t)a * inner] = acc;
    }
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// per-node Gram matrices of the BP update
//   gram[k][p][x][y] = sum_{legs != k} conj(T[p, a.., x, ..]) prod_{j != k} m_j[a_j, b_j] T[p, b.., y, ..]
// (message k before normalisation is gram[k][0] + gram[k][1]; the ZZ-extended message needs both parts).
// P and E are scratch buffers of 2 * D^d elements.  Leg contractions are shared through the prefix P
// (legs < k already contracted): (d-1) + d(d-1)/2 mode products instead of the reference's d(d-1).
// ------------------------------------------------------------------------------------------------
template <typename R, typename G, int DT = 0>
BQA_HDN void node_gram_impl(G g, int d, int D_, const cx<R>* T, const cx<R>* const* msgs, cx<R>* P, cx<R>* E,
                            cx<R>* gram) {
  const int D = DT ? DT : D_;
  const int W = 2 * ipow(D, d);
  for (int i = g.rank(); i < W; i += g.size()) P[i] = T[i];
  g.sync();
  for (int k = 0; k < d; ++k) {
    for (int i = g.rank(); i < W; i += g.size()) E[i] = P[i];
    g.sync();
    for (int j = k + 1; j < d; ++j)
      mode_product_inplace<R, G, DT>(g, E, 2 * ipow(D, j), D, ipow(D, d - 1 - j), msgs[j]);
    // closing contraction over everything but leg k
    const int pre = ipow(D, k), post = ipow(D, d - 1 - k);
    const int half = W / 2;
    for (int o = g.rank(); o < 2 * D * D; o += g.size()) {
      const int p = o / (D * D), x = (o / D) % D, y = o % D;
      cx<R> acc = mk<R>(0, 0);
      const cx<R>* tb = T + (size_t)p * half + (size_t)x * post;
      const cx<R>* eb = E + (size_t)p * half + (size_t)y * post;
      const int step = D * post;
      for (int a = 0; a < pre; ++a, tb += step, eb += step) {
#pragma unroll 4
        for (int i = 0; i < post; ++i) cmacc(acc, tb[i], eb[i]);
      }
      gram[(size_t)k * 2 * D * D + o] = acc;
    }
    g.sync();
    if (k + 1 < d) mode_product_inplace<R, G, DT>(g, P, 2 * pre, D, post, msgs[k]);
  }
}

// bond dimensions 2, 4, 8 get the compile-time variant (16 as well was measured and dropped: the unrolled 16 x 16
// products raise the kernels' register count from 96 to 128 and cost every shape occupancy: D = 16 BP run 544 -> 631 ms) (D = 8 in the r2 capture: index arithmetic, loop control and the
// dynamically indexed fibre made up most of the 7.5 G warp instructions of a 20k-node BP run)
template <typename R, typename G>
BQA_HDN void node_gram(G g, int d, int D, const cx<R>* T, const cx<R>* const* msgs, cx<R>* P, cx<R>* E,
                       cx<R>* gram) {
  if (D == 8) node_gram_impl<R, G, 8>(g, d, D, T, msgs, P, E, gram);
  else if (D == 4) node_gram_impl<R, G, 4>(g, d, D, T, msgs, P, E, gram);
  else if (D == 2) node_gram_impl<R, G, 2>(g, d, D, T, msgs, P, E, gram);
  else node_gram_impl<R, G, 0>(g, d, D, T, msgs, P, E, gram);
}

// complex principal square roots of cos(theta), sin(theta) and the ZZ half-gate factors
//   f0 = sqrt(cos th), f1 = e^{-i pi/4} sqrt(sin th)           (reference backends.py:20-22, 519-526)
template <typename R>
BQA_HD void zz_factors_cs(R c, R s, cx<R>& f0, cx<R>& f1);
template <typename R>
BQA_HD void zz_factors(R theta, cx<R>& f0, cx<R>& f1) {
  zz_factors_cs<R>(mcos(theta), msin(theta), f0, f1);
}
// the same from c = cos(theta), s = sin(theta)
template <typename R>
BQA_HD void zz_factors_cs(R c, R s, cx<R>& f0, cx<R>& f1) {
  f0 = (c >= R(0)) ? mk<R>(msqrt(c), R(0)) : mk<R>(R(0), msqrt(-c));
  const cx<R> rs = (s >= R(0)) ? mk<R>(msqrt(s), R(0)) : mk<R>(R(0), msqrt(-s));
  const R h = R(0.70710678118654752440);
  f1 = mk<R>(h, -h) * rs;
}

// ------------------------------------------------------------------------------------------------
// message epilogues (shared by the generic kernels and the test-only host emulation)
// ------------------------------------------------------------------------------------------------
// BP message k of a node from its Gram parts: new = (g0 + g1) / trace; folds max |new - old|^2 and
// max |new + old|^2 into (mnum, mden) (get_dist, backends.py:492-495) and writes the damped update
// alpha * old + (1 - alpha) * new (or the undamped new) into dst (backends.py:761-764).
template <typename R, typename G>
BQA_HDN void emit_bp_msg(G g, int D, const cx<R>* g0, const cx<R>* g1, const cx<R>* old, cx<R>* dst, R damping,
                         int write_undamped, R& mnum, R& mden, cx<R>* dst2 = nullptr) {
  cx<R> tr = mk<R>(0, 0);
  for (int x = 0; x < D; ++x) tr = tr + g0[x * D + x] + g1[x * D + x];
  const cx<R> itr = cinv(tr);
  for (int o = g.rank(); o < D * D; o += g.size()) {
    const cx<R> nw = itr * (g0[o] + g1[o]);
    const cx<R> od = old[o];
    // a NaN entry must not drop out of the maxima (the reference's np.abs().max() propagates it, state.py:113): it
    // becomes +inf, which survives max() and the bit-pattern atomic max, and makes the residual non-finite on the host
    R a = norm2(nw - od), b = norm2(nw + od);
    if (!(a == a)) a = R(INFINITY);
    if (!(b == b)) b = R(INFINITY);
    mnum = a > mnum ? a : mnum;
    mden = b > mden ? b : mden;
    const cx<R> w = write_undamped ? nw : (damping * od + (R(1) - damping) * nw);
    dst[o] = w;
    if (dst2) dst2[o] = w;                                // halo slot of the peer that owns the receiving node
  }
}

// ZZ-extended message (2D x 2D) from the Gram parts:
//   ext[(s,x),(s',y)] = conj(f_s) f_s' (g0[x,y] + (-1)^{s+s'} g1[x,y]) / trace      (state.py:127-139)
template <typename R, typename G>
BQA_HDN void emit_ext_msg(G g, int D, const cx<R>* g0, const cx<R>* g1, R theta, cx<R>* dst, cx<R>* dst2 = nullptr) {
  const int n = 2 * D;
  cx<R> f[2];
  zz_factors<R>(theta, f[0], f[1]);
  cx<R> tr = mk<R>(0, 0);
  for (int x = 0; x < D; ++x) tr = tr + g0[x * D + x] + g1[x * D + x];
  const cx<R> itr = cinv((norm2(f[0]) + norm2(f[1])) * tr);
  for (int o = g.rank(); o < n * n; o += g.size()) {
    const int r = o / n, c = o - r * n;
    const int s = r / D, x = r - s * D, sp = c / D, y = c - sp * D;
    const cx<R> a0 = g0[x * D + y], a1 = g1[x * D + y];
    const cx<R> v = (s == sp) ? (a0 + a1) : (a0 - a1);
    const cx<R> w = itr * (conj(f[s]) * f[sp] * v);
    dst[o] = w;
    if (dst2) dst2[o] = w;
  }
}

// symmetric-gauge message diag(lambda[:Dn]) / trace                              (state.py:56-57)
template <typename R, typename G>
BQA_HDN void emit_gauge_msg(G g, int Dn, const R* lam, cx<R>* dst) {
  R tr = 0;
  for (int c = 0; c < Dn; ++c) tr += lam[c];
  const R itr = R(1) / tr;
  for (int o = g.rank(); o < Dn * Dn; o += g.size()) {
    const int r = o / Dn, c = o - r * Dn;
    dst[o] = mk<R>(r == c ? lam[c] * itr : R(0), R(0));
  }
}

// projective measurement of one node onto `bit`: zero the other physical slice, renormalise the node
template <typename R, typename G>
BQA_HDN void node_project(G g, cx<R>* t, int half, int bit) {
  cx<R>* keep = t + (size_t)bit * half;
  cx<R>* kill = t + (size_t)(1 - bit) * half;
  R n2 = 0;
  for (int i = g.rank(); i < half; i += g.size()) { n2 += norm2(keep[i]); kill[i] = mk<R>(0, 0); }
  n2 = g.sum(n2);
  const R inv = R(1) / msqrt(n2);
  for (int i = g.rank(); i < half; i += g.size()) keep[i] = inv * keep[i];
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// one-sided (Hestenes) Jacobi SVD of a complex n x n matrix held row-major in `A` with row stride `ld`:
// on exit  A = U diag(sigma) (columns),  V = right singular vectors (same stride),  A_in = U diag(sigma) V^H.
// `order` receives the column permutation that sorts sigma descending.  Lane r owns row r, so a column access
// touches one element per row: in shared memory ld = n + 1 keeps those accesses off a common bank (with ld = n = 16
// complex64 every lane of a column access hit the same two banks: 12.7 conflicts per shared-memory instruction in
// the r2 capture of the n = 16 kernel).
// ------------------------------------------------------------------------------------------------
template <typename R, typename G>
BQA_HDN void jacobi_svd(G g, int n, int ld, cx<R>* A, cx<R>* V, R* sigma, int* order) {
  for (int i = g.rank(); i < n * n; i += g.size()) V[(i / n) * ld + i % n] = mk<R>((i / n == i % n) ? R(1) : R(0), R(0));
  g.sync();
  const R tol = num_traits<R>::eps() * R(2) * msqrt((R)n);
  // columns whose norm has fallen below eps * |A|_F are numerically zero (the absolute accuracy LAPACK's
  // gesdd works to, and far below every pinv_eps cut): they are left alone.  Rotating them is not only
  // wasted work -- once <a_p, a_q> reaches the denormal range the phase g / |g| stops being unimodular and
  // would rescale the healthy column of the pair.
  R fro2 = 0;
  for (int r = g.rank(); r < n; r += g.size())
    for (int j = 0; j < n; ++j) fro2 += norm2(A[r * ld + j]);
  fro2 = g.sum(fro2);
  const R nul = num_traits<R>::eps() * num_traits<R>::eps() * fro2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        R al = 0, be = 0, gr = 0, gi = 0;
        for (int r = g.rank(); r < n; r += g.size()) {
          const cx<R> ap = A[r * ld + p], aq = A[r * ld + q];
          al += norm2(ap);
          be += norm2(aq);
          gr += ap.re * aq.re + ap.im * aq.im;          // conj(ap) * aq
          gi += ap.re * aq.im - ap.im * aq.re;
        }
        al = g.sum(al); be = g.sum(be); gr = g.sum(gr); gi = g.sum(gi);
        const R g2 = gr * gr + gi * gi;
        if (al <= nul || be <= nul || g2 <= tol * tol * al * be) continue;
        rotated = true;
        const R ag = msqrt(g2);
        const R zeta = (be - al) / (R(2) * ag);
        const R t = ((zeta >= R(0)) ? R(1) : R(-1)) / (mabs(zeta) + msqrt(R(1) + zeta * zeta));
        const R c = R(1) / msqrt(R(1) + t * t), s = c * t;
        const cx<R> ph = mk<R>(gr / ag, -gi / ag);      // e^{-i arg(gamma)}
        for (int r = g.rank(); r < n; r += g.size()) {
          cx<R> ap = A[r * ld + p], aq = ph * A[r * ld + q];
          A[r * ld + p] = c * ap - s * aq;
          A[r * ld + q] = s * ap + c * aq;
          ap = V[r * ld + p]; aq = ph * V[r * ld + q];
          V[r * ld + p] = c * ap - s * aq;
          V[r * ld + q] = s * ap + c * aq;
        }
        g.sync();
      }
    }
    if (!rotated) break;
  }
  // singular values = column norms
  for (int j = 0; j < n; ++j) {
    R a = 0;
    for (int r = g.rank(); r < n; r += g.size()) a += norm2(A[r * ld + j]);
    a = g.sum(a);
    if (g.rank() == 0) sigma[j] = msqrt(a);
  }
  g.sync();
  if (g.rank() == 0) {
    for (int j = 0; j < n; ++j) order[j] = j;
    for (int i = 1; i < n; ++i) {                       // stable insertion sort, descending
      const int oi = order[i];
      int j = i - 1;
      while (j >= 0 && sigma[order[j]] < sigma[oi]) { order[j + 1] = order[j]; --j; }
      order[j + 1] = oi;
    }
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// canonicalizers of one undirected edge (forward message slot e, backward slot e + L)
//   scratch: 5 matrices of n rows with stride n + 1 (complex) + 3 n reals + 3 n ints  (see edge_scratch_elems);
//   V of the ker decomposition reuses the storage of V_f, which is dead once ker is built (at n = 32 the sixth matrix
//   cost a quarter of the warps that fit an SM)
// ------------------------------------------------------------------------------------------------
template <typename R>
BQA_HD size_t edge_scratch_elems(int n) { return (size_t)5 * n * (n + 1); }

// which SVD edge_canonicalize runs: the serial cyclic Jacobi above (any group), or a policy of the caller's
// (bqa_generic.cuh: round-robin ordering over the lanes of a warp for n = 16 / 32); `prm` = n / 2 x 4 reals of scratch
struct SerialJacobi {
  template <typename R, typename G>
  static BQA_HDN void run(G g, int n, int ld, cx<R>* A, cx<R>* V, R* sigma, int* order, R* /*prm*/) {
    jacobi_svd<R>(g, n, ld, A, V, sigma, order);
  }
};

template <typename R, typename G, typename SVD = SerialJacobi>
BQA_HDN void edge_canonicalize(G g, int n, const cx<R>* ext_f, const cx<R>* ext_b, R pinv_eps,
                               cx<R>* scratch, R* rscratch, int* iscratch,
                               cx<R>* canon_at_e /* backward */, cx<R>* canon_at_eL /* forward */,
                               R* lmbd_out /* n */, R* prm = nullptr) {
  const int nn = n * n, ld = n + 1, sz = n * ld;              // scratch matrices: row stride ld (see jacobi_svd)
  cx<R>* Af = scratch;           cx<R>* Vf = scratch + sz;
  cx<R>* Ab = scratch + 2 * sz;  cx<R>* Vb = scratch + 3 * sz;
  cx<R>* K = scratch + 4 * sz;   cx<R>* Vk = Vf;              // V_f is dead after the ker build below
  R* sf = rscratch;  R* sb = rscratch + n;  R* sk = rscratch + 2 * n;
  int* of = iscratch;  int* ob = iscratch + n;  int* ok = iscratch + 2 * n;
  for (int i = g.rank(); i < nn; i += g.size()) {
    const int at = (i / n) * ld + i % n;
    Af[at] = ext_f[i];
    Ab[at] = ext_b[i];
  }
  g.sync();
  SVD::template run<R>(g, n, ld, Af, Vf, sf, of, prm);
  SVD::template run<R>(g, n, ld, Ab, Vb, sb, ob, prm);
  // ker[i][j] = sum_k lu_f[i][k] lu_b[j][k],  lu[i][k] = sqrt(s_i) conj(V[k][col_i])  (masked s_i > pinv_eps)
  for (int o = g.rank(); o < nn; o += g.size()) {
    const int i = o / n, j = o - i * n;
    const int ci = of[i], cj = ob[j];
    cx<R> acc = mk<R>(0, 0);
    if (sf[ci] > pinv_eps && sb[cj] > pinv_eps) {
      for (int k = 0; k < n; ++k) cmac(acc, conj(Vf[k * ld + ci]), conj(Vb[k * ld + cj]));
      acc = msqrt(sf[ci] * sb[cj]) * acc;
    }
    K[i * ld + j] = acc;
  }
  g.sync();
  // ul = u * pinv(sqrt s): column c of A (= u_c s_c) scaled by s_c^{-3/2}; stored back into A columns
  for (int o = g.rank(); o < nn; o += g.size()) {
    const int c = o % n, at = (o / n) * ld + c;
    {
      const R s = sf[c];
      const R w = (s > pinv_eps && msqrt(s) > num_traits<R>::eps()) ? R(1) / (s * msqrt(s)) : R(0);
      Af[at] = w * Af[at];
    }
    {
      const R s = sb[c];
      const R w = (s > pinv_eps && msqrt(s) > num_traits<R>::eps()) ? R(1) / (s * msqrt(s)) : R(0);
      Ab[at] = w * Ab[at];
    }
  }
  g.sync();
  SVD::template run<R>(g, n, ld, K, Vk, sk, ok, prm);
  // lambda = masked singular values, L2 normalised (reference state.py:200)
  R nrm2 = 0;
  for (int j = 0; j < n; ++j) { const R s = sk[ok[j]]; if (s > pinv_eps) nrm2 += s * s; }
  const R inrm = R(1) / msqrt(nrm2);
  for (int j = g.rank(); j < n; j += g.size()) {
    const R s = sk[ok[j]];
    lmbd_out[j] = (s > pinv_eps) ? s * inrm : R(0);
  }
  // C_f = ul_f . U2 (slot e + L),  C_b = ul_b . conj(V2) (slot e); columns in sorted order, masked
  for (int o = g.rank(); o < nn; o += g.size()) {
    const int r = o / n, j = o - r * n;
    const int cj = ok[j];
    const R s = sk[cj];
    cx<R> cf = mk<R>(0, 0), cb = mk<R>(0, 0);
    if (s > pinv_eps) {
      const R is = R(1) / s;
      for (int k = 0; k < n; ++k) {
        // U2[k'][cj] = K[k'][cj] / s with k' indexing the *sorted* columns of the forward decomposition
        cmac(cf, Af[r * ld + of[k]], K[k * ld + cj]);
        cmac(cb, Ab[r * ld + ob[k]], conj(Vk[k * ld + cj]));
      }
      cf = is * cf;
    }
    canon_at_eL[o] = cf;
    canon_at_e[o] = cb;
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// simple-update application for one node:  T'[p, c..] = sum_a T[p, a..] prod_j W_j^{(p)}[a_j, c_j]
//   W_j^{(p)}[a, c] = (f0_j C_j[a, c] + (-1)^p f1_j C_j[D + a, c]) sqrt(lambda_j[c]),   c < Dn
// then Rz(phi) Rx(xt) on the physical leg and L2 normalisation.  The reference materialises the doubled
// legs; contracting the two halves of the canonicalizer with the half-gate factors first is the same sum.
//   bufA/bufB: 2 * max(D, Dn)^d elements each;  wbuf: 2 * D * Dn elements.
// ------------------------------------------------------------------------------------------------
template <typename R, typename G>
BQA_HDN void node_apply_update(G g, int d, int D, int Dn, const cx<R>* T, const cx<R>* const* canon,
                               const R* theta, const R* const* lam, R phi, R xt, cx<R>* bufA, cx<R>* bufB,
                               cx<R>* wbuf, cx<R>* Tout) {
  const int n = 2 * D;
  const cx<R>* src = T;
  cx<R>* dst = bufA;
  for (int j = 0; j < d; ++j) {
    cx<R> f0, f1;
    zz_factors<R>(theta[j], f0, f1);
    for (int o = g.rank(); o < 2 * D * Dn; o += g.size()) {
      const int p = o / (D * Dn), a = (o / Dn) % D, c = o % Dn;
      const cx<R> lo = f1 * canon[j][(D + a) * n + c];
      cx<R> w = f0 * canon[j][a * n + c];
      w = p ? (w - lo) : (w + lo);
      wbuf[o] = msqrt(lam[j][c]) * w;
    }
    g.sync();
    const int outer = ipow(Dn, j), inner = ipow(D, d - 1 - j);
    const int total = 2 * outer * Dn * inner;
    for (int o = g.rank(); o < total; o += g.size()) {
      const int i = o % inner, c = (o / inner) % Dn, oo = (o / (inner * Dn)) % outer, p = o / (inner * Dn * outer);
      const cx<R>* s = src + ((size_t)(p * outer + oo) * D) * inner + i;
      const cx<R>* w = wbuf + (size_t)p * D * Dn + c;
      cx<R> acc = mk<R>(0, 0);
      for (int a = 0; a < D; ++a) cmac(acc, s[(size_t)a * inner], w[a * Dn]);
      dst[o] = acc;
    }
    g.sync();
    src = dst;
    dst = (dst == bufA) ? bufB : bufA;
  }
  const int half = ipow(Dn, d);
  const R cp = mcos(phi), sp = msin(phi), cxt = mcos(xt), sxt = msin(xt);
  const cx<R> z0 = mk<R>(cp, -sp), z1 = mk<R>(cp, sp);       // Rz: e^{-i phi}, e^{+i phi}
  R nrm2 = 0;
  for (int i = g.rank(); i < half; i += g.size()) {
    const cx<R> v0 = z0 * src[i], v1 = z1 * src[half + i];
    // Rx: out_p = cos(xt) v_p - i sin(xt) v_{1-p}
    const cx<R> o0 = mk<R>(cxt * v0.re + sxt * v1.im, cxt * v0.im - sxt * v1.re);
    const cx<R> o1 = mk<R>(cxt * v1.re + sxt * v0.im, cxt * v1.im - sxt * v0.re);
    dst[i] = o0;
    dst[half + i] = o1;
    nrm2 += norm2(o0) + norm2(o1);
  }
  nrm2 = g.sum(nrm2);
  g.sync();
  const R inv = R(1) / msqrt(nrm2);
  for (int i = g.rank(); i < 2 * half; i += g.size()) Tout[i] = inv * dst[i];
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// single-qubit marginal: rho[p][q] = sum T[p, b] prod m_j[a_j, b_j] conj(T[q, a]), trace-normalised.
// out4 = (bloch x, y, z, p0).
// ------------------------------------------------------------------------------------------------
template <typename R, typename G>
BQA_HDN void node_density(G g, int d, int D, const cx<R>* T, const cx<R>* const* msgs, cx<R>* E, R* out4) {
  const int half = ipow(D, d), W = 2 * half;
  for (int i = g.rank(); i < W; i += g.size()) E[i] = T[i];
  g.sync();
  for (int j = 0; j < d; ++j) mode_product_inplace<R>(g, E, 2 * ipow(D, j), D, ipow(D, d - 1 - j), msgs[j]);
  R r00 = 0, r11 = 0, r01r = 0, r01i = 0, r10r = 0, r10i = 0;
  for (int i = g.rank(); i < half; i += g.size()) {
    const cx<R> e0 = E[i], e1 = E[half + i], t0 = T[i], t1 = T[half + i];
    cx<R> a = mk<R>(0, 0);
    cmacc(a, t0, e0); r00 += a.re;                      // rho00 = sum e0 conj(t0) (real part; imaginary ~ 0)
    a = mk<R>(0, 0); cmacc(a, t1, e1); r11 += a.re;
    a = mk<R>(0, 0); cmacc(a, t1, e0); r01r += a.re; r01i += a.im;     // rho01 = sum e0 conj(t1)
    a = mk<R>(0, 0); cmacc(a, t0, e1); r10r += a.re; r10i += a.im;     // rho10 = sum e1 conj(t0)
  }
  r00 = g.sum(r00); r11 = g.sum(r11); r01r = g.sum(r01r); r01i = g.sum(r01i); r10r = g.sum(r10r); r10i = g.sum(r10i);
  if (g.rank() == 0) {
    const R it = R(1) / (r00 + r11);
    out4[0] = (r01r + r10r) * it;
    out4[1] = (r10i - r01i) * it;
    out4[2] = (r00 - r11) * it;
    out4[3] = r00 * it;
  }
  g.sync();
}

}  // namespace bqa
