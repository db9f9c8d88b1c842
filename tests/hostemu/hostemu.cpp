// hostemu.cpp -- TEST ONLY.  Builds the C ABI of include/bqa_b200.h on the host from the very same math
// header the CUDA kernels use (bqa_b200/csrc/bqa_core.cuh) with a serial one-lane "group", so that the
// CPU test-suite (no GPU in the build container) can exercise the kernel math and the whole Python engine
// against the oracle.  It is compiled by tests/hostemu/build.py into tests/hostemu/_build/ and is never
// loaded by the bqa_b200 package: the product loads bqa_b200/libbqa_b200.so (CUDA) or fails.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/bqa_b200.h"
#include "../../bqa_b200/csrc/bqa_core.cuh"

using namespace bqa;

static char g_err[256] = "";
static long long g_calls = 0;

template <typename R>
static void bp_or_ext(bool ext, int d, int D, long long B, const void* T_, const void* cur_, void* out_,
                      const int32_t* in_pos, const int32_t* out_pos, const void* ea_, double ztime, double damping,
                      int write_undamped, double bp_eps, int it, void* resid_, int32_t* status) {
  const cx<R>* T = (const cx<R>*)T_;
  const cx<R>* cur = (const cx<R>*)cur_;
  cx<R>* out = (cx<R>*)out_;
  R* resid = (R*)resid_;
  if (!ext && it > 0) {
    if (status[0] != 0) return;
    if (std::sqrt(resid[2 * (it - 1)] / resid[2 * (it - 1) + 1]) < (R)bp_eps) { status[1] = it; status[0] = 1; return; }
  }
  GroupSerial g;
  const int DD = D * D, W = 2 * ipow(D, d);
  std::vector<cx<R>> P(W), E(W), gram((size_t)std::max(d, 1) * 2 * DD);
  R mnum = 0, mden = 0;
  for (long long node = 0; node < B; ++node) {
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = cur + (size_t)in_pos[(size_t)j * B + node] * DD;
    node_gram<R>(g, d, D, T + (size_t)node * W, mp, P.data(), E.data(), gram.data());
    for (int k = 0; k < d; ++k) {
      const cx<R>* g0 = gram.data() + (size_t)k * 2 * DD;
      const cx<R>* g1 = g0 + DD;
      const size_t slot = (size_t)out_pos[(size_t)k * B + node];
      if (!ext)
        emit_bp_msg<R>(g, D, g0, g1, cur + slot * DD, out + slot * DD, (R)damping, write_undamped, mnum, mden);
      else
        emit_ext_msg<R>(g, D, g0, g1, ((const R*)ea_)[(size_t)k * B + node] * (R)ztime, out + slot * 4 * DD);
    }
  }
  if (!ext) {
    resid[2 * it] = std::max(resid[2 * it], mnum);
    resid[2 * it + 1] = std::max(resid[2 * it + 1], mden);
  }
}

template <typename R>
static void canonicalize(int D, long long L, const void* ext_, void* canon_, void* lmbds_, void* colmax_, double eps) {
  GroupSerial g;
  const int n = 2 * D, nn = n * n;
  std::vector<cx<R>> scratch(edge_scratch_elems<R>(n));
  std::vector<R> rs(3 * n);
  std::vector<int> is(3 * n);
  const cx<R>* ext = (const cx<R>*)ext_;
  cx<R>* canon = (cx<R>*)canon_;
  R* lmbds = (R*)lmbds_;
  R* colmax = (R*)colmax_;
  for (long long e = 0; e < L; ++e) {
    edge_canonicalize<R>(g, n, ext + (size_t)e * nn, ext + (size_t)(e + L) * nn, (R)eps, scratch.data(), rs.data(),
                         is.data(), canon + (size_t)e * nn, canon + (size_t)(e + L) * nn, lmbds + (size_t)e * n);
    for (int j = 0; j < n; ++j) colmax[j] = std::max(colmax[j], lmbds[(size_t)e * n + j]);
  }
}

template <typename R>
static void apply_update(int d, int D, int Dn, long long B, const void* Tin_, void* Tout_, const void* canon_,
                         const void* lmbds_, void* msgs_, const int32_t* in_pos, const int32_t* out_pos,
                         const int32_t* lmbd_pos, const void* na_, const void* ea_, double ztime, double xtime) {
  GroupSerial g;
  const int n = 2 * D, Win = 2 * ipow(D, d), Wout = 2 * ipow(Dn, d), Wmax = 2 * ipow(std::max(D, Dn), d);
  std::vector<cx<R>> A(Wmax), Bf(Wmax), wb(2 * D * Dn);
  for (long long node = 0; node < B; ++node) {
    const cx<R>* cp[BQA_MAX_DEGREE];
    const R* lp[BQA_MAX_DEGREE];
    R th[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) {
      cp[j] = (const cx<R>*)canon_ + (size_t)in_pos[(size_t)j * B + node] * n * n;
      lp[j] = (const R*)lmbds_ + (size_t)lmbd_pos[(size_t)j * B + node] * n;
      th[j] = ((const R*)ea_)[(size_t)j * B + node] * (R)ztime;
    }
    node_apply_update<R>(g, d, D, Dn, (const cx<R>*)Tin_ + (size_t)node * Win, cp, th, lp,
                         ((const R*)na_)[node] * (R)ztime, (R)xtime, A.data(), Bf.data(), wb.data(),
                         (cx<R>*)Tout_ + (size_t)node * Wout);
    for (int j = 0; j < d; ++j)
      emit_gauge_msg<R>(g, Dn, lp[j], (cx<R>*)msgs_ + (size_t)out_pos[(size_t)j * B + node] * Dn * Dn);
  }
}

template <typename R>
static void density(int d, int D, long long B, const void* T_, const void* msgs_, const int32_t* in_pos,
                    const int32_t* node_ids, void* bloch_) {
  GroupSerial g;
  const int DD = D * D, W = 2 * ipow(D, d);
  std::vector<cx<R>> E(W);
  for (long long node = 0; node < B; ++node) {
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = (const cx<R>*)msgs_ + (size_t)in_pos[(size_t)j * B + node] * DD;
    node_density<R>(g, d, D, (const cx<R>*)T_ + (size_t)node * W, mp, E.data(), (R*)bloch_ + (size_t)node_ids[node] * 4);
  }
}

template <typename R>
static void argmax(long long N, const void* bloch_, const int32_t* outcomes, int32_t* result, void* p0_) {
  const R* bloch = (const R*)bloch_;
  R best = -1;
  long long bi = -1;
  int cnt = 0;
  for (long long i = 0; i < N; ++i) {
    if (outcomes[i] != 0) continue;
    ++cnt;
    const R key = std::fabs(R(2) * bloch[4 * i + 3] - R(1));
    if (key > best) { best = key; bi = i; }
  }
  result[0] = (int32_t)bi;
  result[1] = cnt;
  ((R*)p0_)[0] = bi >= 0 ? bloch[4 * bi + 3] : R(0);
}

template <typename R>
static void threshold(int d, int D, long long B, void* T_, const int32_t* node_ids, const void* bloch_,
                      int32_t* outcomes, double thr, int32_t* n_proj) {
  const int half = ipow(D, d);
  for (long long node = 0; node < B; ++node) {
    const int32_t id = node_ids[node];
    if (outcomes[id] != 0) continue;
    const R p0 = ((const R*)bloch_)[4 * (size_t)id + 3];
    int bit = -1;
    if (p0 > (R)thr) bit = 0;
    else if (p0 < R(1) - (R)thr) bit = 1;
    if (bit < 0) continue;
    node_project<R>(GroupSerial(), (cx<R>*)T_ + (size_t)node * 2 * half, half, bit);
    outcomes[id] = 1 - 2 * bit;
    ++*n_proj;
  }
}

template <typename R>
static void gauge_msgs(int Dold, int Dn, long long L, const void* lm_, void* msgs_) {
  GroupSerial g;
  for (long long p = 0; p < 2 * L; ++p)
    emit_gauge_msg<R>(g, Dn, (const R*)lm_ + (size_t)(p % L) * 2 * Dold, (cx<R>*)msgs_ + (size_t)p * Dn * Dn);
}

#define DISPATCH(call_f, call_d) do { ++g_calls; if (prec == BQA_C64) { call_f; } else { call_d; } return 0; } while (0)

extern "C" {
const char* bqa_b200_last_error(void) { return g_err; }
int bqa_b200_version(void) { return -1; }   // negative: host emulation
long long bqa_b200_launch_count(void) { return g_calls; }
int bqa_b200_set_kernel_mode(int) { return 0; }
int bqa_b200_canon_stats(unsigned long long* o) { o[0] = o[1] = o[2] = 0; return 0; }
int bqa_b200_set_barrier_timeout(double) { return 0; }
int bqa_b200_set_bp_trace(void*) { return 0; }
int bqa_b200_canon_span(unsigned long long* o) { o[0] = o[1] = 0; return 0; }
int bqa_b200_canon_stats_detail(unsigned long long* o) { for (int i = 0; i < 7; ++i) o[i] = 0; return 0; }
int bqa_b200_gauge_msgs(int prec, int D_old, int D_new, long long L, const void* lmbds, void* msgs_out, void*) {
  DISPATCH(gauge_msgs<float>(D_old, D_new, L, lmbds, msgs_out), gauge_msgs<double>(D_old, D_new, L, lmbds, msgs_out));
}
static int no_p2p() { snprintf(g_err, sizeof(g_err), "peer-memory entry points need the CUDA build"); return 1; }
int bqa_b200_bp_sweep_p2p(int, int, int, long long, const void*, const void*, void*, const int32_t*, const int32_t*, double,
                          int, double, int, void*, int32_t*, void*, size_t, const int32_t*, void* const*, void*) { return no_p2p(); }
int bqa_b200_ext_msgs_p2p(int, int, int, long long, const void*, const void*, void*, const int32_t*, const int32_t*,
                          const void*, double, void*, size_t, const int32_t*, void* const*, void*) { return no_p2p(); }
int bqa_b200_sweep_sync(int, int, int, void* const*, int, void* const*, unsigned, int32_t*, void*) { return no_p2p(); }
int bqa_b200_bp_run(int, int, int, long long, const void*, void*, void*, int, const int32_t*, const int32_t*, double, double,
                    int, void*, int32_t*, const int32_t*, void* const*, void* const*, int, int, void* const*, void* const*,
                    unsigned, void*, void* const*, long long, void*) { return 2; }   // "no single-launch kernel": the engine enqueues sweeps
size_t bqa_b200_workspace_bytes(int, int, int, int) { return 16; }

int bqa_b200_bp_sweep(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur, void* msgs_nxt,
                      const int32_t* in_pos, const int32_t* out_pos, double damping, int write_undamped, double bp_eps,
                      int it, void* resid, int32_t* status, void*, size_t, void*) {
  DISPATCH(bp_or_ext<float>(false, degree, D, B, T, msgs_cur, msgs_nxt, in_pos, out_pos, nullptr, 0, damping, write_undamped, bp_eps, it, resid, status),
           bp_or_ext<double>(false, degree, D, B, T, msgs_cur, msgs_nxt, in_pos, out_pos, nullptr, 0, damping, write_undamped, bp_eps, it, resid, status));
}
int bqa_b200_ext_msgs(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur, void* ext,
                      const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls, double ztime, void*, size_t,
                      void*) {
  DISPATCH(bp_or_ext<float>(true, degree, D, B, T, msgs_cur, ext, in_pos, out_pos, edge_ampls, ztime, 0, 0, 0, 0, nullptr, nullptr),
           bp_or_ext<double>(true, degree, D, B, T, msgs_cur, ext, in_pos, out_pos, edge_ampls, ztime, 0, 0, 0, 0, nullptr, nullptr));
}
int bqa_b200_canonicalize(int prec, int D, long long L, const void* ext, void* canon, void* lmbds, void* colmax,
                          double pinv_eps, int /*n_cols*/, void*) {
  DISPATCH(canonicalize<float>(D, L, ext, canon, lmbds, colmax, pinv_eps),
           canonicalize<double>(D, L, ext, canon, lmbds, colmax, pinv_eps));
}
int bqa_b200_apply_update(int prec, int degree, int D, int D_new, long long B, const void* T_in, void* T_out,
                          const void* canon, const void* lmbds, void* msgs_out, const int32_t* in_pos,
                          const int32_t* out_pos, const int32_t* lmbd_pos, const void* node_ampls,
                          const void* edge_ampls, double ztime, double xtime, void*, size_t, void*) {
  DISPATCH(apply_update<float>(degree, D, D_new, B, T_in, T_out, canon, lmbds, msgs_out, in_pos, out_pos, lmbd_pos, node_ampls, edge_ampls, ztime, xtime),
           apply_update<double>(degree, D, D_new, B, T_in, T_out, canon, lmbds, msgs_out, in_pos, out_pos, lmbd_pos, node_ampls, edge_ampls, ztime, xtime));
}
int bqa_b200_density(int prec, int degree, int D, long long B, const void* T, const void* msgs, const int32_t* in_pos,
                     const int32_t* node_ids, void* bloch, void*, size_t, void*) {
  DISPATCH(density<float>(degree, D, B, T, msgs, in_pos, node_ids, bloch),
           density<double>(degree, D, B, T, msgs, in_pos, node_ids, bloch));
}
int bqa_b200_argmax_unmeasured(int prec, long long N, const void* bloch, const int32_t* outcomes, int32_t* result,
                               void* result_p0, void*) {
  DISPATCH(argmax<float>(N, bloch, outcomes, result, result_p0), argmax<double>(N, bloch, outcomes, result, result_p0));
}
int bqa_b200_project_node(int prec, int degree, int D, void* T, long long pos, int bit, void*) {
  const int half = ipow(D, degree);
  DISPATCH(node_project<float>(GroupSerial(), (cx<float>*)T + (size_t)pos * 2 * half, half, bit),
           node_project<double>(GroupSerial(), (cx<double>*)T + (size_t)pos * 2 * half, half, bit));
}
int bqa_b200_threshold_project(int prec, int degree, int D, long long B, void* T, const int32_t* node_ids,
                               const void* bloch, int32_t* outcomes, double thr, int32_t* n_projected, void*) {
  DISPATCH(threshold<float>(degree, D, B, T, node_ids, bloch, outcomes, thr, n_projected),
           threshold<double>(degree, D, B, T, node_ids, bloch, outcomes, thr, n_projected));
}
}
