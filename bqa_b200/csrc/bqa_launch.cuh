// bqa_launch.cuh -- host-side launch plumbing shared by the .cu files (error string, launch counter,
// declarations of the templated launchers).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#define BQA_GENERIC_MAX_WARPS (148 * 32)
#define BQA_MAX_PEERS 8

struct bqa_b200_class;          // include/bqa_b200.h

namespace bqa {

int set_error(const char* fmt, ...);
int after_launch(const char* what);      // counts the launch, maps cudaGetLastError to a return code

// complex elements of per-warp scratch the generic node kernels need (P, E, gram | bufA, bufB, wbuf)
inline size_t generic_ws_elems_per_warp(int d, int D, int Dn) {
  const int Dm = D > Dn ? D : Dn;
  size_t W = 2;
  for (int i = 0; i < d; ++i) W *= Dm;
  const size_t msgs_part = 2 * W + (size_t)d * 2 * D * D;
  const size_t upd_part = 2 * W + (size_t)2 * D * Dm;
  return (msgs_part > upd_part ? msgs_part : upd_part) + 8;
}

template <typename R>
int launch_node_msgs(bool ext, int d, int D, long long B, const void* T, const void* msgs_cur, void* msgs_out,
                     const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls, double ztime,
                     double damping, int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                     void* ws, size_t ws_bytes, const int32_t* remote_pos, void* const* peers, cudaStream_t st);
template <typename R>
int launch_canonicalize(int D, long long L, const void* ext, void* canon, void* lmbds, void* colmax,
                        double pinv_eps, cudaStream_t st, bool round_robin);
template <typename R>
int launch_apply_update(int d, int D, int Dn, long long B, const void* T_in, void* T_out, const void* canon,
                        const void* lmbds, void* msgs_out, const int32_t* in_pos, const int32_t* out_pos,
                        const int32_t* lmbd_pos, const void* node_ampls, const void* edge_ampls, double ztime,
                        double xtime, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename R>
int launch_density(int d, int D, long long B, const void* T, const void* msgs, const int32_t* in_pos,
                   const int32_t* node_ids, void* bloch, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename R>
int launch_argmax(long long N, const void* bloch, const int32_t* outcomes, int32_t* result, void* result_p0,
                  cudaStream_t st);
template <typename R>
int launch_project(int d, int D, void* T, long long pos, int bit, cudaStream_t st);
template <typename R>
int launch_threshold(int d, int D, long long B, void* T, const int32_t* node_ids, const void* bloch,
                     int32_t* outcomes, double thr, int32_t* n_proj, cudaStream_t st);

// all degree classes in one launch (bqa_multiclass.cuh); kind: 0 extended messages, 1 apply update, 2 whole BP run
template <typename R>
int launch_multiclass(int kind, int n_classes, const bqa_b200_class* cls, int D, int Dn, void* msgs0, void* msgs1,
                      int parity, const void* canon, const void* lmbds, double ztime, double xtime, double damping,
                      double bp_eps, int max_iters, void* resid, int32_t* status, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool allow_fast);

// specialised kernels (bqa_fast_d3D4.cu): degree 3, D = 4, complex64
bool fast_d3D4_available(int prec, int degree, int D, long long B);
// extended messages enqueued behind a single-launch BP run: the buffers and status words the kernel selects from
struct AfterRun {
  const void* msgs[3];
  int nbuf, parity, max_iters;
  const int32_t* status;
};
int launch_fast_msgs_d3D4(bool ext, long long B, const void* T, const void* msgs_cur, void* msgs_out,
                          const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls, double ztime,
                          double damping, int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                          const int32_t* remote_pos, void* const* peers, cudaStream_t st,
                          const AfterRun* after = nullptr);
// the whole BP run of the degree class in one cooperative launch (bqa_fast_d3D4.cu)
int launch_fast_bp_run_d3D4(long long B, const void* T, void* msgs0, void* msgs1, int parity, const int32_t* in_pos,
                            const int32_t* out_pos, double damping, double bp_eps, int max_iters, void* resid,
                            int32_t* status, const int32_t* remote_pos, void* const* peers0, void* const* peers1, int rank,
                            int world, void* const* peer_resid, void* const* peer_flags, unsigned seq_base,
                            void* msgs2, void* const* peers2, long long boundary_nodes, cudaStream_t st);
// symmetric-gauge messages of every slot, msgs[p] = diag(lambda[p mod L][:Dn]) / trace      (state.py:56-57)
template <typename R>
int launch_gauge_msgs(int D_old, int Dn, long long L, const void* lmbds, void* msgs_out, cudaStream_t st);
// msgs_out[out_pos[i]] = diag(lambda[lmbd_pos[i]][:Dn]) / trace for a list of (slot, lambda row) pairs
template <typename R>
int launch_gauge_slots(int D_old, int Dn, long long n, const int32_t* out_pos, const int32_t* lmbd_pos, const void* lmbds,
                       void* msgs_out, cudaStream_t st);
// specialised simple-update application (bqa_fast_apply.cu): degree 3, D = 4 -> 4, complex64
bool fast_apply_available(int prec, int degree, int D, int Dn, long long B);
int launch_fast_apply_d3D4(long long B, const void* T_in, void* T_out, const void* canon, const void* lmbds,
                           const int32_t* in_pos, const int32_t* lmbd_pos, const void* node_ampls,
                           const void* edge_ampls, double ztime, double xtime, cudaStream_t st);
// cross-GPU sweep epilogue over peer memory: residual max to every peer + barrier (bqa_sync.cu)
int launch_sweep_sync(int prec, int rank, int world, void* const* peer_resid, int it, void* const* peer_flags,
                      unsigned seq, int32_t* status, cudaStream_t st);

// cycles a grid / peer barrier waits before it sets status[3] and gives up (bqa_b200_set_barrier_timeout)
namespace fast {
long long barrier_timeout_cycles();
void set_barrier_timeout_cycles(long long cycles);
void set_bp_trace(void* device_buffer);
}

// specialised canonicalizer kernel (bqa_fast_canon8.cu): D = 4 (n = 8), complex64
int launch_fast_canon8(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                       int ncols, cudaStream_t st);

void canon8_stats(unsigned long long* out3);
// second design (bqa_fast_canon8v2.cu): Cholesky factor + column Jacobi for the eigenproblems, stacked SVD of ker
int launch_fast_canon8v2(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                         int ncols, const int32_t* order, void* cost, cudaStream_t st, long long n_edges = -1,
                         const int32_t* remote = nullptr, void* const* peer_canon = nullptr,
                         void* const* peer_lmbds = nullptr, const long long* peer_L = nullptr);
int launch_sort_edges_by_cost(long long L, const void* cost, int32_t* order, cudaStream_t st);
void canon8v2_stats(unsigned long long* out3);
// third layout (bqa_fast_canon8v3.cu): two lanes per matrix, 16 warps per SM
int launch_fast_canon8v3(long long L, const void* ext, void* canon, void* lmbds, void* colmax, double pinv_eps,
                         int ncols, cudaStream_t st);
void canon8v3_stats(unsigned long long* out7);
void canon8v2_stats_detail(unsigned long long* out7);
void canon8v2_span(unsigned long long* out2);

}  // namespace bqa
