/* bqa_b200.h -- C ABI of libbqa_b200.so: the sm_100a kernels behind bqa's backend interface.
 *
 * The reference (LuchnikovI/bqa v0.1.6) has no FFI of its own: its plugin boundary is the Python ABC
 * `bqa.backends.Tensor` plus the registry dict `BACKEND_STR_TO_BACKEND` (src/bqa/backends.py:26-252) and
 * the engine calls in src/bqa/state.py.  Each entry point below is the fused device-side replacement of
 * one group of those calls; the Python side (bqa_b200/engine.py, bqa_b200/backend.py) binds them with
 * ctypes.  See INTEGRATION.md for the binding a bqa maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - complex arrays are interleaved (re, im): complex64 when prec == BQA_C64, complex128 when BQA_C128;
 *     "real" arrays are float / double accordingly;
 *   - tensors of a degree class: (B, 2, D, ..., D) row-major, d bond legs (state.py:24-29);
 *     messages: (2L, D, D) row-major [slot][bra][ket];  index arrays: int32, shape (d, B) leg-major;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success; otherwise bqa_b200_last_error() describes the failure
 *     (thread-local string).  Nothing here falls back to the CPU.
 */
#ifndef BQA_B200_H
#define BQA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BQA_C64 0
#define BQA_C128 1

#define BQA_B200_MAX_BOND_DIM 16
#define BQA_B200_MAX_DEGREE 8

const char* bqa_b200_last_error(void);
int bqa_b200_version(void);
/* number of this library's kernel launches since load (bench.py reports it as gpu_launches) */
long long bqa_b200_launch_count(void);

/* 0 (default): specialised kernels where one exists for (precision, degree, D), generic kernels otherwise;
 * 1: generic kernels only (used by the tests to check the specialised kernels against the generic ones);
 * 2: like 0 with the first-design n = 8 canonicalizer (bqa_fast_canon8.cu) instead of the current one
 *    (bqa_fast_canon8v2.cu): kept for side-by-side measurements.
 * Environment switches read once (side-by-side measurements, results stay within the stated tolerances):
 * BQA_B200_FAST_GRAM=0 (generic node contraction at D = 8), BQA_B200_ROUND_ROBIN=0 (serial Jacobi at n = 16 / 32),
 * BQA_B200_BP_ZIGZAG=0, BQA_B200_FENCE=0|1|2, BQA_B200_MC_SMEM=0|1, BQA_B200_CANON_V=3, BQA_B200_CANON_CONV. */
int bqa_b200_set_kernel_mode(int mode);

/* how long an in-kernel grid barrier or cross-GPU handshake waits before it gives up, sets status[3] (sticky: the engine
 * raises at its next read of the control block) and lets the kernel end; default about 10 s */
int bqa_b200_set_barrier_timeout(double seconds);
/* profiling aid: with a device buffer of 5 * max_iters uint64, CTA 0 of bqa_b200_bp_run writes per sweep the
 * %globaltimer (ns) at: sweep start | its last group done | grid barrier passed | handshake line sent | every peer's
 * line received.  NULL switches it off (default). */
int bqa_b200_set_bp_trace(void* device_buffer);

/* profiling aid: out3[0] = warp-level Jacobi problems (4 matrices each) solved by the n = 8 canonicalizer kernel
 * since load, out3[1] = Jacobi sweeps summed over them, out3[2] = the part of out3[1] spent on the SVD of ker
 * (synchronises the device) */
int bqa_b200_canon_stats(unsigned long long* out3);
/* current n = 8 kernel only: out7[0..2] as above, out7[3] / out7[4] = sweeps summed over the single matrices until each
 * one had converged (eigen phase / SVD phase), out7[5] / out7[6] = matrices: what a warp loses by sweeping until its
 * slowest matrix is done */
int bqa_b200_canon_stats_detail(unsigned long long* out7);
/* profiling aid: out2 = (earliest CTA start, latest CTA end) in %globaltimer ns over the n = 8 launches since the last
 * call (resets; synchronises the device) */
int bqa_b200_canon_span(unsigned long long* out2);

/* bytes of device scratch the node kernels need for a degree class (pass the max over classes) */
size_t bqa_b200_workspace_bytes(int prec, int degree, int D, int D_new);

/* ---- K1 + K2: one BP sweep over a degree class ------------------------------------------------
 * replaces, per sweep and class: batch_slice gathers (state.py:109), Tensor.pass_msgs
 * (backends.py:406-408), assign_at_batch_indices scatters (state.py:111-112), get_dist (state.py:113,
 * backends.py:492-495) and make_inplace_damping_update (state.py:121, backends.py:539-540).
 *
 * reads msgs_cur, writes msgs_nxt[out_pos] = damping * msgs_cur[out_pos] + (1 - damping) * new
 * (the undamped `new` when write_undamped != 0: the reference keeps the undamped last sweep when the
 * iteration cap is hit, state.py:122-123).  resid is a real array [max_iters][2]; sweep `it` folds
 * max |new - old|^2 and max |new + old|^2 into resid[it] with atomic max.  Device-side early exit:
 * sweep it > 0 first tests resid[it-1] (sqrt(num/den) < bp_eps) and, if converged, sets
 * status[0] = 1, status[1] = it and returns without touching the messages -- so the host can enqueue
 * sweeps ahead without a sync per sweep.  resid and status must be zeroed before sweep 0.
 *
 * Messages are Hermitian (BP messages are Hermitian positive semi-definite by construction: state.py:56-57 starts
 * them as diag(lambda) / trace and pass_msgs keeps the property).  The specialised complex64 kernels for degree 3,
 * D = 4 rely on it (3 mode products instead of 6, real diagonals) and write exactly Hermitian outputs; the generic
 * kernels evaluate the reference formula for any input. */
int bqa_b200_bp_sweep(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur,
                      void* msgs_nxt, const int32_t* in_pos, const int32_t* out_pos, double damping,
                      int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU over peer memory (NVLink P2P) -----------------------------------------------------
 * The reference is single-process; these entry points are the B200 scaling path.
 * remote_pos: int32 (d, B) like out_pos; -1 = the message stays local, otherwise (peer << 27 | slot) = where the
 * owner of the receiving node keeps it.  peers: HOST array of BQA_B200_MAX_PEERS device pointers, peers[q] = base
 * of rank q's destination array (msgs_nxt of the same sweep parity / ext) mapped into this process.  The kernel
 * stores such a message twice: locally at out_pos and into the peer's halo slot.  Both NULL = one GPU. */
#define BQA_B200_MAX_PEERS 8
int bqa_b200_bp_sweep_p2p(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur,
                          void* msgs_nxt, const int32_t* in_pos, const int32_t* out_pos, double damping,
                          int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                          void* workspace, size_t workspace_bytes, const int32_t* remote_pos, void* const* peers,
                          void* stream);
int bqa_b200_ext_msgs_p2p(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur,
                          void* ext, const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls,
                          double ztime, void* workspace, size_t workspace_bytes, const int32_t* remote_pos,
                          void* const* peers, void* stream);
/* after the sweep kernels of sweep `it`: atomic-max this rank's resid[it] pair into every peer's resid array
 * (get_dist is a ratio of two GLOBAL maxima, backends.py:492-495) and wait on a flag barrier until every peer
 * has done the same -- their boundary stores have then landed.  it < 0: barrier only.  peer_resid / peer_flags:
 * HOST arrays of `world` device pointers (own arrays at index `rank`); flags: uint32[BQA_B200_MAX_PEERS] per rank,
 * zero-initialised; seq: 1, 2, 3, ... per call.  A peer that never arrives sets status[3] after ~10 s. */
int bqa_b200_sweep_sync(int prec, int rank, int world, void* const* peer_resid, int it, void* const* peer_flags,
                        unsigned seq, int32_t* status, void* stream);
/* Extended messages of the NEXT annealing step, enqueued directly behind bqa_b200_bp_run on the same stream -- before
 * the host has read the run's outcome.  The kernel selects the buffer the run left the messages in from the run's
 * status words (converged: the input of the converging sweep, msgs[(parity + sweeps - 1) % nbuf]; cap reached: the
 * output of the last sweep, msgs[(parity + max_iters) % nbuf]; state.py:118-124), so the GPU works on step k + 1 while
 * the host reads step k's control block.  nbuf = 2 (one GPU, msgs2 = NULL) or 3 (bqa_b200_bp_run across GPUs);
 * parity / max_iters / status: the values given to that bp_run call.  Returns 2 when the shape has no such kernel
 * (the caller then reads the outcome first and calls bqa_b200_ext_msgs).  Replaces: State._get_extended_msgs on
 * the messages _run_bp returned, src/bqa/state.py:127-139 after :97-124. */
int bqa_b200_ext_msgs_after_run(int prec, int degree, int D, long long B, const void* T, const void* msgs0,
                                const void* msgs1, const void* msgs2, int nbuf, int parity, int max_iters,
                                const int32_t* status, void* ext, const int32_t* in_pos, const int32_t* out_pos,
                                const void* edge_ampls, double ztime, const int32_t* remote_pos, void* const* peers,
                                void* stream);

/* The whole BP run of a degree class in ONE cooperative launch (reference _run_bp, state.py:97-124): persistent CTAs
 * iterate sweep -> grid barrier -> residual test on the device.  On return status[0] = converged (0 / 1), status[1] =
 * the reference's sweep count; status[2] (grid-barrier counter) and resid -- (max_iters + 2) x 2 reals, the tail is used
 * as a counter -- must be zero before the call.  Only for graphs whose nodes all sit in ONE degree class that has a
 * specialised kernel; returns 2 (and changes nothing) when there is none -- the caller then enqueues
 * bqa_b200_bp_sweep[_p2p] calls.
 *   world == 1: sweep `it` reads msgs{(parity + it) & 1} and writes the other buffer (msgs2 / peers* unused).
 *   world > 1: THREE message buffers, sweep `it` reads msgs{(parity + it) % 3} and writes the next one; peers0/1/2 =
 *   peer bases of the three buffers; the first `boundary_nodes` nodes of the class are those with a remote out-edge.
 *   Per sweep: boundary groups first; the last CTA to finish them fences and sends a 16-byte DATA line to every peer;
 *   interior groups and the grid barrier run while it is in flight; the RESID line {max |new - old|^2, seq,
 *   max |new + old|^2, seq} follows the barrier and is read by the peers ONE SWEEP LATER (the convergence test lags one
 *   sweep: the third buffer keeps the input of the converging sweep intact, one sweep per run is discarded).  The lines
 *   live 64 bytes into each rank's flag buffer: peer_flags[q] must point to at least 64 + 8 * 16 * BQA_B200_MAX_PEERS
 *   zero-initialised bytes; sequence numbers seq_base + 1 ... seq_base + max_iters + 1 are used.  Nothing of a peer's
 *   control block is written (peer_resid is unused), so no barrier is needed in front of a run. */
int bqa_b200_bp_run(int prec, int degree, int D, long long B, const void* T, void* msgs0, void* msgs1, int parity,
                    const int32_t* in_pos, const int32_t* out_pos, double damping, double bp_eps, int max_iters,
                    void* resid, int32_t* status, const int32_t* remote_pos, void* const* peers0, void* const* peers1,
                    int rank, int world, void* const* peer_resid, void* const* peer_flags, unsigned seq_base,
                    void* msgs2, void* const* peers2, long long boundary_nodes, void* stream);
/* msgs_out[p] = diag(lmbds[p mod L][:D_new]) / trace for every slot p < 2L (state.py:56-57); lmbds: real (L, 2 D_old) */
int bqa_b200_gauge_msgs(int prec, int D_old, int D_new, long long L, const void* lmbds, void* msgs_out, void* stream);

/* ---- K3a: ZZ-extended messages of a degree class -----------------------------------------------
 * replaces _get_extended_msgs (state.py:127-139): pass_msgs with evolution_times, i.e. the open leg is
 * extended D -> 2D by the ZZ half gate (backends.py:519-526) with theta = edge_ampl * ztime.
 * ext: (2L, 2D, 2D).  edge_ampls: real (d, B). */
int bqa_b200_ext_msgs(int prec, int degree, int D, long long B, const void* T, const void* msgs_cur,
                      void* ext, const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls,
                      double ztime, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3b: canonicalizers + lambdas of every undirected edge ------------------------------------
 * replaces _get_canonicalizers (state.py:171-200): masked SVD of every extended message
 * (backends.py:483-490, 709-727), ker = lu_f lu_b^T, masked SVD of ker, canonicalizers = [ul_b vs ; ul_f us],
 * lmbds = s / |s|.  ext, canon: (2L, 2D, 2D); lmbds: real (L, 2D); colmax: real (2D), zeroed by the
 * caller, receives the column-wise max of lmbds over all edges (truncate_lmbds, backends.py:297-299).
 * n_cols: number of leading canonicalizer columns the caller will use (>= min(2D, max_bond_dim), the largest
 * bond dimension the truncation can keep, state.py:233-235); columns >= n_cols of canon may be left unwritten.
 * Kernels: n = 8 in complex64 with n_cols <= D (the bond dimension is capped at D): the Cholesky-factor kernel
 * (bqa_fast_canon8v2.cu); n = 8 in complex64 while the bond dimension may still grow (n_cols > D): the
 * accumulated-rotation kernel (bqa_fast_canon8.cu), whose small singular values are accurate enough for the rank
 * decision at pinv_eps; n = 16 / 32: round-robin Jacobi over the lanes of a warp; else the serial generic routine. */
int bqa_b200_canonicalize(int prec, int D, long long L, const void* ext, void* canon, void* lmbds,
                          void* colmax, double pinv_eps, int n_cols, void* stream);
/* The same with the edges visited in the order `order` (int32 permutation of 0 .. L-1, NULL = identity) and the cost of
 * every edge written to `cost` (one byte per edge, NULL = not wanted): 16 * sweeps of its slower eigenproblem + sweeps of
 * its SVD.  A warp of the n = 8 kernel sweeps until the slowest of its matrices has converged (3.7 / 4.0 sweeps per
 * matrix on average, 5.0 / 5.1 per warp on the 100k benchmark); grouping edges that needed the same number of sweeps
 * last time -- bqa_b200_sort_edges_by_cost: order = edges by descending cost -- removes most of that loss.  Results do
 * not depend on the order (a converged matrix is frozen).  Shapes without the specialised kernel ignore both arrays. */
int bqa_b200_canonicalize_ordered(int prec, int D, long long L, const void* ext, void* canon, void* lmbds, void* colmax,
                                  double pinv_eps, int n_cols, const int32_t* order, void* cost, void* stream);
int bqa_b200_sort_edges_by_cost(long long L, const void* cost, int32_t* order, void* stream);
/* Multi-GPU (no reference counterpart): ONE owner per cut edge instead of both endpoint ranks decomposing it.  This rank
 * decomposes the n_owned edges listed in `owned` (indices into its local edge numbering) and, for those with
 * remote[e] = (peer << 27 | the peer's index of the edge) >= 0, stores C_f, C_b and the lambdas into the peer's arrays as
 * well (peer_canon / peer_lmbds: HOST arrays of peer-mapped bases, peer_L: the ranks' local edge counts), so that every
 * rank ends up with exactly the arrays the duplicate computation gives.  The caller orders the peers' stores before its
 * next kernel (the engine: the max all-reduce of the column maxima that follows on every rank's stream).  Shapes without
 * the specialised n = 8 kernel decompose all L local edges as bqa_b200_canonicalize does. */
int bqa_b200_canonicalize_p2p(int prec, int D, long long L, const void* ext, void* canon, void* lmbds, void* colmax,
                              double pinv_eps, int n_cols, long long n_owned, const int32_t* owned, const int32_t* remote,
                              void* const* peer_canon, void* const* peer_lmbds, const long long* peer_L, void* stream);

/* ---- K3c + K4: apply the simple update to a degree class ---------------------------------------
 * replaces batch_truncate_all_but + apply_canonicalizers_with_extensions (state.py:235-246,
 * backends.py:416-432), _apply_z_layer / _apply_x_layer (state.py:142-156, backends.py:506-510) and
 * _set_to_symmetric_gauge (state.py:219-227, backends.py:450-462) incl. the message re-initialisation
 * msgs[p] = diag(lmbd[p mod L]) / trace (state.py:56-57).
 * T_in: (B, 2, D^d) -> T_out: (B, 2, D_new^d); canon: (2L, 2D, 2D) (first D_new columns are used);
 * lmbds: real (L, 2D) (first D_new entries are used); msgs_out: (2L, D_new, D_new);
 * node_ampls: real (B); edge_ampls: real (d, B). */
int bqa_b200_apply_update(int prec, int degree, int D, int D_new, long long B, const void* T_in,
                          void* T_out, const void* canon, const void* lmbds, void* msgs_out,
                          const int32_t* in_pos, const int32_t* out_pos, const int32_t* lmbd_pos,
                          const void* node_ampls, const void* edge_ampls, double ztime, double xtime,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5a: single-qubit marginals of a degree class ---------------------------------------------
 * replaces get_density_matrices (state.py:77-94, backends.py:440-448) + Bloch conversion (utils.py:23-27).
 * bloch: real (N, 4) = (x, y, z, p0) written at rows node_ids[i]. */
int bqa_b200_density(int prec, int degree, int D, long long B, const void* T, const void* msgs,
                     const int32_t* in_pos, const int32_t* node_ids, void* bloch, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- K5b: sampling helpers (measure, state.py:250-312) -----------------------------------------
 * outcomes: int32 (N), 0 = not measured yet, +1 / -1 = measured (state.py:276).
 * argmax: result[0] = first node id maximising |2 p0 - 1| among unmeasured (state.py:301), result[1] =
 * number of unmeasured nodes; result_p0[0] = its p0.  result: int32[2], result_p0: real[1]. */
int bqa_b200_argmax_unmeasured(int prec, long long N, const void* bloch, const int32_t* outcomes,
                               int32_t* result, void* result_p0, void* stream);
/* project node at class position `pos` onto bit `bit` (zero the other physical slice, renormalise the
 * node; backends.py:729-734 divides the whole class batch instead, which is a pure gauge) */
int bqa_b200_project_node(int prec, int degree, int D, void* T, long long pos, int bit, void* stream);
/* project every unmeasured node of the class with p0 > thr (bit 0) or p0 < 1 - thr (bit 1) and record
 * the outcome (state.py:286-296); n_projected (int32[1]) is incremented by the number of projections */
int bqa_b200_threshold_project(int prec, int degree, int D, long long B, void* T, const int32_t* node_ids,
                               const void* bloch, int32_t* outcomes, double thr, int32_t* n_projected,
                               void* stream);

/* ---- all degree classes of a graph in one launch (bqa_multiclass.cuh) -----------------------------------------
 * The reference loops over the degree classes on the host (state.py:106-112, :127-139, :238-246).  These entry points
 * take the classes as a table (host array, copied into the kernel parameters): a step on a graph with several degree
 * classes is ext_msgs_classes + canonicalize + apply_update_classes + bp_run_classes = 4 launches, and the whole BP run
 * of _run_bp (state.py:97-124) is one cooperative launch with the convergence test on the device (status[0] =
 * converged, status[1] = sweeps executed, status[2] = barrier counter -- zeroed by the caller --, status[3] = abort).
 * Results equal the per-class entry points bit for bit.  Single GPU only (no halo stores). */
typedef struct bqa_b200_class {
  int degree;
  long long B;
  const void* T_in;            /* (B, 2, D, ..., D) */
  void* T_out;                 /* apply_update_classes only */
  const int32_t* in_pos;       /* (degree, B) */
  const int32_t* out_pos;
  const int32_t* lmbd_pos;     /* apply_update_classes only */
  const void* node_ampls;      /* (B) reals, apply_update_classes only */
  const void* edge_ampls;      /* (degree, B) reals */
} bqa_b200_class;
int bqa_b200_ext_msgs_classes(int prec, int n_classes, const bqa_b200_class* classes_host, int D, const void* msgs_cur,
                              void* ext, double ztime, void* workspace, size_t workspace_bytes, void* stream);
int bqa_b200_apply_update_classes(int prec, int n_classes, const bqa_b200_class* classes_host, int D, int D_new,
                                  const void* canon, const void* lmbds, void* msgs_out, double ztime, double xtime,
                                  void* workspace, size_t workspace_bytes, void* stream);
int bqa_b200_bp_run_classes(int prec, int n_classes, const bqa_b200_class* classes_host, int D, void* msgs0, void* msgs1,
                            int parity, double damping, double bp_eps, int max_iters, void* resid, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- raw operations of the backend interface on device arrays (bqa_tensor_ops.cu) ----------------------------
 * The reference's plugin boundary is the ABC bqa.backends.Tensor (src/bqa/backends.py:28-252): 36 abstract raw
 * operations from which it builds every composite.  bqa_b200.tensor_backend.B200Backend implements them with the
 * entry points below (dense row-major device arrays; complex64 / complex128 by `prec`; index arrays int64), so the
 * unmodified engine src/bqa/state.py runs op by op on the GPU.  Each comment names the raw op it serves.           */

/* op: 0 inv_raw (:638), 1 pinv_raw (:719-727; cut at machine epsilon of the precision), 2 sqrt_raw (:683, principal
 * branch), 3 sin_raw, 4 cos_raw (:741-746), 5 conj_raw (:663) */
int bqa_b200_t_unary(int prec, int op, long long n, const void* a, void* out, void* stream);
/* op: 0 mul_raw, 1 sum_raw, 2 sub_raw, 3 div_raw (:694-706, :757) with numpy broadcasting expressed as element
 * strides (0 on a broadcast axis); out is dense, rank <= 8 */
int bqa_b200_t_binary(int prec, int op, int rank, const long long* shape, const long long* strides_a,
                      const long long* strides_b, const void* a, const void* b, void* out, void* stream);
/* strided copy of 4-, 8- or 16-byte elements: transpose_raw (:676), truncate_raw_tensor (:737), take_batch_slice (:688),
 * concatenate (:749), apply_x_to_phys_dim_raw (:632, a negative stride on the physical axis) */
int bqa_b200_t_copy(int elem_bytes, int rank, const long long* shape, const long long* strides_in,
                    const long long* strides_out, const void* in, void* out, void* stream);
/* batched_gather (:607, scatter = 0: out[i] = in[idx[i]]) and assign_at_batch_indices_raw (:646-650, scatter = 1:
 * out[idx[i]] = in[i], in place on `out`); rows of row_elems elements */
int bqa_b200_t_rows(int elem_bytes, int scatter, long long n_idx, long long row_elems, const long long* idx,
                    const void* in, void* out, void* stream);
int bqa_b200_t_fill(int prec, long long n, void* out, double re, double im, void* stream);
/* make_inplace_damping_update_raw (:761-764): dst = alpha dst + beta src */
int bqa_b200_t_axpby(int prec, long long n, void* dst, const void* src, double alpha, double beta, void* stream);
/* max_norm (:621-625): is_full -> one complex element holding max |a|; otherwise |max over the batch axis| per column */
int bqa_b200_t_max_abs(int prec, long long n, const void* a, void* out1, void* stream);
int bqa_b200_t_col_max(int prec, long long batch, long long inner, const void* a, void* out, void* stream);
/* mode 0: batched_l2_norm (:610-618); mode 1: batched_trace (:628) of (batch, n, n) */
int bqa_b200_t_batch_reduce(int prec, int mode, long long batch, long long inner, int n, const void* a, void* out,
                            void* stream);
/* batched_diag (:657-660): out[..., i, j] = a[..., i] (i == j) */
int bqa_b200_t_diag(int prec, long long rows, int n, const void* a, void* out, void* stream);
/* batched_matmul (:666-670): (batch, m, k) @ (batch, k, n) */
int bqa_b200_t_matmul(int prec, long long batch, int m, int k, int n, const void* a, const void* b, void* out,
                      void* stream);
/* batched_svd (:709-717) of square matrices n <= 32: u, s (stored complex), vh, descending, masked at pinv_eps */
size_t bqa_b200_t_svd_scratch_bytes(int prec, int n);
int bqa_b200_t_svd(int prec, long long batch, int n, const void* a, void* u, void* s, void* vh, double pinv_eps,
                   void* scratch, size_t scratch_bytes, void* stream);
/* (B, 2, 2) density matrices from the (x, y, z, p0) rows bqa_b200_density writes (utils.py:23-27 inverted) */
int bqa_b200_t_bloch_to_rho(int prec, long long B, const void* bloch, void* rho, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BQA_B200_H */
