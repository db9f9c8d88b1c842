// bqa_multiclass.cuh -- the node kernels of bqa_generic.cuh over ALL degree classes of a graph in one launch.
//
// The reference loops over the degree classes in Python (state.py:106-112, :127-139, :238-246) and so did the per-class
// entry points: on the small graphs of BASELINE configs 1-3 (127-1000 qubits in 3-4 classes) a step was 3 x classes + 1
// launches plus one launch per class and BP sweep.  Here the classes are rows of a table passed in the kernel
// parameters, a warp takes items of the concatenated node list, and the whole BP run (_run_bp, state.py:97-124) is ONE
// cooperative launch that loops sweep -> grid barrier -> residual test on the device, like k_bp_run_d3D4 does for the
// headline shape.  Per-node arithmetic is the same code (bqa_core.cuh), so results equal the per-class launches bit for
// bit (tests/test_gpu_parity.py::test_multiclass_launches_equal_per_class_launches).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "../../include/bqa_b200.h"
#include "bqa_fast_gram_d3D8.cuh"
#include "bqa_generic.cuh"

namespace bqa {

#define BQA_MAX_CLASSES (BQA_MAX_DEGREE + 1)

template <typename R>
struct ClassRow {
  int d;
  long long B, first;              // nodes of the class; index of its first node in the concatenated list
  const cx<R>* T;
  cx<R>* Tout;
  const int32_t *in_pos, *out_pos, *lmbd_pos;
  const R *node_ampls, *edge_ampls;
};

template <typename R>
struct MultiArgs {
  int n_classes, D, Dn;
  long long total;                 // nodes over all classes
  ClassRow<R> c[BQA_MAX_CLASSES];
  cx<R>* msgs[2];                  // BP run: ping-pong buffers; ext / apply: msgs[0] = input messages, msgs[1] = output
  int parity, max_iters;
  const cx<R>* canon;
  const R* lmbds;
  R damping, bp_eps, ztime, xtime;
  R* resid;
  int32_t* status;
  cx<R>* ws;
  size_t ws_per_warp;
  int ws_in_smem;                  // the per-warp scratch fits into shared memory (small D): no global-memory round trips
  int fast_gram_off;               // FAST kernels: byte offset of the gram8 scratch in the CTA's dynamic shared memory
  long long timeout_cycles;
};

// scratch of this warp: a slab of the CTA's dynamic shared memory when it fits (ws_in_smem), else of the global workspace
template <typename R>
__device__ __forceinline__ cx<R>* warp_scratch(const MultiArgs<R>& a, long long warp) {
  extern __shared__ __align__(16) unsigned char mc_smem[];
  return a.ws_in_smem ? reinterpret_cast<cx<R>*>(mc_smem) + (size_t)(threadIdx.x >> 5) * a.ws_per_warp
                      : a.ws + (size_t)warp * a.ws_per_warp;
}

// Gram parts of one node.  MODE 2 (launched only for D = 8 in complex64): the shared-memory routine for degree 3, the
// generic one for the other classes; MODE 1 (D = 2, 4, 8): the generic routine with the bond dimension as a compile-time
// constant; MODE 0 (any other D): the runtime-D routine alone, so that e.g. D = 16 keeps the 72 registers per thread and the
// occupancy it had before the unrolled variants existed (measured: BP run at D = 16 544 ms against 625 ms)
template <typename R, int MODE>
__device__ __forceinline__ void mc_node_gram(const MultiArgs<R>& a, int d, int D, const cx<R>* T, const cx<R>* const* mp,
                                             cx<R>* P, cx<R>* E, cx<R>* gram) {
  if constexpr (MODE == 2) {
    if (d == 3) {
      extern __shared__ __align__(16) unsigned char mc_smem[];
      float2* sm = reinterpret_cast<float2*>(mc_smem + a.fast_gram_off) + (size_t)(threadIdx.x >> 5) * gram8::kWarpElems;
      gram8::node_gram_d3D8(reinterpret_cast<const float2*>(T), reinterpret_cast<const float2*>(mp[0]),
                            reinterpret_cast<const float2*>(mp[1]), reinterpret_cast<const float2*>(mp[2]),
                            reinterpret_cast<float2*>(gram), sm);
      return;
    }
  }
  GroupWarp g;
  if constexpr (MODE == 0) node_gram_impl<R, GroupWarp, 0>(g, d, D, T, mp, P, E, gram);     // any D: fewest registers
  else node_gram<R>(g, d, D, T, mp, P, E, gram);                                           // D = 2, 4, 8 unrolled
}

template <typename R>
__device__ __forceinline__ int find_class(const MultiArgs<R>& a, long long item) {
  int k = 0;
  while (k + 1 < a.n_classes && item >= a.c[k + 1].first) ++k;
  return k;
}

// one BP sweep over every class: reads `cur`, writes `out`; folds the residual maxima of sweep `it` into a.resid
template <typename R, int MODE>
__device__ __forceinline__ void mc_sweep(const MultiArgs<R>& a, const cx<R>* cur, cx<R>* out, int it, int write_undamped) {
  GroupWarp g;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int D = a.D, DD = D * D;
  R mnum = 0, mden = 0;
  for (long long item = warp; item < a.total; item += nwarps) {
    const ClassRow<R>& c = a.c[find_class(a, item)];
    const long long node = item - c.first;
    const int d = c.d, W = 2 * ipow(D, d);
    cx<R>* P = warp_scratch<R>(a, warp);
    cx<R>* E = P + W;
    cx<R>* gram = E + W;
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = cur + (size_t)c.in_pos[(size_t)j * c.B + node] * DD;
    mc_node_gram<R, MODE>(a, d, D, c.T + (size_t)node * W, mp, P, E, gram);
    for (int k = 0; k < d; ++k) {
      const cx<R>* g0 = gram + (size_t)k * 2 * DD;
      const size_t slot = (size_t)c.out_pos[(size_t)k * c.B + node];
      emit_bp_msg<R>(g, D, g0, g0 + DD, cur + slot * DD, out + slot * DD, a.damping, write_undamped, mnum, mden);
    }
    g.sync();
  }
  mnum = warp_max(mnum);
  mden = warp_max(mden);
  if (lane == 0) {
    atomic_max_nonneg(a.resid + 2 * it, mnum);
    atomic_max_nonneg(a.resid + 2 * it + 1, mden);
  }
}

__device__ __forceinline__ unsigned mc_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all CTAs of a cooperative grid; `counter` only grows (zeroed by the host before the launch); status[3] != 0 aborts
__device__ __forceinline__ bool mc_grid_barrier(unsigned* counter, unsigned& generation, volatile int32_t* status,
                                                long long timeout_cycles) {
  __syncthreads();
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    __threadfence();
    ++generation;
    atomicAdd(counter, 1u);
    const unsigned target = generation * gridDim.x;
    const long long t0 = clock64();
    while (mc_ld_acquire(counter) < target) {
      if (status[3] != 0 || clock64() - t0 > timeout_cycles) { status[3] = 1; ok = 0; break; }
    }
  }
  __syncthreads();
  return ok != 0;
}

// the whole BP run (reference _run_bp, state.py:97-124): status[0] = converged, status[1] = sweeps executed
template <typename R, int MODE>
__global__ void __launch_bounds__(128) k_mc_bp_run(const __grid_constant__ MultiArgs<R> a) {
  unsigned* counter = reinterpret_cast<unsigned*>(a.status + 2);
  unsigned generation = 0;
  int sweeps = a.max_iters, converged = 0;
  for (int it = 0; it < a.max_iters; ++it) {
    const int cur = (a.parity + it) & 1;
    mc_sweep<R, MODE>(a, a.msgs[cur], a.msgs[cur ^ 1], it, it == a.max_iters - 1);    // cap: the undamped sweep is kept (:122-123)
    if (!mc_grid_barrier(counter, generation, a.status, a.timeout_cycles)) return;
    const R num = __ldcg(a.resid + 2 * it), den = __ldcg(a.resid + 2 * it + 1);
    if (msqrt(num / den) < a.bp_eps) { sweeps = it + 1; converged = 1; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.status[1] = sweeps; a.status[0] = converged; }
}

// ZZ-extended messages of every class (_get_extended_msgs, state.py:127-139): msgs[0] -> ext in msgs[1]
template <typename R, int MODE>
__global__ void __launch_bounds__(128) k_mc_ext(const __grid_constant__ MultiArgs<R> a) {
  GroupWarp g;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int D = a.D, DD = D * D;
  for (long long item = warp; item < a.total; item += nwarps) {
    const ClassRow<R>& c = a.c[find_class(a, item)];
    const long long node = item - c.first;
    const int d = c.d, W = 2 * ipow(D, d);
    cx<R>* P = warp_scratch<R>(a, warp);
    cx<R>* E = P + W;
    cx<R>* gram = E + W;
    const cx<R>* mp[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) mp[j] = a.msgs[0] + (size_t)c.in_pos[(size_t)j * c.B + node] * DD;
    mc_node_gram<R, MODE>(a, d, D, c.T + (size_t)node * W, mp, P, E, gram);
    for (int k = 0; k < d; ++k) {
      const cx<R>* g0 = gram + (size_t)k * 2 * DD;
      const size_t slot = (size_t)c.out_pos[(size_t)k * c.B + node];
      emit_ext_msg<R>(g, D, g0, g0 + DD, c.edge_ampls[(size_t)k * c.B + node] * a.ztime, a.msgs[1] + slot * 4 * DD);
    }
    g.sync();
  }
}

// truncated simple update + Rz / Rx + symmetric gauge + message re-initialisation of every class
// (state.py:235-246, :142-156, :219-227, :56-57): T -> Tout, msgs[1][out_pos] = diag(lambda) / trace
template <typename R>
__global__ void __launch_bounds__(128) k_mc_apply(const __grid_constant__ MultiArgs<R> a) {
  GroupWarp g;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int D = a.D, Dn = a.Dn, n = 2 * D;
  for (long long item = warp; item < a.total; item += nwarps) {
    const ClassRow<R>& c = a.c[find_class(a, item)];
    const long long node = item - c.first;
    const int d = c.d;
    const int Win = 2 * ipow(D, d), Wout = 2 * ipow(Dn, d), Wmax = 2 * ipow(D > Dn ? D : Dn, d);
    cx<R>* bufA = warp_scratch<R>(a, warp);
    cx<R>* bufB = bufA + Wmax;
    cx<R>* wbuf = bufB + Wmax;
    const cx<R>* cp[BQA_MAX_DEGREE];
    const R* lp[BQA_MAX_DEGREE];
    R th[BQA_MAX_DEGREE];
    for (int j = 0; j < d; ++j) {
      cp[j] = a.canon + (size_t)c.in_pos[(size_t)j * c.B + node] * n * n;
      lp[j] = a.lmbds + (size_t)c.lmbd_pos[(size_t)j * c.B + node] * n;
      th[j] = c.edge_ampls[(size_t)j * c.B + node] * a.ztime;
    }
    node_apply_update<R>(g, d, D, Dn, c.T + (size_t)node * Win, cp, th, lp, c.node_ampls[node] * a.ztime, a.xtime, bufA,
                         bufB, wbuf, c.Tout + (size_t)node * Wout);
    for (int j = 0; j < d; ++j)
      emit_gauge_msg<R>(g, Dn, lp[j], a.msgs[1] + (size_t)c.out_pos[(size_t)j * c.B + node] * Dn * Dn);
  }
}

// ---- launchers ---------------------------------------------------------------------------------------------------
// kind: 0 = extended messages, 1 = apply update, 2 = BP run
template <typename R>
int launch_multiclass(int kind, int n_classes, const bqa_b200_class* cls, int D, int Dn, void* msgs0, void* msgs1,
                      int parity, const void* canon, const void* lmbds, double ztime, double xtime, double damping,
                      double bp_eps, int max_iters, void* resid, int32_t* status, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool allow_fast) {
  if (n_classes < 1 || n_classes > BQA_MAX_CLASSES) return set_error("%d degree classes outside [1, %d]", n_classes, BQA_MAX_CLASSES);
  MultiArgs<R> a{};
  a.n_classes = 0; a.D = D; a.Dn = Dn;
  size_t per_warp = 16;
  long long total = 0;
  for (int k = 0; k < n_classes; ++k) {
    const bqa_b200_class& s = cls[k];
    if (s.degree < 0 || s.degree > BQA_MAX_DEGREE) return set_error("degree %d outside [0, %d]", s.degree, BQA_MAX_DEGREE);
    if (s.B <= 0 || (kind != 1 && s.degree == 0)) continue;          // isolated qubits send and receive no messages
    ClassRow<R>& r = a.c[a.n_classes++];
    r.d = s.degree; r.B = s.B; r.first = total;
    r.T = (const cx<R>*)s.T_in; r.Tout = (cx<R>*)s.T_out;
    r.in_pos = s.in_pos; r.out_pos = s.out_pos; r.lmbd_pos = s.lmbd_pos;
    r.node_ampls = (const R*)s.node_ampls; r.edge_ampls = (const R*)s.edge_ampls;
    total += s.B;
    const size_t need = generic_ws_elems_per_warp(s.degree, D, Dn);
    if (need > per_warp) per_warp = need;
  }
  if (total == 0) return 0;
  a.total = total;
  a.msgs[0] = (cx<R>*)msgs0; a.msgs[1] = (cx<R>*)msgs1; a.parity = parity & 1; a.max_iters = max_iters;
  a.canon = (const cx<R>*)canon; a.lmbds = (const R*)lmbds;
  a.ztime = (R)ztime; a.xtime = (R)xtime; a.damping = (R)damping; a.bp_eps = (R)bp_eps;
  a.resid = (R*)resid; a.status = status; a.ws = (cx<R>*)ws; a.ws_per_warp = per_warp;
  a.timeout_cycles = fast::barrier_timeout_cycles();
  long long blocks = (total + 3) / 4;
  const long long cap = (long long)BQA_GENERIC_MAX_WARPS / 4;
  if (blocks > cap) blocks = cap;
  // per-warp scratch in shared memory only where it measured faster (r2, profiles/r2_small_configs.jsonl): a footprint the
  // L1 does not hold (>= 18 KB per warp: D = 8 in complex64) on a graph small enough that the CTAs it needs are resident
  // anyway -- heavy-hex 127 at D <= 8: 436 -> 459 steps/s.  Elsewhere the global workspace wins: smaller footprints stay
  // in L1 (grid 20x20, D <= 4: 1258 vs 1216 steps/s), and on a large graph the per-node code is latency bound and the
  // warps lost to the shared-memory footprint cost more than the L2 round trips (20k nodes at D = 8: BP run 29.5 ms
  // against 37.4 ms with 78 KB per CTA).  BQA_B200_MC_SMEM=0/1 forces the choice.
  size_t smem = per_warp * sizeof(cx<R>) * 4;
  static const int smem_mode = [] { const char* e = getenv("BQA_B200_MC_SMEM"); return e ? atoi(e) : -1; }();
  {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long fit = smem ? (long long)((200 * 1024) / smem) * sms : 0;     // CTAs resident with this footprint
    a.ws_in_smem = smem_mode >= 0 ? (smem_mode != 0 && smem <= 100 * 1024) : (smem >= 72 * 1024 && smem <= 100 * 1024 && blocks <= fit);
  }
  if (!a.ws_in_smem) smem = 0;
  // degree 3 at D = 8 in complex64: the node contraction out of shared memory (bqa_fast_gram_d3D8.cuh); P / E of the other
  // classes and the Gram parts stay in the global workspace
  bool fast = false;
  if constexpr (std::is_same<R, float>::value) {
    if (allow_fast && kind != 1 && D == 8)
      for (int k = 0; k < a.n_classes; ++k) fast = fast || a.c[k].d == 3;
  }
  if (fast) {
    a.ws_in_smem = 0;
    a.fast_gram_off = 0;
    smem = (size_t)4 * gram8::kWarpBytes;
  }
  const void* fn = nullptr;
  if constexpr (std::is_same<R, float>::value) {
    if (fast) fn = kind == 0 ? (const void*)k_mc_ext<R, 2> : (const void*)k_mc_bp_run<R, 2>;
  }
  const bool unrolled = D == 2 || D == 4 || D == 8;
  if (!fn && kind == 1) fn = (const void*)k_mc_apply<R>;
  if (!fn && unrolled) fn = kind == 0 ? (const void*)k_mc_ext<R, 1> : (const void*)k_mc_bp_run<R, 1>;
  if (!fn) fn = kind == 0 ? (const void*)k_mc_ext<R, 0> : (const void*)k_mc_bp_run<R, 0>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(multiclass): %s", cudaGetErrorString(e));
  }
  if (kind == 2) {                                                    // every CTA must be resident: cooperative launch
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 128, smem);
    if (occ < 1 || sms < 1) return set_error("bp_run_classes: the kernel cannot be made resident");
    if (blocks > (long long)occ * sms) blocks = (long long)occ * sms;
  }
  if (!a.ws_in_smem && ws_bytes < per_warp * sizeof(cx<R>) * (size_t)blocks * 4)
    return set_error("workspace too small: need %zu bytes, got %zu", per_warp * sizeof(cx<R>) * (size_t)blocks * 4, ws_bytes);
  void* params[] = {&a};
  if (kind == 2) {
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(128), params, smem, st);
    if (e != cudaSuccess) return set_error("cudaLaunchCooperativeKernel(k_mc_bp_run): %s", cudaGetErrorString(e));
    return after_launch("bp_run_classes");
  }
  cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(128), params, smem, st);
  if (e != cudaSuccess) return set_error("cudaLaunchKernel(multiclass): %s", cudaGetErrorString(e));
  return after_launch(kind == 0 ? "ext_msgs_classes" : "apply_update_classes");
}

#define BQA_INSTANTIATE_MULTICLASS(R)                                                                                \
  template int launch_multiclass<R>(int, int, const bqa_b200_class*, int, int, void*, void*, int, const void*,        \
                                    const void*, double, double, double, double, int, void*, int32_t*, void*, size_t, \
                                    cudaStream_t, bool);

}  // namespace bqa
