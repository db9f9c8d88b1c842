// bqa_sync.cu -- the two small kernels of the peer-memory (NVLink P2P) multi-GPU path.
//
//   k_gauge_msgs  : msgs[p] = diag(lambda[p mod L][:Dn]) / trace for EVERY slot (reference state.py:56-57).  On one
//                   GPU the apply-update kernel writes these per node; a partitioned rank also needs the halo
//                   slots, whose lambdas it holds (cut edges are canonicalised on both owners), so it fills all
//                   slots locally instead of exchanging them.
//   k_sweep_sync  : after a BP sweep whose kernels stored the boundary messages straight into the peers' halo
//                   slots: push this rank's residual maxima to every peer (system-scope atomic max over NVLink:
//                   get_dist is a ratio of two GLOBAL maxima, backends.py:492-495) and run a flag barrier, so the
//                   next sweep starts only when every peer's stores and maxima have landed.  it < 0: barrier only.
#include <cuda_runtime.h>

#include "bqa_core.cuh"
#include "bqa_launch.cuh"

namespace bqa {

template <typename R>
__global__ void __launch_bounds__(256) k_gauge_msgs(int stride, int Dn, long long n_slots, long long L, const R* lmbds,
                                                    cx<R>* msgs) {
  const int DD = Dn * Dn;
  const long long total = n_slots * DD;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / DD;
    const int o = (int)(i - p * DD), r = o / Dn, c = o - r * Dn;
    const R* lam = lmbds + (size_t)(p % L) * stride;
    R v = 0;
    if (r == c) {
      R tr = 0;
      for (int k = 0; k < Dn; ++k) tr += lam[k];
      v = lam[c] * (R(1) / tr);                             // same arithmetic as emit_gauge_msg (bqa_core.cuh)
    }
    msgs[i] = mk<R>(v, R(0));
  }
}

// Dn = 4, complex64 (the headline shape): one thread per message ROW -- the 4 lambdas of the row's edge in one 16-byte load,
// the row (4 complex) in two 16-byte stores, shifts instead of the 64-bit divisions of the generic index arithmetic.
// `pos_of(q)` = (slot, lambda row) of message q.  Same arithmetic as emit_gauge_msg (bqa_core.cuh).
template <class PosOf>
__global__ void __launch_bounds__(256) k_gauge_rows4(int stride, long long n, PosOf pos_of, const float* __restrict__ lmbds,
                                                     float4* __restrict__ msgs) {
  const long long total = n * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i >> 2;
    const int r = (int)(i & 3);
    long long slot, row;
    pos_of(q, slot, row);
    const float4 lam = __ldg(reinterpret_cast<const float4*>(lmbds + (size_t)row * stride));
    float tr = 0.f;
    tr += lam.x; tr += lam.y; tr += lam.z; tr += lam.w;
    const float inv = 1.f / tr;
    const float v = (r == 0 ? lam.x : r == 1 ? lam.y : r == 2 ? lam.z : lam.w) * inv;
    float4* dst = msgs + (size_t)slot * 8 + r * 2;          // a message is 8 float4, a row 2
    dst[0] = make_float4(r == 0 ? v : 0.f, 0.f, r == 1 ? v : 0.f, 0.f);
    dst[1] = make_float4(r == 2 ? v : 0.f, 0.f, r == 3 ? v : 0.f, 0.f);
  }
}
struct PosFromArrays {
  const int32_t* out_pos;
  const int32_t* lmbd_pos;
  __device__ __forceinline__ void operator()(long long q, long long& slot, long long& row) const {
    slot = __ldg(out_pos + q);
    row = __ldg(lmbd_pos + q);
  }
};
struct PosAllSlots {
  long long L;
  __device__ __forceinline__ void operator()(long long q, long long& slot, long long& row) const {
    slot = q;
    row = q >= L ? q - L : q;
  }
};
static bool rows4_ok(int D_old, int Dn, const void* lmbds, const void* msgs) {
  return Dn == 4 && (2 * D_old) % 4 == 0 && ((uintptr_t)lmbds & 15) == 0 && ((uintptr_t)msgs & 15) == 0;
}

template <typename R>
int launch_gauge_msgs(int D_old, int Dn, long long L, const void* lmbds, void* msgs_out, cudaStream_t st) {
  if (L == 0) return 0;
  if (sizeof(R) == 4 && rows4_ok(D_old, Dn, lmbds, msgs_out)) {
    long long blocks = (2 * L * 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_gauge_rows4<<<(int)blocks, 256, 0, st>>>(2 * D_old, 2 * L, PosAllSlots{L}, (const float*)lmbds, (float4*)msgs_out);
    return after_launch("gauge_msgs(rows4)");
  }
  const long long total = 2 * L * Dn * Dn;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_gauge_msgs<R><<<(int)blocks, 256, 0, st>>>(2 * D_old, Dn, 2 * L, L, (const R*)lmbds, (cx<R>*)msgs_out);
  return after_launch("gauge_msgs");
}
template int launch_gauge_msgs<float>(int, int, long long, const void*, void*, cudaStream_t);
template int launch_gauge_msgs<double>(int, int, long long, const void*, void*, cudaStream_t);

template <typename R>
__global__ void __launch_bounds__(256) k_gauge_slots(int stride, int Dn, long long n, const int32_t* out_pos,
                                                     const int32_t* lmbd_pos, const R* lmbds, cx<R>* msgs) {
  const int DD = Dn * Dn;
  const long long total = n * DD;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / DD;
    const int o = (int)(i - q * DD), r = o / Dn, c = o - r * Dn;
    const R* lam = lmbds + (size_t)lmbd_pos[q] * stride;
    R v = 0;
    if (r == c) {
      R tr = 0;
      for (int k = 0; k < Dn; ++k) tr += lam[k];
      v = lam[c] * (R(1) / tr);
    }
    msgs[(size_t)out_pos[q] * DD + o] = mk<R>(v, R(0));
  }
}

template <typename R>
int launch_gauge_slots(int D_old, int Dn, long long n, const int32_t* out_pos, const int32_t* lmbd_pos, const void* lmbds,
                       void* msgs_out, cudaStream_t st) {
  if (n == 0) return 0;
  if (sizeof(R) == 4 && rows4_ok(D_old, Dn, lmbds, msgs_out)) {
    long long blocks = (n * 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_gauge_rows4<<<(int)blocks, 256, 0, st>>>(2 * D_old, n, PosFromArrays{out_pos, lmbd_pos}, (const float*)lmbds,
                                               (float4*)msgs_out);
    return after_launch("gauge_slots(rows4)");
  }
  long long blocks = (n * Dn * Dn + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_gauge_slots<R><<<(int)blocks, 256, 0, st>>>(2 * D_old, Dn, n, out_pos, lmbd_pos, (const R*)lmbds, (cx<R>*)msgs_out);
  return after_launch("gauge_slots");
}
template int launch_gauge_slots<float>(int, int, long long, const int32_t*, const int32_t*, const void*, void*, cudaStream_t);
template int launch_gauge_slots<double>(int, int, long long, const int32_t*, const int32_t*, const void*, void*, cudaStream_t);

struct SyncArgs {
  int rank, world, it, dbl;
  unsigned seq;
  long long timeout_cycles;
  void* resid[BQA_MAX_PEERS];              // base of every rank's residual array (peer mapped)
  unsigned* flags[BQA_MAX_PEERS];          // every rank's flag array: flags[q][src] is written by rank src
  int32_t* status;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32) k_sweep_sync(SyncArgs a) {
  const int q = threadIdx.x;
  if (q < a.world && q != a.rank) {
    if (a.it >= 0) {                                        // non-negative reals order like their bit patterns
      if (a.dbl) {
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(a.resid[a.rank]) + 2 * a.it;
        unsigned long long* theirs = reinterpret_cast<unsigned long long*>(a.resid[q]) + 2 * a.it;
        atomicMax_system(theirs, mine[0]);
        atomicMax_system(theirs + 1, mine[1]);
      } else {
        const unsigned* mine = reinterpret_cast<const unsigned*>(a.resid[a.rank]) + 2 * a.it;
        unsigned* theirs = reinterpret_cast<unsigned*>(a.resid[q]) + 2 * a.it;
        atomicMax_system(theirs, mine[0]);
        atomicMax_system(theirs + 1, mine[1]);
      }
    }
    __threadfence_system();
    st_release_sys(a.flags[q] + a.rank, a.seq);             // "rank has finished sweep seq" on peer q
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(a.flags[a.rank] + q) - a.seq) < 0) {
      if (clock64() - t0 > a.timeout_cycles) {              // a peer died: flag the error instead of hanging
        a.status[3] = 1;
        break;
      }
    }
  }
}

int launch_sweep_sync(int prec, int rank, int world, void* const* peer_resid, int it, void* const* peer_flags,
                      unsigned seq, int32_t* status, cudaStream_t st) {
  if (world < 1 || world > BQA_MAX_PEERS || rank < 0 || rank >= world)
    return set_error("sweep_sync: rank %d / world %d outside [1, %d]", rank, world, BQA_MAX_PEERS);
  if (world == 1) return 0;
  SyncArgs a{};
  a.rank = rank; a.world = world; a.it = it; a.dbl = prec == 1; a.seq = seq; a.status = status;
  a.timeout_cycles = fast::barrier_timeout_cycles();
  for (int q = 0; q < world; ++q) { a.resid[q] = peer_resid ? peer_resid[q] : nullptr; a.flags[q] = (unsigned*)peer_flags[q]; }
  if (it >= 0 && !peer_resid) return set_error("sweep_sync: residual arrays missing");
  k_sweep_sync<<<1, 32, 0, st>>>(a);
  return after_launch("sweep_sync");
}

}  // namespace bqa
