// bqa_fast_apply.cu -- simple-update application for the headline shape (degree 3, D = 4 -> D_new = 4, complex64).
//
// replaces, per degree class: batch_truncate_all_but + apply_canonicalizers_with_extensions (src/bqa/state.py:235-246,
// backends.py:416-432), _apply_z_layer / _apply_x_layer (state.py:142-156, backends.py:506-510) and the tensor part
// of _set_to_symmetric_gauge (state.py:219-227, backends.py:450-462):
//     T'[p, c0, c1, c2] = sum_a T[p, a0, a1, a2] prod_j W_j^(p)[a_j, c_j],
//     W_j^(p)[a, c] = (f0_j C_j[a, c] + (-1)^p f1_j C_j[4 + a, c]) sqrt(lambda_j[c])          (c < 4)
//     out_p = cos(xt) e^{-+ i phi} T'_p - i sin(xt) e^{+- i phi} T'_{1-p},   T_out = out / |out|_2
// (the doubled legs of the reference are never materialised).  The re-initialised messages diag(lambda)/trace are
// written by bqa_b200_gauge_msgs.
//
// Same decomposition as the BP kernel (bqa_fast_d3D4.cu): 8 lanes per node, lane (p, a0) holds T[p, a0, :, :]; a warp
// works on 4 consecutive nodes with a two-stage cp.async pipeline that streams the 4 KB of node tensors, the 24
// useful 32-byte row pieces of each gathered canonicalizer and the lambdas of the NEXT group while the current one
// is contracted.  Legs 1 and 2 are thread-local; leg 0 and the Rx mixing of p = 0, 1 go through shared memory; the
// result leaves through shared memory as fully coalesced 16-byte stores.
#include <cuda_runtime.h>

#include "bqa_core.cuh"
#include "bqa_fast_common.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace fast_apply {

using namespace fast;

#ifndef BQA_APPLY_WARPS
#define BQA_APPLY_WARPS 12
#endif
constexpr int kWarps = BQA_APPLY_WARPS;
constexpr int kThreads = kWarps * 32;
constexpr int kSlice = 144;                       // 128-byte slice + 16 bytes of padding
constexpr int kTBytes = 32 * kSlice;              // 4 nodes x 8 slices
constexpr int kCBlk = 8 * 32 + 16;                // 8 row pieces (rows a and 4 + a, columns 0..3) of one canonicalizer
constexpr int kCBytes = 12 * kCBlk;               // 3 legs x 4 nodes
constexpr int kLBytes = 12 * 16;                  // 4 lambdas per (leg, node)
constexpr int kStage = kTBytes + kCBytes + kLBytes;
constexpr int kWBytes = 4 * 3 * 2 * 128;          // W_j^(p) of the 4 nodes
constexpr int kWarpBytes = 2 * kStage + kWBytes;
constexpr int kSmem = kWarps * kWarpBytes;
static_assert(kSmem <= 227 * 1024, "BQA_APPLY_WARPS x (two stages + W tiles) exceeds the 227 KB of shared memory a CTA can own");

struct Args {
  long long B;
  const float2* T;
  float2* Tout;
  const float2* canon;
  const float* lmbds;          // (L, 8)
  const int32_t* in_pos;
  const int32_t* lmbd_pos;
  const float* node_ampls;
  const float* edge_ampls;
  float ztime, xtime;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_s(unsigned dst, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ int ldg_ordered(const int32_t* ptr) {     // see bqa_fast_d3D4.cu
  int v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// Per-lane constants of the copy pipeline (all copy offsets are immediates).  A group is 4 consecutive nodes starting
// at n0 = min(4 g, B - 4): the last group of a class whose size is not a multiple of 4 overlaps its predecessor (those
// nodes are computed and stored twice with identical results), so nothing is predicated.
//   index register: lanes 0..11 hold in_pos[j][n0 + s], lanes 12..23 hold lmbd_pos[j][n0 + s]  (index j * 4 + s)
struct Pipe {
  unsigned sT, sC, sL;             // shared-memory destinations (stage 0) of this lane's first T chunk / canonicalizer piece / lambdas
  const unsigned char* gT;         // a.T + 16 lane
  const unsigned char* gC;         // a.canon + this lane's (row, half) offset inside a canonicalizer
  const int32_t* idx_ptr;          // this lane's row of in_pos / lmbd_pos (+ node slot); nullptr for lanes >= 24
};

__device__ __forceinline__ int load_idx(const Pipe& q, int n0) { return q.idx_ptr ? ldg_ordered(q.idx_ptr + n0) : 0; }

__device__ __forceinline__ void issue_group(const Args& a, const Pipe& q, unsigned stage_off, int n0, int lane, int idx_reg) {
  const unsigned char* src = q.gT + (size_t)(unsigned)n0 * 1024;
  const unsigned dT = q.sT + stage_off, dC = q.sC + stage_off;
#pragma unroll
  for (int i = 0; i < 8; ++i) cp_async16_s(dT + i * 4 * kSlice, src + i * 512);
#pragma unroll
  for (int i = 0; i < 6; ++i) {                  // 192 pieces of 16 bytes: block (leg, node) = i * 2 + lane / 16
    const unsigned slot = (unsigned)__shfl_sync(0xffffffffu, idx_reg, i * 2 + (lane >> 4));
    cp_async16_s(dC + i * 2 * kCBlk, q.gC + (size_t)slot * 512);
  }
  const unsigned lp = (unsigned)__shfl_sync(0xffffffffu, idx_reg, 12 + (lane < 12 ? lane : 0));
  if (lane < 12) cp_async16_s(q.sL + stage_off, reinterpret_cast<const unsigned char*>(a.lmbds) + (size_t)lp * 32);
}

// complex multiply-add with a prepared pair operand: acc += s * w, (wp, iw) = (w, i w) as pairs, s = (sx, sy) scalars
__device__ __forceinline__ p2 cfma_pair(float sx, float sy, p2 wp, p2 iw, p2 acc) {
  return x2::fma2s(sy, iw, x2::fma2s(sx, wp, acc));
}

// Requires 4 <= B < 2^29.
__global__ void __launch_bounds__(kThreads, 1) k_apply_d3D4(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = lane >> 3, t = lane & 7, p = t >> 2, la = t & 3;
  unsigned char* wbase = smem + wib * kWarpBytes;
  unsigned char* Wm = wbase + 2 * kStage + s * (3 * 2 * 128);       // [leg][p][a][c] of this lane's node
  const int B = (int)a.B;
  const int groups = (B + 3) >> 2, tail0 = B - 4;
  const int nwarps = (int)gridDim.x * kWarps;
  int g = wib * (int)gridDim.x + (int)blockIdx.x;                   // warp-major (see bqa_fast_d3D4.cu)
  const float cxt = cosf(a.xtime), sxt = sinf(a.xtime);
  const p2 rot = x2::pk(-1.f, 1.f);                                 // (y, x) * rot = i (x + i y)

  Pipe q;
  q.sT = smem_u32(wbase) + (lane >> 3) * kSlice + (lane & 7) * 16;
  q.sC = smem_u32(wbase) + kTBytes + (lane >> 4) * kCBlk + ((lane >> 1) & 7) * 32 + (lane & 1) * 16;
  q.sL = smem_u32(wbase) + kTBytes + kCBytes + lane * 16;
  q.gT = reinterpret_cast<const unsigned char*>(a.T) + lane * 16;
  q.gC = reinterpret_cast<const unsigned char*>(a.canon) + ((lane >> 1) & 7) * 64 + (lane & 1) * 16;
  {
    const int l = lane < 12 ? lane : lane - 12;
    q.idx_ptr = lane < 24 ? (lane < 12 ? a.in_pos : a.lmbd_pos) + (size_t)(l >> 2) * B + (l & 3) : nullptr;
  }

  int idx_cur = 0, idx_nxt = 0;
  if (g < groups) {
    const int n0 = min(g * 4, tail0);
    idx_cur = load_idx(q, n0);
    issue_group(a, q, 0u, n0, lane, idx_cur);
    cp_async_commit();
    if (g + nwarps < groups) idx_nxt = load_idx(q, min((g + nwarps) * 4, tail0));
  }
  int cur = 0;
#pragma unroll 1
  for (; g < groups; g += nwarps, cur ^= 1) {
    unsigned char* st = wbase + cur * kStage;
    const int n0 = min(g * 4, tail0);
    int idx_nn = 0;
    if (g + nwarps < groups) {
      issue_group(a, q, cur ? 0u : (unsigned)kStage, min((g + nwarps) * 4, tail0), lane, idx_nxt);
      if (g + 2 * nwarps < groups) idx_nn = load_idx(q, min((g + 2 * nwarps) * 4, tail0));
    }
    cp_async_commit();
    // one angle per lane: lane t of a node evaluates leg t % 4 < 3 (theta = J zt) or, t % 4 = 3, the node's field angle
    // (phi = h zt); the ZZ factors and cos / sin(phi) reach the other lanes through shuffles -- one sincos per lane and
    // group instead of five
    const int node = n0 + s, sel = t & 3;
    const float ang = __ldg((sel < 3 ? a.edge_ampls + (size_t)sel * B : a.node_ampls) + node) * a.ztime;
    const float cang = cosf(ang), sang = sinf(ang);
    cx<float> zf0, zf1;
    zz_factors_cs<float>(cang, sang, zf0, zf1);
    cp_async_wait<1>();
    __syncwarp();

    // ---- W_j^(p)[a][c]: lane t builds row a = t % 4, columns 2 (t / 4) and 2 (t / 4) + 1 of both p, for the 3 legs
    {
      const int wa = t & 3, cp2 = t >> 2;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const unsigned char* blk = st + kTBytes + (j * 4 + s) * kCBlk;
        const float4 hi = *reinterpret_cast<const float4*>(blk + wa * 32 + cp2 * 16);
        const float4 lo = *reinterpret_cast<const float4*>(blk + (4 + wa) * 32 + cp2 * 16);
        const float2 lm = *reinterpret_cast<const float2*>(st + kTBytes + kCBytes + (j * 4 + s) * 16 + cp2 * 8);
        const float2 g0 = make_float2(__shfl_sync(0xffffffffu, zf0.re, j, 8), __shfl_sync(0xffffffffu, zf0.im, j, 8));
        const float2 g1 = make_float2(__shfl_sync(0xffffffffu, zf1.re, j, 8), __shfl_sync(0xffffffffu, zf1.im, j, 8));
        const float s0 = sqrtf(lm.x), s1 = sqrtf(lm.y);
        const float2 h0 = cmul(g0, make_float2(hi.x, hi.y)), h1 = cmul(g0, make_float2(hi.z, hi.w));
        const float2 l0 = cmul(g1, make_float2(lo.x, lo.y)), l1 = cmul(g1, make_float2(lo.z, lo.w));
        *reinterpret_cast<float4*>(Wm + (j * 2 + 0) * 128 + wa * 32 + cp2 * 16) =
            make_float4(s0 * (h0.x + l0.x), s0 * (h0.y + l0.y), s1 * (h1.x + l1.x), s1 * (h1.y + l1.y));
        *reinterpret_cast<float4*>(Wm + (j * 2 + 1) * 128 + wa * 32 + cp2 * 16) =
            make_float4(s0 * (h0.x - l0.x), s0 * (h0.y - l0.y), s1 * (h1.x - l1.x), s1 * (h1.y - l1.y));
      }
    }
    __syncwarp();

    // packed arithmetic (FFMA2, see bqa_fast_common.cuh): a complex number is one (re, im) register pair
    unsigned char* Ts = st + (s * 8 + p * 4) * kSlice;
    p2 X[16];
    {
      p2 tt[16], w[16], Y[16];
      lds_tile(tt, Ts + la * kSlice);
      lds_tile(w, Wm + (2 * 2 + p) * 128);                  // W_2^(p)[c][c']
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          CAcc acc;
          cmac<true>(acc, tt[b * 4], w[c2]);
#pragma unroll
          for (int c = 1; c < 4; ++c) cmac<false>(acc, tt[b * 4 + c], w[c * 4 + c2]);
          Y[b * 4 + c2] = cfinish(acc);
        }
      lds_tile(w, Wm + (1 * 2 + p) * 128);                  // W_1^(p)[b][b']
#pragma unroll
      for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          CAcc acc;
          cmac<true>(acc, Y[c2], w[b2]);
#pragma unroll
          for (int b = 1; b < 4; ++b) cmac<false>(acc, Y[b * 4 + c2], w[b * 4 + b2]);
          X[b2 * 4 + c2] = cfinish(acc);
        }
    }
    __syncwarp();                                           // every lane has read its T slice
    sts_tile(Ts + la * kSlice, X);                          // exchange X over leg 0
    __syncwarp();
    p2 R[16];
#pragma unroll
    for (int a0 = 0; a0 < 4; ++a0) {
      p2 xs[16];
      lds_tile(xs, Ts + a0 * kSlice);
      const p2 wp = *reinterpret_cast<const p2*>(Wm + (0 * 2 + p) * 128 + a0 * 32 + la * 8);          // W_0^(p)[a0][c0 = la]
      const p2 iw = x2::mul2(x2::swap(wp), rot);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 xf = x2::unpk(xs[i]);
        R[i] = a0 == 0 ? x2::fma2s(xf.y, iw, x2::mul2s(xf.x, wp)) : cfma_pair(xf.x, xf.y, wp, iw, R[i]);
      }
    }
    __syncwarp();
    sts_tile(Ts + la * kSlice, R);                          // T'[p][c0 = la] for the partner of the Rx mixing
    __syncwarp();
    {
      p2 other[16];
      lds_tile(other, st + (s * 8 + (1 - p) * 4 + la) * kSlice);
      // out_p = cos(xt) z_p T'_p - i sin(xt) z_{1-p} T'_{1-p},  z_0 = e^{-i phi}, z_1 = e^{+i phi}
      const float cph = __shfl_sync(0xffffffffu, cang, 3, 8), sph = __shfl_sync(0xffffffffu, sang, 3, 8);
      const float2 zm = make_float2(cph, p ? sph : -sph), zo = make_float2(cph, p ? -sph : sph);
      const p2 ca = x2::pk(cxt * zm.x, cxt * zm.y), ica = x2::mul2(x2::swap(ca), rot);
      const p2 cb = x2::pk(sxt * zo.y, -sxt * zo.x), icb = x2::mul2(x2::swap(cb), rot);   // -i sin(xt) z_other
      p2 n2p = x2::pk(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 rf = x2::unpk(R[i]), of = x2::unpk(other[i]);
        p2 o = x2::fma2s(rf.y, ica, x2::mul2s(rf.x, ca));
        o = cfma_pair(of.x, of.y, cb, icb, o);
        R[i] = o;
        n2p = x2::fma2(o, o, n2p);
      }
      float n2 = x2::hsum(n2p);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      const float inv = 1.f / sqrtf(n2);
#pragma unroll
      for (int i = 0; i < 16; ++i) R[i] = x2::mul2s(inv, R[i]);
    }
    __syncwarp();                                           // partners have read T'
    sts_tile(Ts + la * kSlice, R);
    __syncwarp();
    // coalesced store of the 4 KB block
    unsigned char* Og = reinterpret_cast<unsigned char*>(a.Tout) + (size_t)(unsigned)n0 * 1024 + lane * 16;
    const unsigned char* Os = st + (lane >> 3) * kSlice + (lane & 7) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(Og + i * 512) = *reinterpret_cast<const float4*>(Os + i * 4 * kSlice);
    idx_cur = idx_nxt;
    idx_nxt = idx_nn;
    __syncwarp();
  }
}

}  // namespace fast_apply

bool fast_apply_available(int prec, int degree, int D, int Dn, long long B) {
  return prec == 0 && degree == 3 && D == 4 && Dn == 4 && (B == 0 || (B >= 4 && B < (1LL << 29)));
}

int launch_fast_apply_d3D4(long long B, const void* T_in, void* T_out, const void* canon, const void* lmbds,
                           const int32_t* in_pos, const int32_t* lmbd_pos, const void* node_ampls,
                           const void* edge_ampls, double ztime, double xtime, cudaStream_t st) {
  using namespace fast_apply;
  if (B == 0) return 0;
  static bool configured[64] = {};           // the attribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (cur_dev < 0 || cur_dev >= 64 || !configured[cur_dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_apply_d3D4, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(k_apply_d3D4): %s", cudaGetErrorString(e));
    if (cur_dev >= 0 && cur_dev < 64) configured[cur_dev] = true;
  }
  Args a{};
  a.B = B; a.T = (const float2*)T_in; a.Tout = (float2*)T_out; a.canon = (const float2*)canon; a.lmbds = (const float*)lmbds;
  a.in_pos = in_pos; a.lmbd_pos = lmbd_pos; a.node_ampls = (const float*)node_ampls; a.edge_ampls = (const float*)edge_ampls;
  a.ztime = (float)ztime; a.xtime = (float)xtime;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long groups = (B + 3) / 4;
  long long grid = (groups + kWarps - 1) / kWarps;
  if (grid > sms) grid = sms;
  k_apply_d3D4<<<(int)grid, kThreads, kSmem, st>>>(a);
  return after_launch("apply_update(d3D4)");
}

}  // namespace bqa
