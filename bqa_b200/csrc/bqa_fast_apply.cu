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

constexpr int kWarps = 8;
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

// lanes 0..11 hold in_pos[j][node0 + s], lanes 12..23 hold lmbd_pos[j][node0 + s]  (index j * 4 + s)
__device__ __forceinline__ int load_idx(const Args& a, long long node0, int lane) {
  int v = 0;
  if (lane < 24) {
    const int l = lane < 12 ? lane : lane - 12;
    long long node = node0 + (l & 3);
    node = node > a.B - 1 ? a.B - 1 : node;
    const int32_t* src = lane < 12 ? a.in_pos : a.lmbd_pos;
    v = __ldg(src + (size_t)(l >> 2) * a.B + node);
  }
  return v;
}

__device__ __forceinline__ void issue_group(const Args& a, unsigned char* stage, long long node0, int lane, int idx_reg) {
  const long long last = a.B - 1;
  const unsigned char* Tg = reinterpret_cast<const unsigned char*>(a.T);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + lane;
    long long node = node0 + (c >> 6);
    node = node > last ? last : node;
    cp_async16(stage + (c >> 3) * kSlice + (c & 7) * 16, Tg + node * 1024 + (c & 63) * 16);
  }
  const unsigned char* Cg = reinterpret_cast<const unsigned char*>(a.canon);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 32 + lane;                 // 192 pieces of 16 bytes: block (leg, node) = c / 16, row = (c % 16) / 2
    const int blk = c >> 4, row = (c >> 1) & 7, half = c & 1;
    const int slot = __shfl_sync(0xffffffffu, idx_reg, blk);
    cp_async16(stage + kTBytes + blk * kCBlk + row * 32 + half * 16, Cg + (size_t)slot * 512 + row * 64 + half * 16);
  }
  const int lp = __shfl_sync(0xffffffffu, idx_reg, 12 + (lane < 12 ? lane : 0));
  if (lane < 12)
    cp_async16(stage + kTBytes + kCBytes + lane * 16, reinterpret_cast<const unsigned char*>(a.lmbds) + (size_t)lp * 32);
}

__global__ void __launch_bounds__(kThreads, 1) k_apply_d3D4(Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = lane >> 3, t = lane & 7, p = t >> 2, la = t & 3;
  unsigned char* wbase = smem + wib * kWarpBytes;
  unsigned char* Wm = wbase + 2 * kStage + s * (3 * 2 * 128);       // [leg][p][a][c] of this lane's node
  const long long groups = (a.B + 3) >> 2;
  const long long nwarps = (long long)gridDim.x * kWarps;
  long long g = (long long)blockIdx.x * kWarps + wib;
  const float cxt = cosf(a.xtime), sxt = sinf(a.xtime);

  int idx_cur = 0, idx_nxt = 0;
  if (g < groups) {
    idx_cur = load_idx(a, g * 4, lane);
    issue_group(a, wbase, g * 4, lane, idx_cur);
    cp_async_commit();
    if (g + nwarps < groups) idx_nxt = load_idx(a, (g + nwarps) * 4, lane);
  }
  int cur = 0;
#pragma unroll 1
  for (; g < groups; g += nwarps, cur ^= 1) {
    unsigned char* st = wbase + cur * kStage;
    int idx_nn = 0;
    if (g + nwarps < groups) {
      issue_group(a, wbase + (cur ^ 1) * kStage, (g + nwarps) * 4, lane, idx_nxt);
      if (g + 2 * nwarps < groups) idx_nn = load_idx(a, (g + 2 * nwarps) * 4, lane);
    }
    cp_async_commit();
    long long node = g * 4 + s;
    const bool live = node < a.B;
    node = live ? node : a.B - 1;
    float th[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) th[j] = __ldg(a.edge_ampls + (size_t)j * a.B + node) * a.ztime;
    const float phi = __ldg(a.node_ampls + node) * a.ztime;
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
        cx<float> f0, f1;
        zz_factors<float>(th[j], f0, f1);
        const float2 g0 = make_float2(f0.re, f0.im), g1 = make_float2(f1.re, f1.im);
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

    unsigned char* Ts = st + (s * 8 + p * 4) * kSlice;
    float2 X[16];
    {
      float2 tt[16], w[16], Y[16];
      lds_tile(tt, Ts + la * kSlice);
      lds_tile(w, Wm + (2 * 2 + p) * 128);                  // W_2^(p)[c][c']
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) fma_c(acc, tt[b * 4 + c], w[c * 4 + c2]);
          Y[b * 4 + c2] = acc;
        }
      lds_tile(w, Wm + (1 * 2 + p) * 128);                  // W_1^(p)[b][b']
#pragma unroll
      for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int b = 0; b < 4; ++b) fma_c(acc, Y[b * 4 + c2], w[b * 4 + b2]);
          X[b2 * 4 + c2] = acc;
        }
    }
    __syncwarp();                                           // every lane has read its T slice
    sts_tile(Ts + la * kSlice, X);                          // exchange X over leg 0
    __syncwarp();
    float2 R[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) R[i] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int a0 = 0; a0 < 4; ++a0) {
      float2 xs[16];
      lds_tile(xs, Ts + a0 * kSlice);
      const float2 w = *reinterpret_cast<const float2*>(Wm + (0 * 2 + p) * 128 + a0 * 32 + la * 8);    // W_0^(p)[a0][c0 = la]
#pragma unroll
      for (int i = 0; i < 16; ++i) fma_c(R[i], w, xs[i]);
    }
    __syncwarp();
    sts_tile(Ts + la * kSlice, R);                          // T'[p][c0 = la] for the partner of the Rx mixing
    __syncwarp();
    {
      float2 other[16];
      lds_tile(other, st + (s * 8 + (1 - p) * 4 + la) * kSlice);
      // out_p = cos(xt) z_p T'_p - i sin(xt) z_{1-p} T'_{1-p},  z_0 = e^{-i phi}, z_1 = e^{+i phi}
      const float cph = cosf(phi), sph = sinf(phi);
      const float2 zm = make_float2(cph, p ? sph : -sph), zo = make_float2(cph, p ? -sph : sph);
      const float2 ca = make_float2(cxt * zm.x, cxt * zm.y);
      const float2 cb = make_float2(sxt * zo.y, -sxt * zo.x);                  // -i sin(xt) z_other
      float n2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 o = cmul(ca, R[i]);
        fma_c(o, cb, other[i]);
        R[i] = o;
        n2 = fmaf(o.x, o.x, n2); n2 = fmaf(o.y, o.y, n2);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      const float inv = 1.f / sqrtf(n2);
#pragma unroll
      for (int i = 0; i < 16; ++i) { R[i].x *= inv; R[i].y *= inv; }
    }
    __syncwarp();                                           // partners have read T'
    sts_tile(Ts + la * kSlice, R);
    __syncwarp();
    // coalesced store of the 4 KB block (ragged last group: only the live nodes)
    unsigned char* Og = reinterpret_cast<unsigned char*>(a.Tout) + (size_t)g * 4096;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 32 + lane;
      if (g * 4 + (c >> 6) < a.B)
        *reinterpret_cast<float4*>(Og + c * 16) = *reinterpret_cast<const float4*>(st + (c >> 3) * kSlice + (c & 7) * 16);
    }
    idx_cur = idx_nxt;
    idx_nxt = idx_nn;
    __syncwarp();
  }
}

}  // namespace fast_apply

bool fast_apply_available(int prec, int degree, int D, int Dn) { return prec == 0 && degree == 3 && D == 4 && Dn == 4; }

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
