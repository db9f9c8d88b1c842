// bqa_fast_d3D4.cu -- specialised sm_100a kernels for the headline shape: degree-3 nodes, bond dimension 4,
// complex64 (random 3-regular QUBO at max_bond_dim 4: BASELINE.json configs 4 and 5).
//
//   k_msgs_d3D4<false, MULTI> : one BP sweep     (Tensor.pass_msgs, src/bqa/backends.py:381-408, + get_dist :492-495
//                                                 + damping :539-540 + the gathers/scatters of state.py:109-112)
//   k_msgs_d3D4<true, true>   : ZZ-extended messages (_get_extended_msgs, state.py:127-139; backends.py:519-526)
//   k_bp_run_d3D4<MULTI>      : the whole BP run (_run_bp, state.py:97-124) in one cooperative launch
//   (MULTI: boundary messages are also stored into the peers' halo slots over NVLink)
//
// Work decomposition.  Eight lanes own a node: lane (p, a) holds the slice T[p, a, :, :] (16 complex) in
// registers, a warp works on 4 consecutive nodes, a CTA of 8 warps is persistent (one per SM) and every warp
// runs its own two-stage cp.async pipeline: the contiguous 4 KB of node tensors and the 24 gathered 128-byte
// messages (3 incoming + 3 previous outgoing per node) of the NEXT group stream into shared memory while the
// current group is contracted out of registers.
//
// Arithmetic.  With Hermitian messages (BP messages are Hermitian positive semi-definite by construction)
//   U_j = T x_j m_j                                   (3 mode products instead of the reference's 6)
//   out_0[x,y] = sum conj(U_1[x,b,c]) U_2[y,b,c],  out_1[x,y] = sum conj(U_2[a,x,c]) U_0[a,y,c],
//   out_2[x,y] = sum conj(U_1[a,b,x]) U_0[a,b,y]
// is the same sum as conj(T) . prod_{j != k} m_j . T.  Legs 1 and 2 are thread-local; leg 0 is spread over the 4 `a`
// lanes, so U_0 and out_0 read the other lanes' slices from shared memory (padded 144-byte slices: conflict-free).
// Only the Hermitian half of every output is computed (real diagonals, upper triangle; out_0: entries (a, a),
// (a + 1, a), (a + 2, a) per lane), the diagonal of a message is real (half a multiply-add), arithmetic is packed
// (FFMA2: a complex number is one register pair): ~600 packed multiply-adds per lane and 4-node group.  The per-lane
// partial sums of out_1 / out_2 are reduce-scattered over the node's 8 lanes with xor shuffles into the 16-byte piece
// of the message each lane stores; the traces are all-reduced early from the partial diagonals.
#include <cuda_runtime.h>

#include <cstdlib>

#include "bqa_core.cuh"
#include "bqa_fast_common.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace fast {

// 8 warps per CTA = 2 per SM sub-partition: the kernel needs ~230 registers to keep three 16-element tiles, the staged
// tile and the copy pipeline's addresses live without spills; 3 warps per sub-partition (12 per CTA) are capped at 168
// registers by the 16 K registers of a sub-partition and measured 9 % slower (62.6 vs 57.4 us per sweep).
#ifndef BQA_BP_WARPS
#define BQA_BP_WARPS 8
#endif
constexpr int kWarps = BQA_BP_WARPS;
constexpr int kThreads = kWarps * 32;
constexpr int kSlice = 144;                       // 128-byte slice + 16 bytes of padding (bank spreading)
constexpr int kTBytes = 32 * kSlice;              // 4 nodes x 8 slices
#ifndef BQA_BP_MSG_PITCH
#define BQA_BP_MSG_PITCH 144
#endif
constexpr int kMsg = BQA_BP_MSG_PITCH;            // pitch of a 128-byte message: 144 spreads the 4 nodes' tiles over the banks
constexpr int kMBytes = 12 * kMsg;                // 3 messages x 4 nodes
constexpr int kRedRow = 80;                       // packed Hermitian partial: 4 diagonal + 6 upper entries (complex)
constexpr int kRedSlot = 8 * kRedRow + 64;        // 8 lanes of a node (+ skew between the two nodes of a half warp)
constexpr int kOut0 = 4 * kRedSlot;               // scratch offset of the out_0 exchange (4 rows of 288 bytes)
constexpr int kOut0Row = 288;
constexpr int kStage = kTBytes + 2 * kMBytes;     // T | incoming messages | previous outgoing messages
constexpr int kScratch = kOut0 + 4 * kOut0Row;     // packed partials of out_1 / out_2 + the out_0 exchange
constexpr int kWarpBytes = 2 * kStage + kScratch; // two stages + reduction scratch
constexpr int kSmem = kWarps * kWarpBytes;
static_assert(kSmem <= 227 * 1024, "BQA_BP_WARPS x (two stages + scratch) exceeds the 227 KB of shared memory a CTA can own");

struct Args {
  long long B;
  const float2* T;
  const float2* msgs_cur;
  float2* msgs_out;
  const int32_t* in_pos;
  const int32_t* out_pos;
  const float* edge_ampls;
  float ztime, damping, bp_eps;
  int write_undamped, it;
  float* resid;
  int32_t* status;
  // multi-GPU over peer memory (see NodeArgs in bqa_generic.cuh)
  const int32_t* remote_pos;
  unsigned char* peers[BQA_MAX_PEERS];
  // extended messages launched BEHIND a single-launch BP run whose outcome the host has not read yet: the kernel picks
  // the buffer the run left the messages in from the run's status words (the host's rule, engine._finish_bp)
  const int32_t* sel_status;       // null: msgs_cur is used
  const float2* sel_msgs[3];
  int sel_parity, sel_nbuf, sel_max_iters;
};

__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// sum over the 8 lanes of a node (aligned groups of 8 lanes)
__device__ __forceinline__ float allreduce8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float abs2_rn(float x, float y) { return fmaf(x, x, __fmul_rn(y, y)); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_s(unsigned dst, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem) : "memory");
}

// Per-lane constants of the copy pipeline.  A group is 4 consecutive nodes starting at n0 = min(4 g, B - 4): the last
// group of a class whose size is not a multiple of 4 overlaps its predecessor (those nodes are computed twice with
// identical results), so no copy and no store is ever predicated.
//   index register: lane 8 s + i holds in_pos[i][n0 + s] (i = 0..2) and out_pos[i - 3][n0 + s] (i = 3..5)
struct Pipe {
  unsigned sT, sM;                 // shared-memory destinations of this lane's first T chunk / message chunk (stage 0)
  const unsigned char* gT;         // a.T + 16 lane
  const unsigned char* gM;         // a.msgs_cur + 16 (lane % 8)
  const int32_t* idx_ptr;          // this lane's row of in_pos / out_pos (+ node slot); nullptr for lanes t >= 6
  const int32_t* rp_ptr;           // this lane's row of remote_pos; nullptr on one GPU and for lanes t >= 3
};

// (volatile: the loads stay where they are written, behind the copies that consume the previous index register --
// hoisted above them they share a scoreboard with the older load and stall its consumers for a full memory latency)
__device__ __forceinline__ int ldg_ordered(const int32_t* ptr) {
  int v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ int load_idx(const Pipe& q, int n0) { return q.idx_ptr ? ldg_ordered(q.idx_ptr + n0) : 0; }
__device__ __forceinline__ int load_rpos(const Pipe& q, int n0) { return q.rp_ptr ? ldg_ordered(q.rp_ptr + n0) : -1; }

// issue the copies of one 4-node group into a stage: T (coalesced 4 KB), 12 incoming and 12 previous outgoing
// messages (each by the 8 lanes of a quarter warp: one full 128-byte line).  All offsets are immediates.
template <bool EXT>
__device__ __forceinline__ void issue_group(const Pipe& q, unsigned stage_off, int n0, int idx_reg) {
  const unsigned char* src = q.gT + (size_t)(unsigned)n0 * 1024;
  const unsigned dT = q.sT + stage_off, dM = q.sM + stage_off;
#pragma unroll
  for (int i = 0; i < 8; ++i) cp_async16_s(dT + i * 4 * kSlice, src + i * 512);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const unsigned pin = (unsigned)__shfl_sync(0xffffffffu, idx_reg, j, 8);
    cp_async16_s(dM + j * 4 * kMsg, q.gM + (size_t)pin * 128);
    if (!EXT) {
      const unsigned pout = (unsigned)__shfl_sync(0xffffffffu, idx_reg, 3 + j, 8);
      cp_async16_s(dM + kMBytes + j * 4 * kMsg, q.gM + (size_t)pout * 128);
    }
  }
}

// One sweep (or one extended-message pass) over the degree class by this CTA's warps.  Requires 4 <= B < 2^29.
// (the per-sweep fields are parameters and the peer table is an accessor so that the single-launch kernel can keep its
// arguments in constant memory: a modified local copy of `Args` would live on the stack because of the dynamic peer index)
struct PeersOfArgs {
  const Args& a;
  __device__ __forceinline__ unsigned char* operator()(int q) const { return a.peers[q]; }
};
struct NoMid {
  __device__ __forceinline__ void operator()() const {}
};
// mid() runs once per warp, after the warp's last group below g_mid (the multi-GPU run: "my boundary groups are stored")
// while the copies of its next group are already in flight
template <bool EXT, bool MULTI, class Peers, class Mid = NoMid>
__device__ __forceinline__ void sweep(const Args& a, const float2* msgs_cur, float2* msgs_out, int it, int write_undamped,
                                      const Peers peers, unsigned char* smem, int g_lo = 0, int g_hi = 0x7fffffff,
                                      bool rev = false, int g_mid = 0, const Mid mid = Mid{}) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int s = lane >> 3, t = lane & 7, p = t >> 2, la = t & 3;     // node slot, lane in node, physical, leg-0 index
  unsigned char* wbase = smem + wib * kWarpBytes;
  unsigned char* scratch = wbase + 2 * kStage;
  const int B = (int)a.B;
  // groups [g_lo, g_hi) of the class (default: all of them; the multi-GPU run sweeps the boundary groups first)
  const int groups = min((B + 3) >> 2, g_hi), tail0 = B - 4;
  const int nwarps = (int)gridDim.x * kWarps;
  // warp-major numbering: the groups of the last, partial round land on one warp each of as many CTAs as possible
  // (a lone warp runs much faster than eight sharing the SM) instead of filling all warps of a few CTAs
  int g = wib * (int)gridDim.x + (int)blockIdx.x;
  if (g < g_lo) g += (g_lo - g + nwarps - 1) / nwarps * nwarps;
  // rev: the groups are visited from the last to the first.  The single-GPU run alternates the direction from sweep to
  // sweep, so a sweep starts on the part of T and of the messages that the previous one left in L2 (the working set of
  // the 100k instance, 180 MB, cycles through the 126 MB L2 otherwise); every sweep reads one buffer and writes the
  // other, so the order does not enter the results
  const int gflip = rev ? groups - 1 : 0, gdir = rev ? -4 : 4;
  auto first_node = [&](int gg) { return min(gflip * 4 + gdir * gg, tail0); };
  // boundary messages are also stored into the peers' halo slots.  Compile-time in the BP kernels; the extended-message
  // kernel has ONE instantiation that tests the pointer (two instantiations contracted its epilogue arithmetic into FMAs
  // differently, and single- and multi-GPU runs must stay bit-identical)
  const bool multi = EXT ? a.remote_pos != nullptr : MULTI;
  float mnum = 0.f, mden = 0.f;

  Pipe q;
  q.sT = smem_u32(wbase) + (lane >> 3) * kSlice + (lane & 7) * 16;
  q.sM = smem_u32(wbase) + kTBytes + s * kMsg + t * 16;
  q.gT = reinterpret_cast<const unsigned char*>(a.T) + lane * 16;
  q.gM = reinterpret_cast<const unsigned char*>(msgs_cur) + t * 16;
  q.idx_ptr = t < 6 ? (t < 3 ? a.in_pos + (size_t)t * B : a.out_pos + (size_t)(t - 3) * B) + s : nullptr;
  q.rp_ptr = (multi && t < 3) ? a.remote_pos + (size_t)t * B + s : nullptr;

  // packed Hermitian partials: entry (x, x) at complex index x, entry (x < y) at 4 + pair(x, y); this lane finally
  // owns (x, y0) and (x, y0 + 1), x = t / 2, y0 = 2 (t % 2); a lower-triangle entry is the conjugate of its mirror
  int offA = 0, offB = 0;
  float sgnA = 1.f, sgnB = 1.f;
  if (EXT) {
    const int x = t >> 1, y0 = (t & 1) * 2;
    auto pack = [](int i, int j) { return i == j ? i : 4 + (i == 0 ? j - 1 : (i == 1 ? j + 1 : 5)); };
    const int ya = y0, yb = y0 + 1;
    offA = 8 * (x <= ya ? pack(x, ya) : pack(ya, x));
    offB = 8 * (x <= yb ? pack(x, yb) : pack(yb, x));
    sgnA = x > ya ? -1.f : 1.f;
    sgnB = x > yb ? -1.f : 1.f;
  }
  // BP path: the packed partials of out_1 / out_2 are reduce-scattered over the 8 lanes of a node with xor shuffles
  // (16 reals -> 2 per lane), then every lane fetches the one pair it still misses from a partner lane.  Lane t owns
  // pair t of  U01 | U02 | (D0, D1) | U13 | U12 | U23 | U03 | (D2, D3)  and needs elements (x, y0), (x, y0 + 1).
  // (tables packed into immediates: partner 4 bits, flags 1 bit, signs 2 bits per lane -- no local-memory arrays)
  const int partner = (lane & ~7) | ((0x53714062u >> (4 * t)) & 7);   // lane whose pair completes this lane's row piece
  const bool a_got = (0xBDu >> t) & 1, b_got = (0x42u >> t) & 1;      // element a = (x, y0), b = (x, y0 + 1): own or fetched pair
  const bool b_usey = (0x84u >> t) & 1;                               // diagonal entry held in the .y slot of (D, D) pairs
  const int a_c = (0xA264u >> (2 * t)) & 3, b_c = (0x2645u >> (2 * t)) & 3;
  const float a_im = a_c == 0 ? 0.f : (a_c == 1 ? 1.f : -1.f);        // sign of the imaginary part (0 on the diagonal)
  const float b_im = b_c == 0 ? 0.f : (b_c == 1 ? 1.f : -1.f);
  const p2 rot = x2::pk(-1.f, 1.f);                                   // (y, x) * rot = (-y, x) = i (x + i y)
  int idx_cur = 0, idx_nxt = 0, rp_cur = -1, rp_nxt = -1;
  if (g < groups) {
    const int n0 = first_node(g);
    idx_cur = load_idx(q, n0);
    rp_cur = load_rpos(q, n0);
    issue_group<EXT>(q, 0u, n0, idx_cur);
    cp_async_commit();
    if (g + nwarps < groups) {
      const int n1 = first_node(g + nwarps);
      idx_nxt = load_idx(q, n1);
      rp_nxt = load_rpos(q, n1);
    }
  }
  int cur = 0;
  if (g >= g_mid) mid();
#pragma unroll 1
  for (; g < groups; g += nwarps, cur ^= 1) {
    unsigned char* st = wbase + cur * kStage;
    const int n0 = first_node(g);
    int idx_nn = 0, rp_nn = -1;
    if (g + nwarps < groups) {
      issue_group<EXT>(q, cur ? 0u : (unsigned)kStage, first_node(g + nwarps), idx_nxt);
      if (g + 2 * nwarps < groups) {
        const int n2 = first_node(g + 2 * nwarps);
        idx_nn = load_idx(q, n2);
        rp_nn = load_rpos(q, n2);
      }
    }
    cp_async_commit();                                      // always commit (possibly empty): one loop body, one wait
    cp_async_wait<1>();
    __syncwarp();

    // ZZ half-gate factors (extended messages): lane t of a node evaluates leg min(t % 4, 2) -- one sincos per lane and
    // group; the epilogue fetches leg k's factors from lane k of the node
    cx<float> zf0 = mk<float>(0.f, 0.f), zf1 = zf0;
    if (EXT) {
      const int sel = (t & 3) < 2 ? (t & 3) : 2;
      zz_factors<float>(__ldg(a.edge_ampls + (size_t)sel * B + n0 + s) * a.ztime, zf0, zf1);
    }
    const unsigned char* Ts = st + (s * 8 + p * 4) * kSlice;          // the four a-slices of (node, p)
    const unsigned char* Min = st + kTBytes;
    const unsigned char* m0row = Min + (0 * 4 + s) * kMsg + la * 32;  // row `la` of m0
    p2 U0[16], U1[16], U2[16];
    {
      p2 tt[16], m[16];
      lds_tile(tt, Ts + la * kSlice);
      // U2[b][c'] = sum_c m2[c'][c] T[b][c]   (the diagonal of a Hermitian message is real: half a multiply-add there)
      lds_tile(m, Min + (2 * 4 + s) * kMsg);
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          CAcc acc;
          if (c2 == 0) cmac_real<true>(acc, m[0], tt[b * 4]); else cmac<true>(acc, m[c2 * 4], tt[b * 4]);
#pragma unroll
          for (int c = 1; c < 4; ++c) {
            if (c == c2) cmac_real<false>(acc, m[c2 * 4 + c], tt[b * 4 + c]);
            else if (c2 == 0 && c == 1) cmac_bfirst<true>(acc, m[c2 * 4 + c], tt[b * 4 + c]);
            else cmac<false>(acc, m[c2 * 4 + c], tt[b * 4 + c]);
          }
          U2[b * 4 + c2] = cfinish(acc);
        }
      // U1[b'][c] = sum_b m1[b'][b] T[b][c]
      lds_tile(m, Min + (1 * 4 + s) * kMsg);
#pragma unroll
      for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          CAcc acc;
          if (b2 == 0) cmac_real<true>(acc, m[0], tt[c]); else cmac<true>(acc, m[b2 * 4], tt[c]);
#pragma unroll
          for (int b = 1; b < 4; ++b) {
            if (b == b2) cmac_real<false>(acc, m[b2 * 4 + b], tt[b * 4 + c]);
            else if (b2 == 0 && b == 1) cmac_bfirst<true>(acc, m[b2 * 4 + b], tt[b * 4 + c]);
            else cmac<false>(acc, m[b2 * 4 + b], tt[b * 4 + c]);
          }
          U1[b2 * 4 + c] = cfinish(acc);
        }
      // U0[a][b][c] = sum_a' m0[a][a'] T[a'][b][c], row a = `la` of m0.  The matrix element is the prepared pair operand
      // (m, i m), the tensor elements enter as broadcast scalars.  Term a' = la: this lane's own slice, still in
      // registers, times the real diagonal element
      {
        const float mr = *reinterpret_cast<const float*>(m0row + la * 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) U0[i] = x2::mul2s(mr, tt[i]);
      }
    }
    // the three other slices from shared memory: the 8 lanes of a node read 8 distinct slices (conflict-free)
    // (unrolled in the BP kernel; rolled in the extended-message kernel, whose body must stay inside the I-cache)
#pragma unroll(EXT ? 1 : 3)
    for (int r = 1; r < 4; ++r) {
      const int a2 = (la + r) & 3;
      p2 tt[16];
      lds_tile(tt, Ts + a2 * kSlice);
      const p2 mm = *reinterpret_cast<const p2*>(m0row + a2 * 8);
      const p2 im = x2::mul2(x2::swap(mm), rot);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 tf = x2::unpk(tt[i]);
        U0[i] = x2::fma2s(tf.x, mm, U0[i]);
        U0[i] = x2::fma2s(tf.y, im, U0[i]);
      }
    }
    __syncwarp();                                           // every lane is done with the T slices
    unsigned char* Xs = st + (s * 8 + p * 4) * kSlice;      // exchange area: U1 slices in the T layout
    sts_tile(Xs + la * kSlice, U1);

    float2 e[3][4];                                         // per message k: [g0 piece (2), g1 piece (2)]; BP sums them
    float tr[3];                                            // traces (real: the outputs are Hermitian), all-reduced over the
                                                            // node's 8 lanes from the lanes' partial diagonals -- three short
                                                            // shuffle chains that overlap the arithmetic instead of ending it
    unsigned char* red = scratch + s * kRedSlot;
    // ---- out_1[x][y] = sum_{a,c} conj(U2[a][x][c]) U0[a][y][c],  out_2[x][y] = sum_{a,b} conj(U1[a][b][x]) U0[a][b][y]
    // (partial over this lane's (p, a); Hermitian: diagonal real parts and the upper triangle only)
#pragma unroll
    for (int k = 1; k <= 2; ++k) {
      float2 acc[10];
#pragma unroll
      for (int x = 0; x < 4; ++x) {                         // diagonal: sum of element-wise pair products
        p2 d2 = x2::mul2((k == 1) ? U2[x * 4] : U1[x], (k == 1) ? U0[x * 4] : U0[x]);
#pragma unroll
        for (int qq = 1; qq < 4; ++qq)
          d2 = x2::fma2((k == 1) ? U2[x * 4 + qq] : U1[qq * 4 + x], (k == 1) ? U0[x * 4 + qq] : U0[qq * 4 + x], d2);
        acc[x] = make_float2(x2::hsum(d2), 0.f);
      }
      {
        int n = 4;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = x + 1; y < 4; ++y) {
            CAcc v;
            cmac<true>(v, (k == 1) ? U2[x * 4] : U1[x], (k == 1) ? U0[y * 4] : U0[y]);
#pragma unroll
            for (int qq = 1; qq < 4; ++qq)
              cmac<false>(v, (k == 1) ? U2[x * 4 + qq] : U1[qq * 4 + x], (k == 1) ? U0[y * 4 + qq] : U0[qq * 4 + y]);
            acc[n++] = cfinish_conj(v);
          }
      }
      tr[k] = allreduce8((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x));
      if (!EXT) {
        // reduce-scatter over the node's 8 lanes (p and a): 14 shuffles, then one pair from the partner lane
        float z[16] = {acc[4].x, acc[4].y, acc[5].x, acc[5].y, acc[0].x, acc[1].x, acc[8].x, acc[8].y,
                       acc[7].x, acc[7].y, acc[9].x, acc[9].y, acc[6].x, acc[6].y, acc[2].x, acc[3].x};
        const bool b2 = t & 4, b1 = t & 2, b0 = t & 1;
        float y8[8], y4[4], y2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          y8[i] = (b2 ? z[8 + i] : z[i]) + __shfl_xor_sync(0xffffffffu, b2 ? z[i] : z[8 + i], 4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          y4[i] = (b1 ? y8[4 + i] : y8[i]) + __shfl_xor_sync(0xffffffffu, b1 ? y8[i] : y8[4 + i], 2);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          y2[i] = (b0 ? y4[2 + i] : y4[i]) + __shfl_xor_sync(0xffffffffu, b0 ? y4[i] : y4[2 + i], 1);
        const float gx = __shfl_sync(0xffffffffu, y2[0], partner), gy = __shfl_sync(0xffffffffu, y2[1], partner);
        const float ax = a_got ? gx : y2[0], ay = a_got ? gy : y2[1];
        const float bx = b_got ? gx : y2[0], by = b_got ? gy : y2[1];
        e[k][0] = make_float2(ax, a_im * ay);
        e[k][1] = make_float2(b_usey ? by : bx, b_im * by);
        e[k][2] = make_float2(0.f, 0.f);
        e[k][3] = make_float2(0.f, 0.f);
      } else {
        __syncwarp();                                       // previous readers of the scratch are done
#pragma unroll
        for (int i = 0; i < 5; ++i)
          *reinterpret_cast<float4*>(red + t * kRedRow + 16 * i) =
              make_float4(acc[2 * i].x, acc[2 * i].y, acc[2 * i + 1].x, acc[2 * i + 1].y);
        __syncwarp();
        float2 g0a = make_float2(0.f, 0.f), g0b = g0a, g1a = g0a, g1b = g0a;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 a0 = *reinterpret_cast<const float2*>(red + r * kRedRow + offA);
          const float2 b0 = *reinterpret_cast<const float2*>(red + r * kRedRow + offB);
          const float2 a1 = *reinterpret_cast<const float2*>(red + (4 + r) * kRedRow + offA);
          const float2 b1 = *reinterpret_cast<const float2*>(red + (4 + r) * kRedRow + offB);
          g0a.x += a0.x; g0a.y += a0.y; g0b.x += b0.x; g0b.y += b0.y;
          g1a.x += a1.x; g1a.y += a1.y; g1b.x += b1.x; g1b.y += b1.y;
        }
        e[k][0] = make_float2(g0a.x, sgnA * g0a.y); e[k][1] = make_float2(g0b.x, sgnB * g0b.y);
        e[k][2] = make_float2(g1a.x, sgnA * g1a.y); e[k][3] = make_float2(g1b.x, sgnB * g1b.y);
      }
    }
    // ---- out_0[x][y] = sum_{b,c} conj(U1[x][b][c]) U2[y][b][c]: lane (p, y = la) against the exchanged U1 slices.
    // Hermitian: the lane computes (la, la), (la + 1, la) and -- lanes la < 2 only keep it -- (la + 2, la); the mirrored
    // entries are stored as conjugates, (la + 3, la) is the mirror of the next lane's (la + 1, la).
    {
      unsigned char* o0 = scratch + kOut0 + s * 64;         // [x][node][p][y]: a warp-wide store is 256 contiguous bytes
      {                                                     // x = la: this lane's own U1 slice is still in registers
        p2 d0 = x2::mul2(U1[0], U2[0]), d1 = x2::mul2(U1[1], U2[1]);     // real part only: element-wise pair products
#pragma unroll
        for (int i = 2; i < 16; i += 2) { d0 = x2::fma2(U1[i], U2[i], d0); d1 = x2::fma2(U1[i + 1], U2[i + 1], d1); }
        const float dg = x2::hsum(d0) + x2::hsum(d1);
        tr[0] = allreduce8(dg);
        *reinterpret_cast<float2*>(o0 + la * kOut0Row + p * 32 + la * 8) = make_float2(dg, 0.f);
      }
      __syncwarp();                                         // the U1 slices of the other lanes are visible
#pragma unroll(EXT ? 1 : 2)
      for (int r = 1; r < 3; ++r) {                         // the 8 lanes of a node read 8 distinct slices
        const int x = (la + r) & 3;
        p2 ux[16];
        lds_tile(ux, Xs + x * kSlice);
        CAcc v0, v1;                                        // two chains
        cmac<true>(v0, ux[0], U2[0]);
        cmac<true>(v1, ux[1], U2[1]);
#pragma unroll
        for (int i = 2; i < 16; i += 2) { cmac<false>(v0, ux[i], U2[i]); cmac<false>(v1, ux[i + 1], U2[i + 1]); }
        const float2 r0 = cfinish_conj(v0), r1 = cfinish_conj(v1);
        const float2 v = make_float2(r0.x + r1.x, r0.y + r1.y);
        if (r == 1 || la < 2) {
          *reinterpret_cast<float2*>(o0 + x * kOut0Row + p * 32 + la * 8) = v;
          *reinterpret_cast<float2*>(o0 + la * kOut0Row + p * 32 + x * 8) = make_float2(v.x, -v.y);
        }
      }
      __syncwarp();
      const float4 v0 = *reinterpret_cast<const float4*>(o0 + (t >> 1) * kOut0Row + (t & 1) * 16);
      const float4 v1 = *reinterpret_cast<const float4*>(o0 + (t >> 1) * kOut0Row + 32 + (t & 1) * 16);
      e[0][0] = make_float2(v0.x, v0.y); e[0][1] = make_float2(v0.z, v0.w);
      e[0][2] = make_float2(v1.x, v1.y); e[0][3] = make_float2(v1.z, v1.w);
    }

    // ---- epilogue: lane t owns elements (x, y0) and (x, y0 + 1) of every message, x = t / 2, y0 = 2 (t % 2)
    const int x = t >> 1, y0 = (t & 1) * 2;
    unsigned char* const out_base = reinterpret_cast<unsigned char*>(msgs_out) + t * 16;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float2 sa = make_float2(e[k][0].x + e[k][2].x, e[k][0].y + e[k][2].y);
      const float2 sb = make_float2(e[k][1].x + e[k][3].x, e[k][1].y + e[k][3].y);
      const unsigned slot = (unsigned)__shfl_sync(0xffffffffu, idx_cur, 3 + k, 8);
      unsigned char* far = nullptr;
      if (multi) {
        const int rp = __shfl_sync(0xffffffffu, rp_cur, k, 8);
        if (rp >= 0) far = peers(rp >> 27) + (size_t)(rp & ((1 << 27) - 1)) * (EXT ? 512 : 128) + (EXT ? 0 : t * 16);
      }
      if (!EXT) {
        const float d = rcp_approx(tr[k]);
        const float2 na = make_float2(__fmul_rn(sa.x, d), __fmul_rn(sa.y, d)), nb = make_float2(__fmul_rn(sb.x, d), __fmul_rn(sb.y, d));
        const float4 ov = *reinterpret_cast<const float4*>(st + kTBytes + kMBytes + (k * 4 + s) * kMsg + t * 16);
        // (explicit fused / unfused operations: the result must not depend on how an instantiation was contracted)
        mnum = fmaxf(mnum, fmaxf(abs2_rn(na.x - ov.x, na.y - ov.y), abs2_rn(nb.x - ov.z, nb.y - ov.w)));
        mden = fmaxf(mden, fmaxf(abs2_rn(na.x + ov.x, na.y + ov.y), abs2_rn(nb.x + ov.z, nb.y + ov.w)));
        // fmaxf drops NaN operands.  Any non-finite entry of the node tensor or of an incoming message reaches the trace
        // (a sum over every contracted entry), so a non-finite trace poisons the residual like np.abs().max() would
        if (!(fabsf(tr[k]) < INFINITY)) { mnum = INFINITY; mden = INFINITY; }
        float4 w;
        if (write_undamped) {
          w = make_float4(na.x, na.y, nb.x, nb.y);
        } else {
          const float al = a.damping, be = 1.f - a.damping;
          w = make_float4(fmaf(al, ov.x, __fmul_rn(be, na.x)), fmaf(al, ov.y, __fmul_rn(be, na.y)),
                          fmaf(al, ov.z, __fmul_rn(be, nb.x)), fmaf(al, ov.w, __fmul_rn(be, nb.y)));
        }
        *reinterpret_cast<float4*>(out_base + (size_t)slot * 128) = w;
        if (far) *reinterpret_cast<float4*>(far) = w;      // halo slot on the peer that owns the receiver
      } else {
        // ext[(s1,x),(s2,y)] = conj(f_s1) f_s2 (g0 + (-1)^(s1+s2) g1)[x][y] / ((|f0|^2 + |f1|^2) trace)
        const float2 ff[2] = {make_float2(__shfl_sync(0xffffffffu, zf0.re, k, 8), __shfl_sync(0xffffffffu, zf0.im, k, 8)),
                              make_float2(__shfl_sync(0xffffffffu, zf1.re, k, 8), __shfl_sync(0xffffffffu, zf1.im, k, 8))};
        const float w = (ff[0].x * ff[0].x + ff[0].y * ff[0].y + ff[1].x * ff[1].x + ff[1].y * ff[1].y);
        const float2 itr = make_float2(1.f / (w * tr[k]), 0.f);
        const float2 da = make_float2(e[k][0].x - e[k][2].x, e[k][0].y - e[k][2].y);
        const float2 db = make_float2(e[k][1].x - e[k][3].x, e[k][1].y - e[k][3].y);
        unsigned char* dst = reinterpret_cast<unsigned char*>(msgs_out) + (size_t)slot * 512;
#pragma unroll
        for (int s1 = 0; s1 < 2; ++s1)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const float2 cf = cmul(itr, cmul(make_float2(ff[s1].x, -ff[s1].y), ff[s2]));
            const float2 va = cmul(cf, s1 == s2 ? sa : da), vb = cmul(cf, s1 == s2 ? sb : db);
            const float4 w4 = make_float4(va.x, va.y, vb.x, vb.y);
            *reinterpret_cast<float4*>(dst + ((s1 * 4 + x) * 8 + s2 * 4 + y0) * 8) = w4;
            if (far) *reinterpret_cast<float4*>(far + ((s1 * 4 + x) * 8 + s2 * 4 + y0) * 8) = w4;
          }
      }
    }
    idx_cur = idx_nxt;
    idx_nxt = idx_nn;
    rp_cur = rp_nxt;
    rp_nxt = rp_nn;
    __syncwarp();                                           // stage `cur` may be overwritten by the next issue
    if (g < g_mid && g + nwarps >= g_mid) mid();
  }
  if (!EXT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnum = fmaxf(mnum, __shfl_xor_sync(0xffffffffu, mnum, o));
      mden = fmaxf(mden, __shfl_xor_sync(0xffffffffu, mden, o));
    }
    if (lane == 0) {
      atomicMax(reinterpret_cast<unsigned int*>(a.resid + 2 * it), __float_as_uint(mnum));
      atomicMax(reinterpret_cast<unsigned int*>(a.resid + 2 * it + 1), __float_as_uint(mden));
    }
  }
}

template <bool EXT, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) k_msgs_d3D4(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (!EXT && a.it > 0) {                                   // device-side early exit after convergence
    if (*((volatile int32_t*)a.status) != 0) return;
    const float num = a.resid[2 * (a.it - 1)], den = a.resid[2 * (a.it - 1) + 1];
    if (sqrtf(num / den) < a.bp_eps) {
      if (threadIdx.x == 0) { a.status[1] = a.it; __threadfence(); a.status[0] = 1; }
      return;
    }
  }
  const float2* cur = a.msgs_cur;
  if (EXT && a.sel_status != nullptr) {
    // converged: the input of the converging sweep; cap reached: the output of the last sweep (state.py:118-124)
    const int conv = a.sel_status[0], sw = a.sel_status[1];
    const int idx = (a.sel_parity + (conv ? sw - 1 : a.sel_max_iters)) % a.sel_nbuf;
    cur = idx == 0 ? a.sel_msgs[0] : (idx == 1 ? a.sel_msgs[1] : a.sel_msgs[2]);
  }
  sweep<EXT, MULTI>(a, cur, a.msgs_out, a.it, a.write_undamped, PeersOfArgs{a}, smem);
}

// ---- whole BP run in ONE cooperative launch (reference _run_bp, state.py:97-124) ----------------------------------
// The persistent CTAs (one per SM, all co-resident) iterate: sweep -> grid barrier -> (multi-GPU: CTA 0 pushes the
// residual maxima to the peers and runs the cross-GPU flag barrier, like bqa_sync.cu) -> every CTA reads the global
// residual and stops or goes on.  No kernel launch and no host round trip per sweep.
struct RunArgs {
  Args base;                       // msgs_cur / msgs_out / peers / it / write_undamped are filled per sweep
  float2* msgs[3];                 // message buffers; sweep `it` reads msgs[(parity + it) % nbuf], writes the next one
  unsigned char* peers[3][BQA_MAX_PEERS];
  int nbuf;                        // 2 on one GPU; 3 across GPUs (the convergence test lags one sweep, see k_bp_run_d3D4)
  int boundary_groups;             // the first groups of the class hold every node with a remote out-edge
  int parity, max_iters;
  int rank, world;                 // world == 1: single GPU
  uint4* peer_xchg[BQA_MAX_PEERS]; // every rank's handshake lines [2][BQA_MAX_PEERS] (peer mapped): 64 bytes into its flag buffer
  unsigned seq_base;               // cross-GPU sequence numbers seq_base + 1, + 2, ... (one per sweep)
  long long timeout_cycles;        // a peer that stays silent this long aborts the run (status[3])
  // Fences of the handshake (BQA_B200_FENCE): the sender orders its halo stores before the line with a system-scope
  // fence in every mode.  Receiver, after it has seen the line: 0 = fence.sc.sys, 1 = fence.acq_rel.sys (3.5 us each on
  // B200 with peer mappings: measured, profiles/r2_multigpu.md), 2 (default) = fence.acq_rel.gpu, 0.8 us -- the halo slots
  // and the line are in THIS GPU's memory, read through its L2 (cp.async.cg, ld.volatile), which is where the peer's
  // stores land in the order its fence gave them.
  int fence_mode;
  int zigzag;                        // 1: odd sweeps of the single-GPU run visit the groups in reverse (L2 reuse)
  unsigned long long* trace;       // profiling aid (bqa_b200_set_bp_trace): 5 globaltimer stamps per sweep from CTA 0, or null
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_release_sys(int mode) {
  if (mode == 0) __threadfence_system(); else asm volatile("fence.acq_rel.sys;" ::: "memory");
}
__device__ __forceinline__ void fence_acquire_sys(int mode) {
  if (mode == 0) __threadfence_system();
  else if (mode == 1) asm volatile("fence.acq_rel.sys;" ::: "memory");
  else asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// all CTAs of the grid; `counter` only grows (zeroed by the host before the launch); status[3] != 0 aborts everyone
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned& generation, volatile int32_t* status,
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
    while (ld_acquire_gpu_u32(counter) < target) {
      if (status[3] != 0 || clock64() - t0 > timeout_cycles) { status[3] = 1; ok = 0; break; }
    }
  }
  __syncthreads();
  return ok != 0;
}

struct PeersOfRun {
  const RunArgs& r;
  int buf;
  __device__ __forceinline__ unsigned char* operator()(int q) const { return r.peers[buf][q]; }
};

template <bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) k_bp_run_d3D4(const __grid_constant__ RunArgs r) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Args& a = r.base;
  unsigned* counter = reinterpret_cast<unsigned*>(a.status + 2);
  unsigned generation = 0;
  int sweeps = r.max_iters, converged = 0;
  const bool tracing = r.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (!(MULTI && r.world > 1)) {
    for (int it = 0; it < r.max_iters; ++it) {
      const int cur = (r.parity + it) & 1;
      if (tracing) r.trace[5 * it] = globaltimer_ns();
      // cap reached: the undamped sweep is kept (state.py:122-123)
      sweep<false, MULTI>(a, r.msgs[cur], r.msgs[cur ^ 1], it, it == r.max_iters - 1, PeersOfRun{r, cur ^ 1}, smem, 0,
                          0x7fffffff, (it & r.zigzag) != 0);
      if (tracing) r.trace[5 * it + 1] = globaltimer_ns();
      if (!grid_barrier(counter, generation, a.status, r.timeout_cycles)) return;
      if (tracing) { r.trace[5 * it + 2] = globaltimer_ns(); r.trace[5 * it + 3] = r.trace[5 * it + 4] = r.trace[5 * it + 2]; }
      const float num = __ldcg(a.resid + 2 * it), den = __ldcg(a.resid + 2 * it + 1);
      if (sqrtf(num / den) < a.bp_eps) { sweeps = it + 1; converged = 1; break; }
    }
  } else {
    // ---- across GPUs -------------------------------------------------------------------------------------------------
    // A sweep needs the peers' halo messages of the previous sweep, and the convergence test needs the peers' residual
    // maxima (get_dist is a ratio of two GLOBAL maxima, backends.py:492-495).  Both travel as 16-byte lines in the
    // layout of NCCL's low-latency protocol (two 8-byte words, each carrying data and the sequence number: only 8-byte
    // stores are atomic over NVLink), and both are taken off the critical path:
    //   * the nodes with a remote out-edge come FIRST in the class.  The CTA that is the last to finish its boundary
    //     groups fences (system scope: every CTA fenced its own halo stores before it arrived) and sends the DATA line
    //     of the sweep; the interior groups and the grid barrier run while it is in flight;
    //   * the RESID line {max |new - old|^2, seq, max |new + old|^2, seq} is sent after the grid barrier and read ONE
    //     SWEEP LATER: sweep it + 1 starts without knowing whether sweep it converged.  The sweeps rotate through THREE
    //     message buffers, so the input of the converging sweep -- what the reference keeps (state.py:118-120) -- is
    //     still intact when the test arrives; the price is one discarded sweep per BP run.
    // No remote atomics, no second grid barrier, nothing of a peer's control block is written.
    unsigned* bcount = reinterpret_cast<unsigned*>(a.resid + 2 * r.max_iters);     // zeroed with the residuals
    const int groups = (int)((a.B + 3) >> 2);
    const int gb = min(r.boundary_groups, groups);
    __shared__ unsigned s_max[2];
    __shared__ unsigned s_bdone;                             // warps of this CTA that finished their boundary groups (all sweeps)
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_bdone = 0;
    __syncthreads();
    auto fold_resid = [&](int k) -> bool {                   // global residual of sweep k; true = converged
      const unsigned seq = r.seq_base + k + 1;
      if (threadIdx.x == 0) {
        s_max[0] = __float_as_uint(__ldcg(a.resid + 2 * k));
        s_max[1] = __float_as_uint(__ldcg(a.resid + 2 * k + 1));
        s_abort = 0;
      }
      __syncthreads();
      if (threadIdx.x < r.world && (int)threadIdx.x != r.rank) {
        const uint4* mine = r.peer_xchg[r.rank] + (4 + (k & 3)) * BQA_MAX_PEERS + threadIdx.x;
        uint4 line = ld_volatile_v4(mine);
        const long long t0 = clock64();
        while (line.y != seq || line.w != seq) {
          if (*((volatile int32_t*)a.status + 3) != 0 || clock64() - t0 > r.timeout_cycles) { a.status[3] = 1; s_abort = 1; break; }
          line = ld_volatile_v4(mine);
        }
        atomicMax(&s_max[0], line.x);                       // non-negative reals order like their bit patterns
        atomicMax(&s_max[1], line.z);
      }
      __syncthreads();
      const float num = __uint_as_float(s_max[0]), den = __uint_as_float(s_max[1]);
      if (blockIdx.x == 0 && threadIdx.x == 0) { a.resid[2 * k] = num; a.resid[2 * k + 1] = den; }   // the host reads the GLOBAL one
      return sqrtf(num / den) < a.bp_eps;
    };
    for (int it = 0; it < r.max_iters; ++it) {
      const int cur = (r.parity + it) % r.nbuf, nxt = (cur + 1) % r.nbuf;
      const unsigned seq = r.seq_base + it + 1;
      const int undamped = it == r.max_iters - 1;
      if (tracing) r.trace[5 * it] = globaltimer_ns();
      // 1. boundary groups; the last warp of the last CTA to finish them sends the DATA lines.  No CTA-wide barrier and
      // one system-scope fence per rank and sweep: every warp releases its halo stores at gpu scope (fence + counter),
      // the warp that completes the count fences at system scope -- cumulativity carries the other warps' stores -- and
      // stores the lines; everybody else is already sweeping interior groups.
      // ONE pass over all groups (boundary groups come first in the node order): the copies of a warp's first interior
      // group are in flight while it finishes its last boundary group and signals
      auto boundary_done = [&]() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
          __threadfence();
          if (atomicAdd(&s_bdone, 1u) == (unsigned)(it + 1) * kWarps - 1u) {
            __threadfence();
            if (atomicAdd(bcount, 1u) == (unsigned)(it + 1) * gridDim.x - 1u) {
              fence_release_sys(r.fence_mode);
              for (int q = 0; q < r.world; ++q)
                if (q != r.rank) st_volatile_v4(r.peer_xchg[q] + (it & 3) * BQA_MAX_PEERS + r.rank, make_uint4(seq, seq, seq, seq));
            }
          }
          if (tracing) r.trace[5 * it + 1] = globaltimer_ns();
        }
      };
      // 2. interior groups, grid barrier (local messages and local residual of the sweep complete)
      sweep<false, MULTI>(a, r.msgs[cur], r.msgs[nxt], it, undamped, PeersOfRun{r, nxt}, smem, 0, groups, false, gb, boundary_done);
      if (!grid_barrier(counter, generation, a.status, r.timeout_cycles)) return;
      if (tracing) r.trace[5 * it + 2] = globaltimer_ns();
      // 3. RESID line of this sweep (read by the peers one sweep later)
      if (blockIdx.x == 0 && threadIdx.x < r.world && (int)threadIdx.x != r.rank) {
        const unsigned num = __float_as_uint(__ldcg(a.resid + 2 * it)), den = __float_as_uint(__ldcg(a.resid + 2 * it + 1));
        st_volatile_v4(r.peer_xchg[threadIdx.x] + (4 + (it & 3)) * BQA_MAX_PEERS + r.rank, make_uint4(num, seq, den, seq));
      }
      if (tracing) r.trace[5 * it + 3] = globaltimer_ns();
      // 4. the peers' halo messages of this sweep must have landed before the next sweep reads them
      if (threadIdx.x < r.world && (int)threadIdx.x != r.rank) {
        const uint4* mine = r.peer_xchg[r.rank] + (it & 3) * BQA_MAX_PEERS + threadIdx.x;
        uint4 line = ld_volatile_v4(mine);
        const long long t0 = clock64();
        while (line.x != seq || line.w != seq) {
          if (*((volatile int32_t*)a.status + 3) != 0 || clock64() - t0 > r.timeout_cycles) { a.status[3] = 1; break; }
          line = ld_volatile_v4(mine);
        }
        fence_acquire_sys(r.fence_mode);                    // the peer's halo stores are ordered before its line
      }
      __syncthreads();
      if (tracing) r.trace[5 * it + 4] = globaltimer_ns();
      if (*((volatile int32_t*)a.status + 3) != 0) return;
      // 5. convergence test of the PREVIOUS sweep
      if (it >= 1 && fold_resid(it - 1)) { sweeps = it; converged = 1; break; }
      if (*((volatile int32_t*)a.status + 3) != 0) return;
    }
    if (!converged) {                                        // the last sweep's own test (cap reached or not)
      if (fold_resid(r.max_iters - 1)) { sweeps = r.max_iters; converged = 1; }
      if (*((volatile int32_t*)a.status + 3) != 0) return;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.status[1] = sweeps; a.status[0] = converged; }
}

static unsigned long long* g_bp_trace = nullptr;            // device buffer of 5 * max_iters stamps, or null
void set_bp_trace(void* p) { g_bp_trace = (unsigned long long*)p; }
static long long g_timeout_cycles = 20000000000LL;          // ~10 s at 2 GHz
long long barrier_timeout_cycles() { return g_timeout_cycles; }
void set_barrier_timeout_cycles(long long c) { g_timeout_cycles = c > 0 ? c : 20000000000LL; }

static int sm_count() {                     // of the current device (a process may drive several)
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace fast

int launch_fast_bp_run_d3D4(long long B, const void* T, void* msgs0, void* msgs1, int parity, const int32_t* in_pos,
                            const int32_t* out_pos, double damping, double bp_eps, int max_iters, void* resid,
                            int32_t* status, const int32_t* remote_pos, void* const* peers0, void* const* peers1, int rank,
                            int world, void* const* peer_resid, void* const* peer_flags, unsigned seq_base,
                            void* msgs2, void* const* peers2, long long boundary_nodes, cudaStream_t st) {
  using namespace fast;
  if (B == 0) return 0;
  if (world < 1 || world > BQA_MAX_PEERS) return set_error("bp_run: world %d outside [1, %d]", world, BQA_MAX_PEERS);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_bp_run_d3D4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_bp_run_d3D4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(k_bp_run_d3D4): %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  RunArgs r{};
  Args& a = r.base;
  a.B = B; a.T = (const float2*)T; a.in_pos = in_pos; a.out_pos = out_pos;
  a.damping = (float)damping; a.bp_eps = (float)bp_eps; a.resid = (float*)resid; a.status = status;
  a.remote_pos = (world > 1 && peers0 && peers1) ? remote_pos : nullptr;
  r.msgs[0] = (float2*)msgs0; r.msgs[1] = (float2*)msgs1; r.msgs[2] = (float2*)msgs2;
  r.nbuf = world > 1 ? 3 : 2;
  if (world > 1 && (!msgs2 || !peers2)) return set_error("bp_run: the multi-GPU run needs a third message buffer");
  if (boundary_nodes < 0 || boundary_nodes > B) return set_error("bp_run: %lld boundary nodes of %lld", boundary_nodes, B);
  r.boundary_groups = (int)((boundary_nodes + 3) / 4);
  r.parity = world > 1 ? ((parity % 3) + 3) % 3 : (parity & 1); r.max_iters = max_iters; r.rank = rank; r.world = world; r.seq_base = seq_base;
  for (int q = 0; q < BQA_MAX_PEERS; ++q) {
    r.peers[0][q] = (a.remote_pos && peers0) ? (unsigned char*)peers0[q] : nullptr;
    r.peers[1][q] = (a.remote_pos && peers1) ? (unsigned char*)peers1[q] : nullptr;
    r.peers[2][q] = (a.remote_pos && peers2) ? (unsigned char*)peers2[q] : nullptr;
    r.peer_xchg[q] = (world > 1 && q < world) ? (uint4*)((unsigned char*)peer_flags[q] + 64) : nullptr;
  }
  (void)peer_resid;
  r.timeout_cycles = barrier_timeout_cycles();
  r.trace = g_bp_trace;
  static const int fence_mode = [] { const char* e = getenv("BQA_B200_FENCE"); return e ? atoi(e) : 2; }();
  r.fence_mode = fence_mode;
  static const int zigzag = [] { const char* e = getenv("BQA_B200_BP_ZIGZAG"); return e ? atoi(e) & 1 : 1; }();
  r.zigzag = zigzag;
  const long long groups = (B + 3) / 4;
  long long grid = (groups + kWarps - 1) / kWarps;
  if (grid > sm_count()) grid = sm_count();
  void* params[] = {&r};
  const bool multi = world > 1 || a.remote_pos != nullptr;
  const void* fn = multi ? (const void*)k_bp_run_d3D4<true> : (const void*)k_bp_run_d3D4<false>;
  // (pinning T in L2 with an access-policy window was tried in r2: no change at 48 MB, see profiles/r2_bp_l2_experiments.md)
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kThreads), params, (size_t)kSmem, st);
  if (e != cudaSuccess) return set_error("cudaLaunchCooperativeKernel(k_bp_run_d3D4): %s", cudaGetErrorString(e));
  return after_launch("bp_run(d3D4)");
}

// a group is 4 nodes (the last one may overlap its predecessor), node offsets are 32-bit: 4 <= B < 2^29 (or empty)
bool fast_d3D4_available(int prec, int degree, int D, long long B) {
  return prec == 0 && degree == 3 && D == 4 && (B == 0 || (B >= 4 && B < (1LL << 29)));
}

int launch_fast_msgs_d3D4(bool ext, long long B, const void* T, const void* msgs_cur, void* msgs_out,
                          const int32_t* in_pos, const int32_t* out_pos, const void* edge_ampls, double ztime,
                          double damping, int write_undamped, double bp_eps, int it, void* resid, int32_t* status,
                          const int32_t* remote_pos, void* const* peers, cudaStream_t st, const AfterRun* after) {
  using namespace fast;
  if (B == 0) return 0;
  static bool configured[64] = {};           // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e1 = cudaFuncSetAttribute(k_msgs_d3D4<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaError_t e2 = cudaFuncSetAttribute(k_msgs_d3D4<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_msgs_d3D4<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e1 != cudaSuccess || e2 != cudaSuccess)
      return set_error("cudaFuncSetAttribute(k_msgs_d3D4): %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  Args a{};
  a.B = B; a.T = (const float2*)T; a.msgs_cur = (const float2*)msgs_cur; a.msgs_out = (float2*)msgs_out;
  a.in_pos = in_pos; a.out_pos = out_pos; a.edge_ampls = (const float*)edge_ampls;
  a.ztime = (float)ztime; a.damping = (float)damping; a.bp_eps = (float)bp_eps;
  a.write_undamped = write_undamped; a.it = it; a.resid = (float*)resid; a.status = status;
  a.remote_pos = peers ? remote_pos : nullptr;
  for (int q = 0; q < BQA_MAX_PEERS; ++q) a.peers[q] = peers ? (unsigned char*)peers[q] : nullptr;
  if (after) {
    if (!ext || !after->status || after->nbuf < 2 || after->nbuf > 3 || !after->msgs[0] || !after->msgs[1] ||
        (after->nbuf == 3 && !after->msgs[2]))
      return set_error("ext_msgs_after_run: needs the run's status words and its %d message buffers", after->nbuf);
    a.sel_status = after->status;
    for (int k = 0; k < 3; ++k) a.sel_msgs[k] = (const float2*)after->msgs[k];
    a.sel_nbuf = after->nbuf; a.sel_max_iters = after->max_iters;
    a.sel_parity = ((after->parity % after->nbuf) + after->nbuf) % after->nbuf;
  }
  const long long groups = (B + 3) / 4;
  long long grid = (groups + kWarps - 1) / kWarps;
  if (grid > sm_count()) grid = sm_count();
  const bool multi = a.remote_pos != nullptr;
  if (ext) k_msgs_d3D4<true, true><<<(int)grid, kThreads, kSmem, st>>>(a);      // one instantiation: see sweep()
  else if (multi) k_msgs_d3D4<false, true><<<(int)grid, kThreads, kSmem, st>>>(a);
  else k_msgs_d3D4<false, false><<<(int)grid, kThreads, kSmem, st>>>(a);
  return after_launch(ext ? "ext_msgs(d3D4)" : "bp_sweep(d3D4)");
}

}  // namespace bqa
