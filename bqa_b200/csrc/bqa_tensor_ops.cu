// bqa_tensor_ops.cu -- the raw operations of bqa's backend interface on device arrays.
//
// The reference's plugin boundary is the ABC `bqa.backends.Tensor` (src/bqa/backends.py:28-252): 36 abstract raw
// operations on a backend-owned array, from which the ABC builds every composite (batch_tensordot, pass_msgs, ...).
// `bqa_b200.tensor_backend.B200Backend` implements them on device buffers with the kernels below, so that the
// unmodified engine `bqa.state` runs op by op on the GPU; the fused entry points of include/bqa_b200.h override the hot
// composites.  Every array is dense row-major; complex64 / complex128 interleaved (prec = BQA_C64 / BQA_C128); index
// arrays are int64 like the reference's (backends.py:583).  These are small, launch-bound kernels by design (one per
// raw op, like the CuPy backend's elementwise kernels, backends.py:770-1023): the throughput path is the fused one.
#include <cuda_runtime.h>

#include "../../include/bqa_b200.h"
#include "bqa_core.cuh"
#include "bqa_launch.cuh"

namespace bqa {
namespace tops {

constexpr int kThreads = 256;
static inline int grid_for(long long n) {
  long long g = (n + kThreads - 1) / kThreads;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

struct Nd {                                 // up to 8 dimensions; strides in elements (0 = broadcast, negative = reversed)
  int rank;
  long long shape[8], sa[8], sb[8], so[8];
};
__device__ __forceinline__ void offsets(const Nd& d, long long idx, long long& oa, long long& ob, long long& oo) {
  oa = ob = oo = 0;
#pragma unroll 1
  for (int k = d.rank - 1; k >= 0; --k) {
    const long long q = idx / d.shape[k], i = idx - q * d.shape[k];
    oa += i * d.sa[k]; ob += i * d.sb[k]; oo += i * d.so[k];
    idx = q;
  }
}

// ---- complex helpers -------------------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ cx<R> c_inv(cx<R> a) {
  const R d = R(1) / (a.re * a.re + a.im * a.im);
  return mk<R>(a.re * d, -a.im * d);
}
template <typename R> __device__ __forceinline__ cx<R> c_div(cx<R> a, cx<R> b) { return a * c_inv(b); }
template <typename R> __device__ __forceinline__ cx<R> c_sqrt(cx<R> a) {       // principal branch, like numpy.sqrt
  const R m = msqrt(a.re * a.re + a.im * a.im);
  if (m == R(0)) return mk<R>(0, a.im);
  if (a.re >= R(0)) {
    const R t = msqrt(R(0.5) * (m + a.re));
    return mk<R>(t, a.im / (R(2) * t));
  }
  const R t = msqrt(R(0.5) * (m - a.re));
  return mk<R>(mabs(a.im) / (R(2) * t), a.im < R(0) || (a.im == R(0) && signbit(a.im)) ? -t : t);
}
__device__ __forceinline__ float h_sinh(float x) { return sinhf(x); }
__device__ __forceinline__ double h_sinh(double x) { return sinh(x); }
__device__ __forceinline__ float h_cosh(float x) { return coshf(x); }
__device__ __forceinline__ double h_cosh(double x) { return cosh(x); }

enum UnaryOp { U_INV = 0, U_PINV = 1, U_SQRT = 2, U_SIN = 3, U_COS = 4, U_CONJ = 5 };
enum BinaryOp { B_MUL = 0, B_ADD = 1, B_SUB = 2, B_DIV = 3 };

template <typename R>
__global__ void k_unary(int op, long long n, const cx<R>* __restrict__ a, cx<R>* __restrict__ o, R eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const cx<R> v = a[i];
    cx<R> r;
    switch (op) {
      case U_INV: r = c_inv(v); break;
      // numpy.divide(1, x, where = x > eps) on a complex array compares lexicographically (backends.py:719-727)
      case U_PINV: r = (v.re > eps || (v.re == eps && v.im > R(0))) ? c_inv(v) : mk<R>(0, 0); break;
      case U_SQRT: r = c_sqrt(v); break;
      case U_SIN: r = mk<R>(msin(v.re) * h_cosh(v.im), mcos(v.re) * h_sinh(v.im)); break;
      case U_COS: r = mk<R>(mcos(v.re) * h_cosh(v.im), -msin(v.re) * h_sinh(v.im)); break;
      default: r = mk<R>(v.re, -v.im); break;
    }
    o[i] = r;
  }
}

template <typename R>
__global__ void k_binary(int op, Nd d, long long n, const cx<R>* __restrict__ a, const cx<R>* __restrict__ b,
                         cx<R>* __restrict__ o) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long oa, ob, oo;
    offsets(d, i, oa, ob, oo);
    const cx<R> x = a[oa], y = b[ob];
    cx<R> r;
    switch (op) {
      case B_MUL: r = x * y; break;
      case B_ADD: r = x + y; break;
      case B_SUB: r = x - y; break;
      default: r = c_div(x, y); break;
    }
    o[oo] = r;
  }
}

// strided copy of 4-, 8- or 16-byte elements (float; complex64 / int64 / double; complex128): transpose, slicing,
// concatenation, the reversal of the physical axis, real parts of complex arrays (stride 2 in units of the real type)
template <typename E>
__global__ void k_copy(Nd d, long long n, const E* __restrict__ a, E* __restrict__ o) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long oa, ob, oo;
    offsets(d, i, oa, ob, oo);
    o[oo] = a[oa];
  }
}

// rows of `row` elements: out[i] = in[idx[i]] (gather) or out[idx[i]] = in[i] (scatter)
template <typename E>
__global__ void k_rows(int scatter, long long n_idx, long long row, const long long* __restrict__ idx,
                       const E* __restrict__ in, E* __restrict__ out) {
  const long long n = n_idx * row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row, c = i - r * row;
    if (scatter) out[idx[r] * row + c] = in[i];
    else out[i] = in[idx[r] * row + c];
  }
}

template <typename R>
__global__ void k_fill(long long n, cx<R>* o, R re, R im) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = mk<R>(re, im);
}

// dst = alpha * dst + beta * src
template <typename R>
__global__ void k_axpby(long long n, cx<R>* dst, const cx<R>* __restrict__ src, R alpha, R beta) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const cx<R> d = dst[i], s = src[i];
    dst[i] = mk<R>(alpha * d.re + beta * s.re, alpha * d.im + beta * s.im);
  }
}

template <typename R> __device__ __forceinline__ R block_reduce(R v, bool is_max) {
  __shared__ R sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const R w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (w > v || w != w ? w : v) : v + w;
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : (is_max ? R(0) : R(0));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const R w = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? (w > v || w != w ? w : v) : v + w;
    }
  }
  __syncthreads();
  return v;
}

__device__ __forceinline__ void atomic_max_bits(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_bits(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// out (one complex, zeroed by the caller) = max |a_i|   (np.abs(a).max(), backends.py:621-623; NaN propagates)
template <typename R>
__global__ void k_max_abs(long long n, const cx<R>* __restrict__ a, cx<R>* out) {
  R m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const R v = msqrt(norm2(a[i]));
    m = (v > m || v != v) ? v : m;
  }
  m = block_reduce(m, true);
  if (threadIdx.x == 0) atomic_max_bits(&out->re, m);      // a NaN's bit pattern is above every finite one
}

// out[c] = |max over the batch axis of a[:, c]| with numpy's lexicographic order on complex numbers
// (np.abs(a.max(0)), backends.py:625): one thread per column
template <typename R>
__global__ void k_col_max(long long batch, long long inner, const cx<R>* __restrict__ a, cx<R>* __restrict__ out) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < inner; c += (long long)gridDim.x * blockDim.x) {
    cx<R> best = a[c];
    for (long long b = 1; b < batch; ++b) {
      const cx<R> v = a[b * inner + c];
      if (v.re > best.re || (v.re == best.re && v.im > best.im)) best = v;
    }
    out[c] = mk<R>(msqrt(norm2(best)), 0);
  }
}

// per batch item: L2 norm (mode 0) over `inner` elements, or trace (mode 1) of an n x n matrix (inner = n * n)
template <typename R>
__global__ void k_batch_reduce(int mode, long long batch, long long inner, int n, const cx<R>* __restrict__ a,
                               cx<R>* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < batch; b += nwarps) {
    const cx<R>* p = a + b * inner;
    R re = 0, im = 0;
    if (mode == 0) {
      for (long long i = lane; i < inner; i += 32) re += norm2(p[i]);
    } else {
      for (int i = lane; i < n; i += 32) { re += p[(long long)i * n + i].re; im += p[(long long)i * n + i].im; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) out[b] = mode == 0 ? mk<R>(msqrt(re), 0) : mk<R>(re, im);
  }
}

// out[..., i, j] = a[..., i] * (i == j)
template <typename R>
__global__ void k_diag(long long rows, int n, const cx<R>* __restrict__ a, cx<R>* __restrict__ out) {
  const long long total = rows * n * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ((long long)n * n);
    const int ij = (int)(i - r * n * n), ii = ij / n, jj = ij - ii * n;
    out[i] = ii == jj ? a[r * n + ii] : mk<R>(0, 0);
  }
}

// out[b] = a[b] (m x k) . c[b] (k x n): one thread per output element
template <typename R>
__global__ void k_matmul(long long batch, int m, int k, int n, const cx<R>* __restrict__ a, const cx<R>* __restrict__ c,
                         cx<R>* __restrict__ out) {
  const long long total = batch * m * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ((long long)m * n);
    const int rc = (int)(i - b * m * n), r = rc / n, col = rc - r * n;
    const cx<R>* pa = a + (b * m + r) * k;
    const cx<R>* pc = c + b * k * n + col;
    cx<R> acc = mk<R>(0, 0);
    for (int t = 0; t < k; ++t) cmac(acc, pa[t], pc[(long long)t * n]);
    out[i] = acc;
  }
}

// masked SVD of square matrices, one warp per matrix (one-sided Jacobi of bqa_core.cuh in a per-warp global scratch):
// u (n x n), s (n, stored complex like numpy's astype(NP_DTYPE), backends.py:709-717), vh (n x n), singular values
// descending, entries with s <= pinv_eps zeroed together with their columns of u and rows of vh
template <typename R>
__global__ void k_svd(long long batch, int n, const cx<R>* __restrict__ a, cx<R>* __restrict__ u, cx<R>* __restrict__ s,
                      cx<R>* __restrict__ vh, R pinv_eps, cx<R>* scratch) {
  GroupWarp g;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nn = n * n;
  cx<R>* A = scratch + warp * (2 * nn + 2 * n);
  cx<R>* V = A + nn;
  R* sigma = reinterpret_cast<R*>(V + nn);
  int* order = reinterpret_cast<int*>(sigma + n);
  for (long long b = warp; b < batch; b += nwarps) {
    for (int i = lane; i < nn; i += 32) A[i] = a[b * nn + i];
    g.sync();
    jacobi_svd<R>(g, n, n, A, V, sigma, order);
    for (int i = lane; i < nn; i += 32) {
      const int r = i / n, c = i - r * n;
      const int col = order[c];
      const R sv = sigma[col];
      const bool keep = sv > pinv_eps;
      u[b * nn + i] = keep ? (R(1) / sv) * A[r * n + col] : mk<R>(0, 0);
      // vh[c][r] = conj(V[r][col_c])
      vh[b * nn + (long long)c * n + r] = keep ? conj(V[r * n + col]) : mk<R>(0, 0);
    }
    for (int c = lane; c < n; c += 32) {
      const R sv = sigma[order[c]];
      s[b * n + c] = mk<R>(sv > pinv_eps ? sv : R(0), 0);
    }
    g.sync();
  }
}

// rho (B, 2, 2) from the (x, y, z, p0) rows the fused marginal kernel writes (utils.py:23-27 inverted)
template <typename R>
__global__ void k_bloch_to_rho(long long B, const R* __restrict__ bloch, cx<R>* __restrict__ rho) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    const R x = bloch[4 * i], y = bloch[4 * i + 1], z = bloch[4 * i + 2];
    rho[4 * i] = mk<R>(R(0.5) * (R(1) + z), 0);
    rho[4 * i + 1] = mk<R>(R(0.5) * x, R(-0.5) * y);
    rho[4 * i + 2] = mk<R>(R(0.5) * x, R(0.5) * y);
    rho[4 * i + 3] = mk<R>(R(0.5) * (R(1) - z), 0);
  }
}

static int fill_nd(Nd& d, int rank, const long long* shape, const long long* sa, const long long* sb, const long long* so) {
  if (rank < 0 || rank > 8) return set_error("tensor rank %d outside [0, 8]", rank);
  d.rank = rank;
  for (int k = 0; k < 8; ++k) {
    d.shape[k] = k < rank ? shape[k] : 1;
    d.sa[k] = (k < rank && sa) ? sa[k] : 0;
    d.sb[k] = (k < rank && sb) ? sb[k] : 0;
    d.so[k] = (k < rank && so) ? so[k] : 0;
  }
  return 0;
}
static long long numel(int rank, const long long* shape) {
  long long n = 1;
  for (int k = 0; k < rank; ++k) n *= shape[k];
  return n;
}

}  // namespace tops
}  // namespace bqa

using namespace bqa;
using namespace bqa::tops;

#define BQA_PREC_DISPATCH(prec, CALL_F, CALL_D)                                   \
  if ((prec) == BQA_C64) { CALL_F; } else if ((prec) == BQA_C128) { CALL_D; }     \
  else return set_error("unknown precision code %d", (prec));

extern "C" {

int bqa_b200_t_unary(int prec, int op, long long n, const void* a, void* out, void* stream) {
  if (op < 0 || op > 5) return set_error("unknown unary op %d", op);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec,
      (k_unary<float><<<grid_for(n), kThreads, 0, st>>>(op, n, (const cx<float>*)a, (cx<float>*)out, 1.1920929e-07f)),
      (k_unary<double><<<grid_for(n), kThreads, 0, st>>>(op, n, (const cx<double>*)a, (cx<double>*)out, 2.220446049250313e-16)))
  return after_launch("t_unary");
}

int bqa_b200_t_binary(int prec, int op, int rank, const long long* shape, const long long* strides_a,
                      const long long* strides_b, const void* a, const void* b, void* out, void* stream) {
  if (op < 0 || op > 3) return set_error("unknown binary op %d", op);
  Nd d;
  if (int rc = fill_nd(d, rank, shape, strides_a, strides_b, nullptr)) return rc;
  long long so = 1;                                       // dense output
  for (int k = rank - 1; k >= 0; --k) { d.so[k] = so; so *= d.shape[k]; }
  const long long n = numel(rank, shape);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec,
      (k_binary<float><<<grid_for(n), kThreads, 0, st>>>(op, d, n, (const cx<float>*)a, (const cx<float>*)b, (cx<float>*)out)),
      (k_binary<double><<<grid_for(n), kThreads, 0, st>>>(op, d, n, (const cx<double>*)a, (const cx<double>*)b, (cx<double>*)out)))
  return after_launch("t_binary");
}

int bqa_b200_t_copy(int elem_bytes, int rank, const long long* shape, const long long* strides_in,
                    const long long* strides_out, const void* in, void* out, void* stream) {
  Nd d;
  if (int rc = fill_nd(d, rank, shape, strides_in, nullptr, strides_out)) return rc;
  const long long n = numel(rank, shape);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 8) k_copy<unsigned long long><<<grid_for(n), kThreads, 0, st>>>(d, n, (const unsigned long long*)in, (unsigned long long*)out);
  else if (elem_bytes == 16) k_copy<ulonglong2><<<grid_for(n), kThreads, 0, st>>>(d, n, (const ulonglong2*)in, (ulonglong2*)out);
  else if (elem_bytes == 4) k_copy<unsigned int><<<grid_for(n), kThreads, 0, st>>>(d, n, (const unsigned int*)in, (unsigned int*)out);
  else return set_error("t_copy: element size %d is not 4, 8 or 16 bytes", elem_bytes);
  return after_launch("t_copy");
}

int bqa_b200_t_rows(int elem_bytes, int scatter, long long n_idx, long long row_elems, const long long* idx,
                    const void* in, void* out, void* stream) {
  const long long n = n_idx * row_elems;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 8) k_rows<unsigned long long><<<grid_for(n), kThreads, 0, st>>>(scatter, n_idx, row_elems, idx, (const unsigned long long*)in, (unsigned long long*)out);
  else if (elem_bytes == 16) k_rows<ulonglong2><<<grid_for(n), kThreads, 0, st>>>(scatter, n_idx, row_elems, idx, (const ulonglong2*)in, (ulonglong2*)out);
  else return set_error("t_rows: element size %d is not 8 or 16 bytes", elem_bytes);
  return after_launch("t_rows");
}

int bqa_b200_t_fill(int prec, long long n, void* out, double re, double im, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_fill<float><<<grid_for(n), kThreads, 0, st>>>(n, (cx<float>*)out, (float)re, (float)im)),
                    (k_fill<double><<<grid_for(n), kThreads, 0, st>>>(n, (cx<double>*)out, re, im)))
  return after_launch("t_fill");
}

int bqa_b200_t_axpby(int prec, long long n, void* dst, const void* src, double alpha, double beta, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_axpby<float><<<grid_for(n), kThreads, 0, st>>>(n, (cx<float>*)dst, (const cx<float>*)src, (float)alpha, (float)beta)),
                    (k_axpby<double><<<grid_for(n), kThreads, 0, st>>>(n, (cx<double>*)dst, (const cx<double>*)src, alpha, beta)))
  return after_launch("t_axpby");
}

int bqa_b200_t_max_abs(int prec, long long n, const void* a, void* out1, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = prec == BQA_C64 ? 8 : 16;
  cudaMemsetAsync(out1, 0, esz, st);
  if (n <= 0) return 0;
  BQA_PREC_DISPATCH(prec, (k_max_abs<float><<<grid_for(n), kThreads, 0, st>>>(n, (const cx<float>*)a, (cx<float>*)out1)),
                    (k_max_abs<double><<<grid_for(n), kThreads, 0, st>>>(n, (const cx<double>*)a, (cx<double>*)out1)))
  return after_launch("t_max_abs");
}

int bqa_b200_t_col_max(int prec, long long batch, long long inner, const void* a, void* out, void* stream) {
  if (batch <= 0 || inner <= 0) return set_error("t_col_max: empty array");
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_col_max<float><<<grid_for(inner), kThreads, 0, st>>>(batch, inner, (const cx<float>*)a, (cx<float>*)out)),
                    (k_col_max<double><<<grid_for(inner), kThreads, 0, st>>>(batch, inner, (const cx<double>*)a, (cx<double>*)out)))
  return after_launch("t_col_max");
}

int bqa_b200_t_batch_reduce(int prec, int mode, long long batch, long long inner, int n, const void* a, void* out,
                            void* stream) {
  if (mode != 0 && mode != 1) return set_error("t_batch_reduce: mode %d is not 0 (l2 norm) or 1 (trace)", mode);
  if (batch <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(batch * 32);
  BQA_PREC_DISPATCH(prec, (k_batch_reduce<float><<<grid, kThreads, 0, st>>>(mode, batch, inner, n, (const cx<float>*)a, (cx<float>*)out)),
                    (k_batch_reduce<double><<<grid, kThreads, 0, st>>>(mode, batch, inner, n, (const cx<double>*)a, (cx<double>*)out)))
  return after_launch("t_batch_reduce");
}

int bqa_b200_t_diag(int prec, long long rows, int n, const void* a, void* out, void* stream) {
  const long long total = rows * n * n;
  if (total <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_diag<float><<<grid_for(total), kThreads, 0, st>>>(rows, n, (const cx<float>*)a, (cx<float>*)out)),
                    (k_diag<double><<<grid_for(total), kThreads, 0, st>>>(rows, n, (const cx<double>*)a, (cx<double>*)out)))
  return after_launch("t_diag");
}

int bqa_b200_t_matmul(int prec, long long batch, int m, int k, int n, const void* a, const void* b, void* out,
                      void* stream) {
  const long long total = batch * m * n;
  if (total <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_matmul<float><<<grid_for(total), kThreads, 0, st>>>(batch, m, k, n, (const cx<float>*)a, (const cx<float>*)b, (cx<float>*)out)),
                    (k_matmul<double><<<grid_for(total), kThreads, 0, st>>>(batch, m, k, n, (const cx<double>*)a, (const cx<double>*)b, (cx<double>*)out)))
  return after_launch("t_matmul");
}

size_t bqa_b200_t_svd_scratch_bytes(int prec, int n) {
  const size_t esz = prec == BQA_C64 ? 8 : 16;
  return (size_t)(2 * n * n + 2 * n) * esz * BQA_GENERIC_MAX_WARPS;
}

int bqa_b200_t_svd(int prec, long long batch, int n, const void* a, void* u, void* s, void* vh, double pinv_eps,
                   void* scratch, size_t scratch_bytes, void* stream) {
  if (n < 1 || n > BQA_MAX_N) return set_error("t_svd: matrix size %d outside [1, %d]", n, BQA_MAX_N);
  if (batch <= 0) return 0;
  if (scratch_bytes < bqa_b200_t_svd_scratch_bytes(prec, n)) return set_error("t_svd: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  long long warps = batch < BQA_GENERIC_MAX_WARPS ? batch : BQA_GENERIC_MAX_WARPS;
  const int grid = (int)((warps * 32 + kThreads - 1) / kThreads);
  BQA_PREC_DISPATCH(prec, (k_svd<float><<<grid, kThreads, 0, st>>>(batch, n, (const cx<float>*)a, (cx<float>*)u, (cx<float>*)s, (cx<float>*)vh, (float)pinv_eps, (cx<float>*)scratch)),
                    (k_svd<double><<<grid, kThreads, 0, st>>>(batch, n, (const cx<double>*)a, (cx<double>*)u, (cx<double>*)s, (cx<double>*)vh, pinv_eps, (cx<double>*)scratch)))
  return after_launch("t_svd");
}

int bqa_b200_t_bloch_to_rho(int prec, long long B, const void* bloch, void* rho, void* stream) {
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BQA_PREC_DISPATCH(prec, (k_bloch_to_rho<float><<<grid_for(B), kThreads, 0, st>>>(B, (const float*)bloch, (cx<float>*)rho)),
                    (k_bloch_to_rho<double><<<grid_for(B), kThreads, 0, st>>>(B, (const double*)bloch, (cx<double>*)rho)))
  return after_launch("t_bloch_to_rho");
}

}  // extern "C"
