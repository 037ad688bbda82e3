// bn.cu — batch norm over the rows of F [N, C] fused with ReLU and the residual add of a
// BasicBlock, forward and backward.  Pure HBM-bound passes: every kernel reads each operand
// once with 128-bit (fp32) / 64-bit (bf16) vector accesses and keeps per-channel reductions in
// registers -> shared memory -> one fp64 atomic per channel per block.
#include <cooperative_groups.h>
#include <mutex>
#include "common.cuh"

namespace gcd {
namespace {
constexpr int kRedX = 32, kRedY = 8;        // reduction kernels: 32 channel lanes x 8 row lanes
constexpr int kMaxChanIter = 32;            // supports up to 1024 channels (MinkUNet50/101 bottlenecks at tensor stride 16: 256 x 4)
constexpr int kMaxChannels = kRedX * kMaxChanIter;
constexpr int kFusedBwdMaxChannels = 512;      // two-phase backward kernel: six per-channel arrays in shared memory next to the reduction scratch

// Folded scale / shift of a training-mode batch norm, with explicit roundings: the backward pass of conv -> BN -> ReLU units
// re-derives the ReLU mask from x as (x * scale + shift > 0) instead of reading the stored activation, which is only the
// mask the forward pass applied if both sides compute scale and shift to the same bits.
__device__ __forceinline__ float bn_scale_of(float g, float is) { return __fmul_rn(g, is); }
__device__ __forceinline__ float bn_shift_of(float b, float m, float g, float is) { return __fmaf_rn(-__fmul_rn(m, g), is, b); }
// dx of the batch-norm backward, likewise with explicit roundings: the two-phase kernels (two template variants) and the
// stand-alone pass inline this into different surroundings, and the fma contraction ptxas chooses must not differ between them
// (the C-sequenced blocks and the per-launch path are held bit-identical in bf16, tests/test_gpu_fused_block.py)
__device__ __forceinline__ float bn_dx_of(float g, float x, float k, float mean, float is, float sg, float sgx) {
  const float xhat = __fmul_rn(__fsub_rn(x, mean), is);
  return __fmul_rn(k, __fsub_rn(__fsub_rn(g, sg), __fmul_rn(xhat, sgx)));
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t; t.x = *reinterpret_cast<unsigned*>(&a); t.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// Generic two-quantity column reduction.  F(row, ch) -> (a, b); sums[ch] += a, sums[c + ch] += b.
template <typename F>
__device__ __forceinline__ void column_reduce2(int64_t n, int c, double* __restrict__ sums, F f) {
  __shared__ float red_a[kRedY][kRedX + 1], red_b[kRedY][kRedX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int cb = 0; cb < c; cb += kRedX) {
    const int ch = cb + tx;
    float sa = 0.f, sb = 0.f;
    if (ch < c)
      for (int64_t r = (int64_t)blockIdx.x * kRedY + ty; r < n; r += (int64_t)gridDim.x * kRedY) {
        float a, b;
        f(r, ch, a, b);
        sa += a; sb += b;
      }
    red_a[ty][tx] = sa; red_b[ty][tx] = sb;
    __syncthreads();
    if (ty == 0 && ch < c) {
      float ta = 0.f, tb = 0.f;
#pragma unroll
      for (int j = 0; j < kRedY; ++j) { ta += red_a[j][tx]; tb += red_b[j][tx]; }
      atomicAdd(&sums[ch], (double)ta);
      atomicAdd(&sums[c + ch], (double)tb);
    }
    __syncthreads();
  }
}

// Vectorised form (c % 4 == 0, aligned rows): a thread owns 4 consecutive channels (one 8/16-byte load per row) and a
// row lane; a block streams kVecRows rows.  F(row, ch, a[4], b[4]).
constexpr int kVecThreads = 256;
constexpr int kVecRows = 256;
template <typename F>
__device__ __forceinline__ void column_reduce2_vec(int64_t n, int c, int rows_per_block, double* __restrict__ sums, F f) {
  __shared__ float red[kVecThreads][9];
  const int groups = c >> 2;                        // <= 128
  const int lanes = kVecThreads / groups;           // row lanes per block (>= 2)
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < lanes) {
    const int64_t row_end = min((int64_t)(blockIdx.x + 1) * rows_per_block, n);
    for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rl; r < row_end; r += lanes) {
      float a[4], b[4];
      f(r, g * 4, a, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) { sa[j] += a[j]; sb[j] += b[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = sa[j]; red[threadIdx.x][4 + j] = sb[j]; }
  __syncthreads();
  if (threadIdx.x < groups) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += red[l * groups + threadIdx.x][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sums[threadIdx.x * 4 + j], (double)t[j]);
      atomicAdd(&sums[c + threadIdx.x * 4 + j], (double)t[4 + j]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_stats_vec_kernel(const T* __restrict__ x, int64_t ld, int64_t n, int c, int rows_per_block, double* __restrict__ stats) {
  column_reduce2_vec(n, c, rows_per_block, stats, [&](int64_t r, int ch, float (&a)[4], float (&b)[4]) {
    Vec4<T>::load(x + r * ld + ch, a);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = a[j] * a[j];
  });
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_bwd_reduce_vec_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                                         const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                                         const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                                                         int rows_per_block, double* __restrict__ sums) {
  column_reduce2_vec(n, c, rows_per_block, sums, [&](int64_t r, int ch, float (&a)[4], float (&b)[4]) {
    float xv[4], yv[4] = {1.f, 1.f, 1.f, 1.f};
    Vec4<T>::load(dy + r * ld_dy + ch, a);
    Vec4<T>::load(x + r * ld_x + ch, xv);
    if (relu) Vec4<T>::load(y + r * ld_y + ch, yv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (relu && !(yv[j] > 0.f)) a[j] = 0.f;
      b[j] = a[j] * (xv[j] - mean[ch + j]) * invstd[ch + j];
    }
  });
}

// Streaming elementwise passes: a block first derives the per-channel constants into shared memory (fp64 batch sums ->
// fp32 scale/shift; cheap once per block, ruinous once per thread), then streams kVecRows rows, 4 channels per thread.
template <typename T, bool kTrain>
__global__ void __launch_bounds__(kVecThreads) bn_apply_vec_kernel(const T* __restrict__ x, int64_t ld_x, int64_t n, int c,
                                                                    const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float eps, float momentum,
                                                                    float* running_mean, float* running_var, float* __restrict__ mean_out,
                                                                    float* __restrict__ invstd_out, const float* __restrict__ scale_in,
                                                                    const float* __restrict__ shift_in, const T* __restrict__ res, int64_t ld_res,
                                                                    int relu, T* __restrict__ y, int64_t ld_y) {
  __shared__ float s_scale[kMaxChannels], s_shift[kMaxChannels];
  for (int ch = threadIdx.x; ch < c; ch += kVecThreads) {
    if (kTrain) {
      const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
      const double m = stats[ch] * inv_n;
      double var = stats[c + ch] * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      const float is = rsqrtf((float)var + eps);
      const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
      s_scale[ch] = bn_scale_of(g, is);
      s_shift[ch] = bn_shift_of(b, (float)m, g, is);
      if (blockIdx.x == 0) {
        mean_out[ch] = (float)m;
        invstd_out[ch] = is;
        if (running_mean) {
          const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
          running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
          running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
        }
      }
    } else {
      s_scale[ch] = scale_in[ch];
      s_shift[ch] = shift_in[ch];
    }
  }
  __syncthreads();
  const int groups = c >> 2;
  const int step_r = kVecThreads / groups, step_g = kVecThreads % groups;
  int g = threadIdx.x % groups;
  int64_t r = (int64_t)blockIdx.x * kVecRows + threadIdx.x / groups;
  const int64_t row_end = min((int64_t)(blockIdx.x + 1) * kVecRows, n);
  while (r < row_end) {
    const int ch = g * 4;
    float v[4], rs[4] = {0.f, 0.f, 0.f, 0.f};
    Vec4<T>::load(x + r * ld_x + ch, v);
    if (res) Vec4<T>::load(res + r * ld_res + ch, rs);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float o = fmaf(v[j], s_scale[ch + j], s_shift[ch + j]) + rs[j];
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
    Vec4<T>::store(y + r * ld_y + ch, v);
    g += step_g; r += step_r;
    if (g >= groups) { g -= groups; ++r; }
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_bwd_apply_vec_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                                        const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                        const float* __restrict__ gamma, const double* __restrict__ sums, int relu,
                                                                        int training, T* __restrict__ dx, int64_t ld_dx, T* __restrict__ dres,
                                                                        int64_t ld_dres, float* dgamma, float* dbeta) {
  __shared__ float s_k[kMaxChannels], s_mean[kMaxChannels], s_is[kMaxChannels], s_sg[kMaxChannels], s_sgx[kMaxChannels];
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  for (int ch = threadIdx.x; ch < c; ch += kVecThreads) {
    const float is = invstd[ch];
    s_k[ch] = (gamma ? gamma[ch] : 1.f) * is;
    s_mean[ch] = mean[ch];
    s_is[ch] = is;
    s_sg[ch] = training ? (float)sums[ch] * inv_n : 0.f;
    s_sgx[ch] = training ? (float)sums[c + ch] * inv_n : 0.f;
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[ch] += (float)sums[ch];
      if (dgamma) dgamma[ch] += (float)sums[c + ch];
    }
  }
  __syncthreads();
  const int groups = c >> 2;
  const int step_r = kVecThreads / groups, step_g = kVecThreads % groups;
  int g = threadIdx.x % groups;
  int64_t r = (int64_t)blockIdx.x * kVecRows + threadIdx.x / groups;
  const int64_t row_end = min((int64_t)(blockIdx.x + 1) * kVecRows, n);
  while (r < row_end) {
    const int ch = g * 4;
    float gv[4], xv[4], yv[4] = {1.f, 1.f, 1.f, 1.f}, o[4];
    Vec4<T>::load(dy + r * ld_dy + ch, gv);
    Vec4<T>::load(x + r * ld_x + ch, xv);
    if (relu) Vec4<T>::load(y + r * ld_y + ch, yv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (relu && !(yv[j] > 0.f)) gv[j] = 0.f;
      const float xhat = (xv[j] - s_mean[ch + j]) * s_is[ch + j];
      o[j] = s_k[ch + j] * (gv[j] - s_sg[ch + j] - xhat * s_sgx[ch + j]);
    }
    Vec4<T>::store(dx + r * ld_dx + ch, o);
    if (dres) Vec4<T>::store(dres + r * ld_dres + ch, gv);
    g += step_g; r += step_r;
    if (g >= groups) { g -= groups; ++r; }
  }
}

// ---- 16-byte-wide streaming kernels (c % V == 0 with V = 4 fp32 / 8 bf16 channels per thread) -----------------------
// Sized by the host so that small tensors still fill the machine: an elementwise block takes `rows_per_block` rows with
// about two to eight 16-byte items per thread (two in flight at a time); deep MinkUNet levels have few rows and many
// channels, and blocks of 256 rows left most SMs idle there (16..70 blocks, 75 us for a 4 MB tensor).
template <typename T> struct VecW;
template <> struct VecW<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecW<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Walks the (row, channel group) items of a block two at a time: item i of thread t is t + i * kVecThreads.
struct ItemWalk {
  int g; int64_t r; int step_g; int step_r; int groups;
  __device__ __forceinline__ ItemWalk(int groups_, int64_t row0) : groups(groups_) {
    g = threadIdx.x % groups; r = row0 + threadIdx.x / groups;
    step_g = kVecThreads % groups; step_r = kVecThreads / groups;
  }
  __device__ __forceinline__ void next() { g += step_g; r += step_r; if (g >= groups) { g -= groups; ++r; } }
  // the same items backwards (two-phase kernels: the elementwise pass starts with the rows the reduction read last, which
  // are the ones still in the L2 when the layer's operands do not fit it): position on this thread's k-th item
  __device__ __forceinline__ void seek(int64_t row0, int64_t k) {
    const int64_t item = (int64_t)threadIdx.x + k * kVecThreads;
    r = row0 + item / groups; g = (int)(item % groups);
  }
  __device__ __forceinline__ void prev() { g -= step_g; r -= step_r; if (g < 0) { g += groups; --r; } }
};
// number of items thread t walks in a block of `rows` rows
__device__ __forceinline__ int64_t items_of_thread(int64_t rows, int groups) {
  const int64_t n_items = rows * groups;
  return n_items > (int64_t)threadIdx.x ? (n_items - 1 - threadIdx.x) / kVecThreads + 1 : 0;
}

template <typename T, int V, int U, typename F>
__device__ __forceinline__ void column_reduce2_wide(int64_t n, int c, int rows_per_block, double* __restrict__ sums, F f) {
  __shared__ float red[kVecThreads][2 * V + 1];
  const int groups = c / V;                         // <= 128
  const int lanes = kVecThreads / groups;           // row lanes per block (>= 2)
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  float sa[V], sb[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
  if (rl < lanes) {
    const int64_t row_end = min((int64_t)(blockIdx.x + 1) * rows_per_block, n);
    int64_t r = (int64_t)blockIdx.x * rows_per_block + rl;
    // U rows in flight per thread (HBM latency x bandwidth needs ~6 MB outstanding over the whole GPU; U = 4 for the
    // one-operand statistics pass, 2 for the three-operand backward reduction, which is register-bound)
    for (; r + (U - 1) * (int64_t)lanes < row_end; r += U * lanes) {
      float a[U][V], b[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) f(r + u * lanes, g * V, a[u], b[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) { sa[j] += a[u][j]; sb[j] += b[u][j]; }
    }
    for (; r < row_end; r += lanes) {
      float a0[V], b0[V];
      f(r, g * V, a0, b0);
#pragma unroll
      for (int j = 0; j < V; ++j) { sa[j] += a0[j]; sb[j] += b0[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { red[threadIdx.x][j] = sa[j]; red[threadIdx.x][V + j] = sb[j]; }
  __syncthreads();
  // one thread per (channel, quantity): 2c <= 1024 sums over the row lanes
  for (int q = threadIdx.x; q < 2 * c; q += kVecThreads) {
    const int ch = q < c ? q : q - c;
    const int grp = ch / V, j = (ch % V) + (q < c ? 0 : V);
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * groups + grp][j];
    atomicAdd(&sums[q], (double)t);
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_stats_wide_kernel(const T* __restrict__ x, int64_t ld, int64_t n, int c, int rows_per_block, double* __restrict__ stats) {
  pdl_trigger(); pdl_wait();
  constexpr int V = VecW<T>::V;
  column_reduce2_wide<T, V, 4>(n, c, rows_per_block, stats, [&](int64_t r, int ch, float (&a)[V], float (&b)[V]) {
    VecW<T>::load(x + r * ld + ch, a);
#pragma unroll
    for (int j = 0; j < V; ++j) b[j] = a[j] * a[j];
  });
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_bwd_reduce_wide_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                                          const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                                          const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                                                          int rows_per_block, double* __restrict__ sums) {
  pdl_trigger(); pdl_wait();
  constexpr int V = VecW<T>::V;
  __shared__ float s_mean[kMaxChannels], s_is[kMaxChannels];
  for (int ch = threadIdx.x; ch < c; ch += kVecThreads) { s_mean[ch] = mean[ch]; s_is[ch] = invstd[ch]; }
  __syncthreads();
  column_reduce2_wide<T, V, 2>(n, c, rows_per_block, sums, [&](int64_t r, int ch, float (&a)[V], float (&b)[V]) {
    float xv[V], yv[V];
    VecW<T>::load(dy + r * ld_dy + ch, a);
    VecW<T>::load(x + r * ld_x + ch, xv);
    if (relu) VecW<T>::load(y + r * ld_y + ch, yv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (relu && !(yv[j] > 0.f)) a[j] = 0.f;
      b[j] = a[j] * (xv[j] - s_mean[ch + j]) * s_is[ch + j];
    }
  });
}

// kCoherent: the statistics were accumulated by this very kernel (fused two-phase form below): read them from L2, not through
// the read-only path.
template <typename T, bool kTrain, bool kCoherent = false>
__device__ __forceinline__ void bn_apply_wide_body(const T* __restrict__ x, int64_t ld_x, int64_t n, int c,
                                                   const double* stats, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps, float momentum,
                                                   float* running_mean, float* running_var, float* __restrict__ mean_out,
                                                   float* __restrict__ invstd_out, const float* __restrict__ scale_in,
                                                   const float* __restrict__ shift_in, const T* __restrict__ res, int64_t ld_res,
                                                   int relu, T* __restrict__ y, int64_t ld_y, int rows_per_block) {
  constexpr int V = VecW<T>::V;
  __shared__ float s_scale[kMaxChannels], s_shift[kMaxChannels];
  for (int ch = threadIdx.x; ch < c; ch += kVecThreads) {
    if (kTrain) {
      const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
      const double m = (kCoherent ? __ldcg(&stats[ch]) : stats[ch]) * inv_n;
      double var = (kCoherent ? __ldcg(&stats[c + ch]) : stats[c + ch]) * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      const float is = rsqrtf((float)var + eps);
      const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
      s_scale[ch] = bn_scale_of(g, is);
      s_shift[ch] = bn_shift_of(b, (float)m, g, is);
      if (blockIdx.x == 0) {
        mean_out[ch] = (float)m;
        invstd_out[ch] = is;
        if (running_mean) {
          const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
          running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
          running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
        }
      }
    } else {
      s_scale[ch] = scale_in[ch];
      s_shift[ch] = shift_in[ch];
    }
  }
  __syncthreads();
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(row0 + rows_per_block, n);
  ItemWalk a(c / V, row0);
  // kCoherent (two-phase kernel): walk backwards, most recently read rows first
  int64_t left = kCoherent ? items_of_thread(row_end > row0 ? row_end - row0 : 0, c / V) : 0;
  if (kCoherent && left > 0) a.seek(row0, left - 1);
  while (kCoherent ? left > 0 : a.r < row_end) {
    ItemWalk b = a;
    if (kCoherent) b.prev(); else b.next();
    const bool two = kCoherent ? left > 1 : b.r < row_end;
    float va[V], vb[V], ra[V], rb[V];
    VecW<T>::load(x + a.r * ld_x + a.g * V, va);
    if (two) VecW<T>::load(x + b.r * ld_x + b.g * V, vb);
    if (res) {
      VecW<T>::load(res + a.r * ld_res + a.g * V, ra);
      if (two) VecW<T>::load(res + b.r * ld_res + b.g * V, rb);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float o = fmaf(va[j], s_scale[a.g * V + j], s_shift[a.g * V + j]);
      if (res) o += ra[j];
      va[j] = relu ? fmaxf(o, 0.f) : o;
    }
    VecW<T>::store(y + a.r * ld_y + a.g * V, va);
    if (two) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float o = fmaf(vb[j], s_scale[b.g * V + j], s_shift[b.g * V + j]);
        if (res) o += rb[j];
        vb[j] = relu ? fmaxf(o, 0.f) : o;
      }
      VecW<T>::store(y + b.r * ld_y + b.g * V, vb);
    }
    a = b;
    if (kCoherent) { a.prev(); left -= 2; } else a.next();
  }
}

template <typename T, bool kTrain>
__global__ void __launch_bounds__(kVecThreads) bn_apply_wide_kernel(const T* __restrict__ x, int64_t ld_x, int64_t n, int c,
                                                                     const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps, float momentum,
                                                                     float* running_mean, float* running_var, float* __restrict__ mean_out,
                                                                     float* __restrict__ invstd_out, const float* __restrict__ scale_in,
                                                                     const float* __restrict__ shift_in, const T* __restrict__ res, int64_t ld_res,
                                                                     int relu, T* __restrict__ y, int64_t ld_y, int rows_per_block) {
  pdl_trigger(); pdl_wait();
  bn_apply_wide_body<T, kTrain>(x, ld_x, n, c, stats, gamma, beta, eps, momentum, running_mean, running_var, mean_out, invstd_out, scale_in, shift_in,
                                res, ld_res, relu, y, ld_y, rows_per_block);
}

// kMaskFromX (two-phase kernel, ReLU without residual): the mask is (x * scale + shift > 0) re-derived from x (beta_mask != NULL)
// and y is not read at all; kMaxC bounds the channel count (size of the per-channel arrays).
template <typename T, bool kCoherent = false, int kMaxC = kMaxChannels, bool kMaskFromX = false>
__device__ __forceinline__ void bn_bwd_apply_wide_body(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                       const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       const float* __restrict__ gamma, const double* sums, int relu,
                                                       int training, T* __restrict__ dx, int64_t ld_dx, T* __restrict__ dres,
                                                       int64_t ld_dres, float* dgamma, float* dbeta, int rows_per_block,
                                                       const float* __restrict__ beta_mask = nullptr) {
  constexpr int V = VecW<T>::V;
  __shared__ float s_k[kMaxC], s_mean[kMaxC], s_is[kMaxC], s_sg[kMaxC], s_sgx[kMaxC];
  __shared__ float s_shift[kMaskFromX ? kMaxC : 1];
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  for (int ch = threadIdx.x; ch < c; ch += kVecThreads) {
    const float is = invstd[ch];
    const double sum_g = kCoherent ? __ldcg(&sums[ch]) : sums[ch], sum_gx = kCoherent ? __ldcg(&sums[c + ch]) : sums[c + ch];
    s_k[ch] = bn_scale_of(gamma ? gamma[ch] : 1.f, is);
    if (kMaskFromX) s_shift[ch] = bn_shift_of(beta_mask[ch], mean[ch], gamma ? gamma[ch] : 1.f, is);
    s_mean[ch] = mean[ch];
    s_is[ch] = is;
    s_sg[ch] = training ? (float)sum_g * inv_n : 0.f;
    s_sgx[ch] = training ? (float)sum_gx * inv_n : 0.f;
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[ch] += (float)sum_g;
      if (dgamma) dgamma[ch] += (float)sum_gx;
    }
  }
  __syncthreads();
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(row0 + rows_per_block, n);
  ItemWalk a(c / V, row0);
  int64_t left = kCoherent ? items_of_thread(row_end > row0 ? row_end - row0 : 0, c / V) : 0;      // two-phase kernel: backwards (see ItemWalk::seek)
  if (kCoherent && left > 0) a.seek(row0, left - 1);
  while (kCoherent ? left > 0 : a.r < row_end) {
    ItemWalk b = a;
    if (kCoherent) b.prev(); else b.next();
    const bool two = kCoherent ? left > 1 : b.r < row_end;
    float ga[V], gb[V], xa[V], xb[V], ya[V], yb[V], o[V];
    VecW<T>::load(dy + a.r * ld_dy + a.g * V, ga);
    VecW<T>::load(x + a.r * ld_x + a.g * V, xa);
    if (relu && !kMaskFromX) VecW<T>::load(y + a.r * ld_y + a.g * V, ya);
    if (two) {
      VecW<T>::load(dy + b.r * ld_dy + b.g * V, gb);
      VecW<T>::load(x + b.r * ld_x + b.g * V, xb);
      if (relu && !kMaskFromX) VecW<T>::load(y + b.r * ld_y + b.g * V, yb);
    }
    if (kMaskFromX) {           // what the forward pass stored: relu(x * scale + shift)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        ya[j] = fmaf(xa[j], s_k[a.g * V + j], s_shift[a.g * V + j]);
        if (two) yb[j] = fmaf(xb[j], s_k[b.g * V + j], s_shift[b.g * V + j]);
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int ch = a.g * V + j;
      if (relu && !(ya[j] > 0.f)) ga[j] = 0.f;
      o[j] = bn_dx_of(ga[j], xa[j], s_k[ch], s_mean[ch], s_is[ch], s_sg[ch], s_sgx[ch]);
    }
    VecW<T>::store(dx + a.r * ld_dx + a.g * V, o);
    if (dres) VecW<T>::store(dres + a.r * ld_dres + a.g * V, ga);
    if (two) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int ch = b.g * V + j;
        if (relu && !(yb[j] > 0.f)) gb[j] = 0.f;
        o[j] = bn_dx_of(gb[j], xb[j], s_k[ch], s_mean[ch], s_is[ch], s_sg[ch], s_sgx[ch]);
      }
      VecW<T>::store(dx + b.r * ld_dx + b.g * V, o);
      if (dres) VecW<T>::store(dres + b.r * ld_dres + b.g * V, gb);
    }
    a = b;
    if (kCoherent) { a.prev(); left -= 2; } else a.next();
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_bwd_apply_wide_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                                         const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                         const float* __restrict__ gamma, const double* __restrict__ sums, int relu,
                                                                         int training, T* __restrict__ dx, int64_t ld_dx, T* __restrict__ dres,
                                                                         int64_t ld_dres, float* dgamma, float* dbeta, int rows_per_block) {
  pdl_trigger(); pdl_wait();
  bn_bwd_apply_wide_body<T>(dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, dx, ld_dx, dres, ld_dres, dgamma, dbeta,
                            rows_per_block);
}

// ---- two-phase forms (cooperative launch): the column reduction, a grid-wide barrier, then the elementwise pass over the
// SAME rows of the same block, which the block has just pulled through the L2.  One launch instead of two per batch norm and
// direction (124 fewer launches and dependent-launch gaps per MinkUNet34 step), and the second read of the operands is an L2
// hit instead of a pass over HBM whenever the layer's tensors fit the 126 MB L2.
struct BnFwdFusedArgs {
  const void* x; int64_t ld_x; int64_t n; int c;
  double* stats; const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var; float* mean; float* invstd;
  const void* res; int64_t ld_res; int relu; void* y; int64_t ld_y; int rows_per_block;
};
template <typename T>
__global__ void __launch_bounds__(kVecThreads) bn_fwd_fused_wide_kernel(const BnFwdFusedArgs a) {
  pdl_trigger(); pdl_wait();
  constexpr int V = VecW<T>::V;
  const T* x = static_cast<const T*>(a.x);
  column_reduce2_wide<T, V, 4>(a.n, a.c, a.rows_per_block, a.stats, [&](int64_t r, int ch, float (&va)[V], float (&vb)[V]) {
    VecW<T>::load(x + r * a.ld_x + ch, va);
#pragma unroll
    for (int j = 0; j < V; ++j) vb[j] = va[j] * va[j];
  });
  __threadfence();
  cooperative_groups::this_grid().sync();
  bn_apply_wide_body<T, true, true>(x, a.ld_x, a.n, a.c, a.stats, a.gamma, a.beta, a.eps, a.momentum, a.running_mean, a.running_var, a.mean, a.invstd,
                                    nullptr, nullptr, static_cast<const T*>(a.res), a.ld_res, a.relu, static_cast<T*>(a.y), a.ld_y, a.rows_per_block);
}

struct BnBwdFusedArgs {
  const void* dy; int64_t ld_dy; const void* x; int64_t ld_x; const void* y; int64_t ld_y; int64_t n; int c;
  const float* mean; const float* invstd; const float* gamma; double* sums; int relu;
  void* dx; int64_t ld_dx; void* dres; int64_t ld_dres; float* dgamma; float* dbeta; int rows_per_block;
  const float* beta;      // kMaskFromX kernels: the forward pass's beta (mask = x * scale + shift > 0)
};
// kMaskFromX: conv -> BN -> ReLU units (no residual): both phases read dy and x only; the activation (one of three operand
// streams of the reduction, one of three reads of the elementwise phase) stays in HBM.
template <typename T, bool kMaskFromX>
__global__ void __launch_bounds__(kVecThreads, 3) bn_bwd_fused_wide_kernel(const BnBwdFusedArgs a) {
  pdl_trigger(); pdl_wait();
  constexpr int V = VecW<T>::V;
  const T* dy = static_cast<const T*>(a.dy);
  const T* x = static_cast<const T*>(a.x);
  const T* y = static_cast<const T*>(a.y);
  {
    __shared__ float s_mean[kFusedBwdMaxChannels], s_is[kFusedBwdMaxChannels];
    for (int ch = threadIdx.x; ch < a.c; ch += kVecThreads) { s_mean[ch] = a.mean[ch]; s_is[ch] = a.invstd[ch]; }
    __syncthreads();
    // a thread of the reduction owns one group of V channels for all its rows: its folded scale / shift live in registers
    float sc[V], sh[V];
    if (kMaskFromX) {
      const int ch0 = (int)(threadIdx.x % (a.c / V)) * V;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float g = a.gamma ? a.gamma[ch0 + j] : 1.f;
        sc[j] = bn_scale_of(g, s_is[ch0 + j]);
        sh[j] = bn_shift_of(a.beta[ch0 + j], s_mean[ch0 + j], g, s_is[ch0 + j]);
      }
    }
    column_reduce2_wide<T, V, 2>(a.n, a.c, a.rows_per_block, a.sums, [&](int64_t r, int ch, float (&va)[V], float (&vb)[V]) {
      float xv[V], yv[V];
      VecW<T>::load(dy + r * a.ld_dy + ch, va);
      VecW<T>::load(x + r * a.ld_x + ch, xv);
      if (a.relu && !kMaskFromX) VecW<T>::load(y + r * a.ld_y + ch, yv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (kMaskFromX) yv[j] = fmaf(xv[j], sc[j], sh[j]);
        if (a.relu && !(yv[j] > 0.f)) va[j] = 0.f;
        vb[j] = va[j] * (xv[j] - s_mean[ch + j]) * s_is[ch + j];
      }
    });
  }
  __threadfence();
  cooperative_groups::this_grid().sync();
  bn_bwd_apply_wide_body<T, true, kFusedBwdMaxChannels, kMaskFromX>(dy, a.ld_dy, x, a.ld_x, y, a.ld_y, a.n, a.c, a.mean, a.invstd, a.gamma, a.sums,
                                                                    a.relu, 1, static_cast<T*>(a.dx), a.ld_dx, static_cast<T*>(a.dres), a.ld_dres,
                                                                    a.dgamma, a.dbeta, a.rows_per_block, a.beta);
}

template <typename T>
__global__ void __launch_bounds__(kRedX* kRedY) bn_stats_kernel(const T* __restrict__ x, int64_t ld, int64_t n, int c, double* __restrict__ stats) {
  column_reduce2(n, c, stats, [&](int64_t r, int ch, float& a, float& b) {
    const float v = to_f32<T>(x[r * ld + ch]);
    a = v; b = v * v;
  });
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int64_t n, int c, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
  const double m = stats[ch] * inv_n;
  double var = stats[c + ch] * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[ch] = (float)m;
  invstd[ch] = is;
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  scale[ch] = g * is;
  shift[ch] = b - (float)m * g * is;
  if (running_mean) {
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

__global__ void bn_fold_eval_kernel(int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float is = 1.f / sqrtf(rv[ch] + eps);
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  scale[ch] = g * is;
  shift[ch] = b - rm[ch] * g * is;
}

// y = act(x*scale + shift + residual); one thread per 4 consecutive channels (c % 4 == 0 fast path).
template <typename T, bool kVec>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, int64_t ld_x, int64_t n, int c,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        const T* __restrict__ res, int64_t ld_res, int relu, T* __restrict__ y, int64_t ld_y) {
  const int groups = (c + 3) / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * groups) return;
  const int64_t r = t / groups;
  const int ch = (int)(t - r * groups) * 4;
  float v[4], rs[4] = {0.f, 0.f, 0.f, 0.f};
  if (kVec) {
    Vec4<T>::load(x + r * ld_x + ch, v);
    if (res) Vec4<T>::load(res + r * ld_res + ch, rs);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = ch + j < c ? to_f32<T>(x[r * ld_x + ch + j]) : 0.f;
      if (res && ch + j < c) rs[j] = to_f32<T>(res[r * ld_res + ch + j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (ch + j < c) {
      float o = fmaf(v[j], scale[ch + j], shift[ch + j]) + rs[j];
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
  }
  if (kVec) Vec4<T>::store(y + r * ld_y + ch, v);
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (ch + j < c) y[r * ld_y + ch + j] = from_f32<T>(v[j]);
  }
}

// Training-mode normalise in one pass: every thread derives scale/shift for its four channels from the fp64 batch
// sums (cheap next to the memory traffic), block 0 also publishes mean / invstd for the backward pass and updates the
// running statistics, so no separate "finalise" launch is needed.
template <typename T, bool kVec>
__global__ void __launch_bounds__(256) bn_apply_train_kernel(const T* __restrict__ x, int64_t ld_x, int64_t n, int c,
                                                              const double* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, float momentum,
                                                              float* running_mean, float* running_var, float* __restrict__ mean_out,
                                                              float* __restrict__ invstd_out, const T* __restrict__ res, int64_t ld_res,
                                                              int relu, T* __restrict__ y, int64_t ld_y) {
  const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
  if (blockIdx.x == 0) {
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const double m = stats[ch] * inv_n;
      double var = stats[c + ch] * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      mean_out[ch] = (float)m;
      invstd_out[ch] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean) {
        const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
      }
    }
  }
  const int groups = (c + 3) / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * groups) return;
  const int64_t r = t / groups;
  const int ch = (int)(t - r * groups) * 4;
  float v[4], rs[4] = {0.f, 0.f, 0.f, 0.f};
  if (kVec) {
    Vec4<T>::load(x + r * ld_x + ch, v);
    if (res) Vec4<T>::load(res + r * ld_res + ch, rs);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = ch + j < c ? to_f32<T>(x[r * ld_x + ch + j]) : 0.f;
      if (res && ch + j < c) rs[j] = to_f32<T>(res[r * ld_res + ch + j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (ch + j < c) {
      const double m = stats[ch + j] * inv_n;
      double var = stats[c + ch + j] * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      const float is = (float)(1.0 / sqrt(var + (double)eps));
      const float g = gamma ? gamma[ch + j] : 1.f, b = beta ? beta[ch + j] : 0.f;
      const float scale = g * is, shift = b - (float)m * scale;
      const float o = fmaf(v[j], scale, shift) + rs[j];
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
  }
  if (kVec) Vec4<T>::store(y + r * ld_y + ch, v);
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (ch + j < c) y[r * ld_y + ch + j] = from_f32<T>(v[j]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kRedX* kRedY) bn_bwd_reduce_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                                     const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                                                     double* __restrict__ sums) {
  column_reduce2(n, c, sums, [&](int64_t r, int ch, float& a, float& b) {
    float g = to_f32<T>(dy[r * ld_dy + ch]);
    if (relu && !(to_f32<T>(y[r * ld_y + ch]) > 0.f)) g = 0.f;
    const float xhat = (to_f32<T>(x[r * ld_x + ch]) - mean[ch]) * invstd[ch];
    a = g; b = g * xhat;
  });
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x, int64_t ld_x,
                                                            const T* __restrict__ y, int64_t ld_y, int64_t n, int c,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const double* __restrict__ sums, int relu,
                                                            int training, T* __restrict__ dx, int64_t ld_dx, T* __restrict__ dres, int64_t ld_dres,
                                                            float* dgamma, float* dbeta) {
  if (blockIdx.x == 0) {                              // parameter gradients: dbeta = sum g, dgamma = sum g * xhat (accumulated)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      if (dbeta) dbeta[ch] += (float)sums[ch];
      if (dgamma) dgamma[ch] += (float)sums[c + ch];
    }
  }
  const int groups = (c + 3) / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * groups) return;
  const int64_t r = t / groups;
  const int ch = (int)(t - r * groups) * 4;
  float g[4], xv[4], yv[4] = {1.f, 1.f, 1.f, 1.f};
  if (kVec) {
    Vec4<T>::load(dy + r * ld_dy + ch, g);
    Vec4<T>::load(x + r * ld_x + ch, xv);
    if (relu) Vec4<T>::load(y + r * ld_y + ch, yv);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = ch + j < c;
      g[j] = ok ? to_f32<T>(dy[r * ld_dy + ch + j]) : 0.f;
      xv[j] = ok ? to_f32<T>(x[r * ld_x + ch + j]) : 0.f;
      if (relu && ok) yv[j] = to_f32<T>(y[r * ld_y + ch + j]);
    }
  }
  const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[j] = 0.f;
    if (ch + j < c) {
      if (relu && !(yv[j] > 0.f)) g[j] = 0.f;
      const float is = invstd[ch + j], gm = gamma ? gamma[ch + j] : 1.f;
      if (training) {
        const float xhat = (xv[j] - mean[ch + j]) * is;
        const float sg = (float)sums[ch + j] * inv_n, sgx = (float)sums[c + ch + j] * inv_n;
        o[j] = gm * is * (g[j] - sg - xhat * sgx);
      } else {
        o[j] = gm * is * g[j];
      }
    }
  }
  if (kVec) {
    Vec4<T>::store(dx + r * ld_dx + ch, o);
    if (dres) Vec4<T>::store(dres + r * ld_dres + ch, g);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (ch + j < c) {
      dx[r * ld_dx + ch + j] = from_f32<T>(o[j]);
      if (dres) dres[r * ld_dres + ch + j] = from_f32<T>(g[j]);
    }
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, int c, float* dgamma, float* dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  if (dbeta) dbeta[ch] += (float)sums[ch];
  if (dgamma) dgamma[ch] += (float)sums[c + ch];
}

template <typename T>
__global__ void __launch_bounds__(256) relu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t numel) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < numel) y[t] = from_f32<T>(fmaxf(to_f32<T>(x[t]), 0.f));
}
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t numel) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < numel) dx[t] = to_f32<T>(y[t]) > 0.f ? dy[t] : from_f32<T>(0.f);
}

inline unsigned reduce_grid(int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kRedY * 8), (int64_t)kNumSMs * 8)); }
template <typename T> bool vec_ok(int c, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  const int64_t a = sizeof(T) == 4 ? 16 : 8;
  if (c % 4) return false;
  for (int64_t ld : lds) if (ld % 4) return false;
  for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) % a)) return false;
  return true;
}
}  // namespace
}  // namespace gcd

using namespace gcd;

namespace {
inline unsigned vec_grid(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div(n, kVecRows)); }
// reductions end in one fp64 atomic per channel per block: few, fat blocks (about 4 per SM) keep the atomics rare
inline int red_rows(int64_t n) { return (int)std::max<int64_t>(kVecRows, ceil_div(n, (int64_t)kNumSMs * 4)); }
inline unsigned red_grid(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div(n, red_rows(n))); }
inline bool vec_shape_ok(int c) { return c % 4 == 0 && c >= 8 && c <= kMaxChannels; }
// 16-byte path: V channels per thread, every leading dimension a multiple of V, every pointer 16-byte aligned
template <typename T> bool wide_ok(int c, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  const int V = 16 / (int)sizeof(T);
  if (c % V || c < 2 * V || c > kMaxChannels) return false;
  for (int64_t ld : lds) if (ld % V) return false;
  for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) % 16)) return false;
  return true;
}
// elementwise passes: ~8 blocks per SM when the tensor allows, never less than two items per thread
template <typename T> int wide_rows(int64_t n, int c) {
  const int groups = c / (16 / (int)sizeof(T));
  const int64_t min_rows = ceil_div(2 * kVecThreads, groups);
  return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, (int64_t)kNumSMs * 8), min_rows), 4096);
}
// reductions: ~4 blocks per SM (one fp64 atomic per channel per block), at least four rows per row lane
template <typename T> int wide_red_rows(int64_t n, int c) {
  const int groups = c / (16 / (int)sizeof(T));
  const int64_t min_rows = (int64_t)(kVecThreads / groups) * 4;
  return (int)std::max<int64_t>(ceil_div(n, (int64_t)kNumSMs * 4), min_rows);
}
inline unsigned rows_grid(int64_t n, int rows) { return (unsigned)std::max<int64_t>(1, ceil_div(n, rows)); }
}  // namespace

namespace gcd {
namespace {
// co-resident blocks per SM of the two cooperative kernels (queried once)
template <typename Kernel> int coop_blocks_per_sm(Kernel k) {
  int b = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k, kVecThreads, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  return b;
}
template <typename T> int fwd_fused_occupancy() { static const int v = coop_blocks_per_sm(bn_fwd_fused_wide_kernel<T>); return v; }
template <typename T> int bwd_fused_occupancy() {
  static const int v = std::min(coop_blocks_per_sm(bn_bwd_fused_wide_kernel<T, false>), coop_blocks_per_sm(bn_bwd_fused_wide_kernel<T, true>));
  return v;
}
// rows per block of a cooperative launch: the partition of the stand-alone reduction when it fits on the GPU at once, fatter
// blocks otherwise
template <typename T> int coop_rows(int64_t n, int c, int occupancy) {
  int rows = wide_red_rows<T>(n, c);
  const int64_t max_grid = (int64_t)occupancy * kNumSMs;
  if (ceil_div(n, rows) > max_grid) rows = (int)ceil_div(n, max_grid);
  return rows;
}
}  // namespace
template <typename T> int bwd_partition_rows(int64_t n, int c) { return coop_rows<T>(n, c, std::max(bwd_fused_occupancy<T>(), 1)); }
template <typename T> int fwd_partition_rows(int64_t n, int c) { return coop_rows<T>(n, c, std::max(fwd_fused_occupancy<T>(), 1)); }

// batch statistics + normalise(+ReLU)(+residual) of a training-mode batch norm as ONE launch when the 16-byte path applies and
// the option allows it, else the two stand-alone passes.  stats [2c] zero on entry.
int32_t bn_forward_train(const void* x, int64_t ld_x, int64_t n, int32_t c, double* stats, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, float* mean, float* invstd, const void* res,
                         int64_t ld_res, int32_t relu, void* y, int64_t ld_y, int32_t dtype, void* stream) {
  if (n > 0 && option(GCD_OPT_BN_FUSED)) {
    BnFwdFusedArgs a{x, ld_x, n, c, stats, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, res, ld_res, relu, y, ld_y, 0};
    cudaError_t e = cudaErrorInvalidValue;
    if (dtype == GCD_BF16 && wide_ok<__nv_bfloat16>(c, {ld_x, ld_y, res ? ld_res : 0}, {x, y, res}) && fwd_fused_occupancy<__nv_bfloat16>() > 0) {
      a.rows_per_block = fwd_partition_rows<__nv_bfloat16>(n, c);
      e = launch_coop_pdl(bn_fwd_fused_wide_kernel<__nv_bfloat16>, dim3(rows_grid(n, a.rows_per_block)), dim3(kVecThreads), as_stream(stream), a);
    } else if (dtype == GCD_F32 && wide_ok<float>(c, {ld_x, ld_y, res ? ld_res : 0}, {x, y, res}) && fwd_fused_occupancy<float>() > 0) {
      a.rows_per_block = fwd_partition_rows<float>(n, c);
      e = launch_coop_pdl(bn_fwd_fused_wide_kernel<float>, dim3(rows_grid(n, a.rows_per_block)), dim3(kVecThreads), as_stream(stream), a);
    }
    if (e == cudaSuccess) return GCD_OK;
    cudaGetLastError();          // not applicable / not launchable here: the two-pass form below
  }
  int32_t rc = gcd_bn_stats(x, ld_x, n, c, dtype, stats, stream);
  if (rc != GCD_OK) return rc;
  return gcd_bn_apply_train(x, ld_x, n, c, stats, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, res, ld_res, relu, y, ld_y,
                            dtype, stream);
}

// the same for the backward pass (training mode): sums [2c] zero on entry
// beta_mask (optional): the unit is conv -> BN -> ReLU without a residual and its forward pass ran through bn_forward_train of
// this library: the two-phase kernel re-derives the ReLU mask from x and does not read y (GCD_OPT_BN_MASK_FROM_X).
int32_t bn_backward_train(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y, int64_t n, int32_t c,
                          const float* mean, const float* invstd, const float* gamma, double* sums, int32_t relu, void* dx, int64_t ld_dx,
                          void* dres, int64_t ld_dres, float* dgamma, float* dbeta, int32_t dtype, void* stream, const float* beta_mask) {
  if (n > 0 && option(GCD_OPT_BN_FUSED) && c <= kFusedBwdMaxChannels) {
    const bool from_x = relu && beta_mask != nullptr && dres == nullptr && option(GCD_OPT_BN_MASK_FROM_X);
    BnBwdFusedArgs a{dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, gamma, sums, relu, dx, ld_dx, dres, ld_dres, dgamma, dbeta, 0, beta_mask};
    cudaError_t e = cudaErrorInvalidValue;
    if (dtype == GCD_BF16 && wide_ok<__nv_bfloat16>(c, {ld_dy, ld_x, relu ? ld_y : 0, ld_dx, dres ? ld_dres : 0}, {dy, x, relu ? y : nullptr, dx, dres}) &&
        bwd_fused_occupancy<__nv_bfloat16>() > 0) {
      a.rows_per_block = bwd_partition_rows<__nv_bfloat16>(n, c);
      e = launch_coop_pdl(from_x ? bn_bwd_fused_wide_kernel<__nv_bfloat16, true> : bn_bwd_fused_wide_kernel<__nv_bfloat16, false>,
                          dim3(rows_grid(n, a.rows_per_block)), dim3(kVecThreads), as_stream(stream), a);
    } else if (dtype == GCD_F32 && wide_ok<float>(c, {ld_dy, ld_x, relu ? ld_y : 0, ld_dx, dres ? ld_dres : 0}, {dy, x, relu ? y : nullptr, dx, dres}) &&
               bwd_fused_occupancy<float>() > 0) {
      a.rows_per_block = bwd_partition_rows<float>(n, c);
      e = launch_coop_pdl(from_x ? bn_bwd_fused_wide_kernel<float, true> : bn_bwd_fused_wide_kernel<float, false>,
                          dim3(rows_grid(n, a.rows_per_block)), dim3(kVecThreads), as_stream(stream), a);
    }
    if (e == cudaSuccess) return GCD_OK;
    cudaGetLastError();
  }
  int32_t rc = gcd_bn_backward_reduce(dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, relu, dtype, sums, stream);
  if (rc != GCD_OK) return rc;
  return gcd_bn_backward_apply(dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, gamma, sums, relu, 1, dx, ld_dx, dres, ld_dres, dgamma, dbeta, dtype,
                               stream);
}
}  // namespace gcd

extern "C" int32_t gcd_bn_stats(const void* x, int64_t ld, int64_t n, int32_t c, int32_t dtype, double* stats, void* stream) {
  GCD_REQUIRE(c >= 1 && c <= kRedX * kMaxChanIter, "gcd_bn_stats: channel count %d out of range", c);
  if (n == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  dim3 block(kRedX, kRedY);
  if (dtype == GCD_F32) {
    using T = float;
    if (wide_ok<T>(c, {ld}, {x})) { const int rows = gcd::fwd_partition_rows<T>(n, c); launch_pdl(bn_stats_wide_kernel<T>, dim3(rows_grid(n, rows)), dim3(kVecThreads), 0, st, (const T*)x, ld, n, c, rows, stats); }
    else if (vec_shape_ok(c) && vec_ok<T>(c, {ld}, {x})) bn_stats_vec_kernel<T><<<red_grid(n), kVecThreads, 0, st>>>((const T*)x, ld, n, c, red_rows(n), stats);
    else bn_stats_kernel<T><<<reduce_grid(n), block, 0, st>>>((const T*)x, ld, n, c, stats);
  } else {
    using T = __nv_bfloat16;
    if (wide_ok<T>(c, {ld}, {x})) { const int rows = gcd::fwd_partition_rows<T>(n, c); launch_pdl(bn_stats_wide_kernel<T>, dim3(rows_grid(n, rows)), dim3(kVecThreads), 0, st, (const T*)x, ld, n, c, rows, stats); }
    else if (vec_shape_ok(c) && vec_ok<T>(c, {ld}, {x})) bn_stats_vec_kernel<T><<<red_grid(n), kVecThreads, 0, st>>>((const T*)x, ld, n, c, red_rows(n), stats);
    else bn_stats_kernel<T><<<reduce_grid(n), block, 0, st>>>((const T*)x, ld, n, c, stats);
  }
  GCD_LAUNCH_CHECK("gcd_bn_stats");
  return GCD_OK;
}

extern "C" int32_t gcd_bn_finalize(const double* stats, int64_t n, int32_t c, const float* gamma, const float* beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                                   float* scale, float* shift, void* stream) {
  bn_finalize_kernel<<<(unsigned)ceil_div(c, 128), 128, 0, as_stream(stream)>>>(stats, n, c, gamma, beta, eps, momentum, running_mean,
                                                                                running_var, mean, invstd, scale, shift);
  GCD_LAUNCH_CHECK("gcd_bn_finalize");
  return GCD_OK;
}

extern "C" int32_t gcd_bn_fold_eval(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                                    const float* running_var, float eps, float* scale, float* shift, void* stream) {
  bn_fold_eval_kernel<<<(unsigned)ceil_div(c, 128), 128, 0, as_stream(stream)>>>(c, gamma, beta, running_mean, running_var, eps, scale, shift);
  GCD_LAUNCH_CHECK("gcd_bn_fold_eval");
  return GCD_OK;
}

namespace {
template <typename T>
void launch_apply(bool train, const void* x, int64_t ld_x, int64_t n, int c, const double* stats, const float* gamma, const float* beta,
                  float eps, float momentum, float* rm, float* rv, float* mean, float* invstd, const float* scale, const float* shift,
                  const void* res, int64_t ld_res, int relu, void* y, int64_t ld_y, cudaStream_t st) {
  if (wide_ok<T>(c, {ld_x, ld_y, res ? ld_res : 0}, {x, y, res})) {
    const int rows = wide_rows<T>(n, c);
    const unsigned g = rows_grid(n, rows);
    if (train) launch_pdl(bn_apply_wide_kernel<T, true>, dim3(g), dim3(kVecThreads), 0, st, (const T*)x, ld_x, n, c, stats, gamma, beta, eps, momentum, rm, rv, mean, invstd, nullptr, nullptr, (const T*)res, ld_res, relu, (T*)y, ld_y, rows);
    else launch_pdl(bn_apply_wide_kernel<T, false>, dim3(g), dim3(kVecThreads), 0, st, (const T*)x, ld_x, n, c, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, scale, shift, (const T*)res, ld_res, relu, (T*)y, ld_y, rows);
    return;
  }
  const bool vec = vec_ok<T>(c, {ld_x, ld_y, res ? ld_res : 0}, {x, y, res});
  if (vec && vec_shape_ok(c)) {
    if (train) bn_apply_vec_kernel<T, true><<<vec_grid(n), kVecThreads, 0, st>>>((const T*)x, ld_x, n, c, stats, gamma, beta, eps, momentum, rm, rv, mean, invstd, nullptr, nullptr, (const T*)res, ld_res, relu, (T*)y, ld_y);
    else bn_apply_vec_kernel<T, false><<<vec_grid(n), kVecThreads, 0, st>>>((const T*)x, ld_x, n, c, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, scale, shift, (const T*)res, ld_res, relu, (T*)y, ld_y);
    return;
  }
  const unsigned g = (unsigned)std::max<int64_t>(1, ceil_div(n * ((c + 3) / 4), 256));
  if (train) {
    if (vec) bn_apply_train_kernel<T, true><<<g, 256, 0, st>>>((const T*)x, ld_x, n, c, stats, gamma, beta, eps, momentum, rm, rv, mean, invstd, (const T*)res, ld_res, relu, (T*)y, ld_y);
    else bn_apply_train_kernel<T, false><<<g, 256, 0, st>>>((const T*)x, ld_x, n, c, stats, gamma, beta, eps, momentum, rm, rv, mean, invstd, (const T*)res, ld_res, relu, (T*)y, ld_y);
  } else {
    if (vec) bn_apply_kernel<T, true><<<g, 256, 0, st>>>((const T*)x, ld_x, n, c, scale, shift, (const T*)res, ld_res, relu, (T*)y, ld_y);
    else bn_apply_kernel<T, false><<<g, 256, 0, st>>>((const T*)x, ld_x, n, c, scale, shift, (const T*)res, ld_res, relu, (T*)y, ld_y);
  }
}
}  // namespace

extern "C" int32_t gcd_bn_apply(const void* x, int64_t ld_x, int64_t n, int32_t c, const float* scale, const float* shift,
                                const void* residual, int64_t ld_res, int32_t relu, void* y, int64_t ld_y, int32_t dtype, void* stream) {
  if (n == 0) return GCD_OK;
  GCD_REQUIRE(c >= 1 && c <= kMaxChannels, "gcd_bn_apply: channel count %d out of range", c);
  cudaStream_t st = as_stream(stream);
  if (dtype == GCD_F32) launch_apply<float>(false, x, ld_x, n, c, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, scale, shift, residual, ld_res, relu, y, ld_y, st);
  else launch_apply<__nv_bfloat16>(false, x, ld_x, n, c, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, scale, shift, residual, ld_res, relu, y, ld_y, st);
  GCD_LAUNCH_CHECK("gcd_bn_apply");
  return GCD_OK;
}

extern "C" int32_t gcd_bn_apply_train(const void* x, int64_t ld_x, int64_t n, int32_t c, const double* stats, const float* gamma,
                                      const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                      float* mean, float* invstd, const void* residual, int64_t ld_res, int32_t relu, void* y,
                                      int64_t ld_y, int32_t dtype, void* stream) {
  GCD_REQUIRE(stats && mean && invstd && c >= 1 && c <= kMaxChannels, "gcd_bn_apply_train: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (dtype == GCD_F32) launch_apply<float>(true, x, ld_x, n, c, stats, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, nullptr, nullptr, residual, ld_res, relu, y, ld_y, st);
  else launch_apply<__nv_bfloat16>(true, x, ld_x, n, c, stats, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, nullptr, nullptr, residual, ld_res, relu, y, ld_y, st);
  GCD_LAUNCH_CHECK("gcd_bn_apply_train");
  return GCD_OK;
}

extern "C" int32_t gcd_bn_backward_reduce(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                                          int64_t n, int32_t c, const float* mean, const float* invstd, int32_t relu, int32_t dtype,
                                          double* sums, void* stream) {
  GCD_REQUIRE(c >= 1 && c <= kRedX * kMaxChanIter, "gcd_bn_backward_reduce: channel count %d out of range", c);
  if (n == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  dim3 block(kRedX, kRedY);
  if (dtype == GCD_F32) {
    using T = float;
    if (wide_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0}, {dy, x, relu ? y : nullptr})) {
      const int rows = gcd::bwd_partition_rows<T>(n, c);      // the partition of the fused two-phase form: both paths sum the same partials
      launch_pdl(bn_bwd_reduce_wide_kernel<T>, dim3(rows_grid(n, rows)), dim3(kVecThreads), 0, st, (const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, rows, sums);
    } else if (vec_shape_ok(c) && vec_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0}, {dy, x, relu ? y : nullptr}))
      bn_bwd_reduce_vec_kernel<T><<<red_grid(n), kVecThreads, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, red_rows(n), sums);
    else
      bn_bwd_reduce_kernel<T><<<reduce_grid(n), block, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, sums);
  } else {
    using T = __nv_bfloat16;
    if (wide_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0}, {dy, x, relu ? y : nullptr})) {
      const int rows = gcd::bwd_partition_rows<T>(n, c);      // the partition of the fused two-phase form: both paths sum the same partials
      launch_pdl(bn_bwd_reduce_wide_kernel<T>, dim3(rows_grid(n, rows)), dim3(kVecThreads), 0, st, (const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, rows, sums);
    } else if (vec_shape_ok(c) && vec_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0}, {dy, x, relu ? y : nullptr}))
      bn_bwd_reduce_vec_kernel<T><<<red_grid(n), kVecThreads, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, red_rows(n), sums);
    else
      bn_bwd_reduce_kernel<T><<<reduce_grid(n), block, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, relu, sums);
  }
  GCD_LAUNCH_CHECK("gcd_bn_backward_reduce");
  return GCD_OK;
}

namespace {
template <typename T>
void launch_bwd_apply(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y, int64_t n, int c,
                      const float* mean, const float* invstd, const float* gamma, const double* sums, int relu, int training, void* dx,
                      int64_t ld_dx, void* dres, int64_t ld_dres, float* dgamma, float* dbeta, cudaStream_t st) {
  if (wide_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0, ld_dx, dres ? ld_dres : 0}, {dy, x, relu ? y : nullptr, dx, dres})) {
    const int rows = wide_rows<T>(n, c);
    launch_pdl(bn_bwd_apply_wide_kernel<T>, dim3(rows_grid(n, rows)), dim3(kVecThreads), 0, st, (const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, (T*)dx, ld_dx, (T*)dres, ld_dres, dgamma, dbeta, rows);
    return;
  }
  const bool vec = vec_ok<T>(c, {ld_dy, ld_x, relu ? ld_y : 0, ld_dx, dres ? ld_dres : 0}, {dy, x, relu ? y : nullptr, dx, dres});
  if (vec && vec_shape_ok(c)) {
    bn_bwd_apply_vec_kernel<T><<<vec_grid(n), kVecThreads, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, (T*)dx, ld_dx, (T*)dres, ld_dres, dgamma, dbeta);
    return;
  }
  const unsigned g = (unsigned)ceil_div(n * ((c + 3) / 4), 256);
  if (vec) bn_bwd_apply_kernel<T, true><<<g, 256, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, (T*)dx, ld_dx, (T*)dres, ld_dres, dgamma, dbeta);
  else bn_bwd_apply_kernel<T, false><<<g, 256, 0, st>>>((const T*)dy, ld_dy, (const T*)x, ld_x, (const T*)y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, (T*)dx, ld_dx, (T*)dres, ld_dres, dgamma, dbeta);
}
}  // namespace

extern "C" int32_t gcd_bn_backward_apply(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                                         int64_t n, int32_t c, const float* mean, const float* invstd, const float* gamma,
                                         const double* sums, int32_t relu, int32_t training, void* dx, int64_t ld_dx, void* dres,
                                         int64_t ld_dres, float* dgamma, float* dbeta, int32_t dtype, void* stream) {
  GCD_REQUIRE(c >= 1 && c <= kMaxChannels, "gcd_bn_backward_apply: channel count %d out of range", c);
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    if (dtype == GCD_F32) launch_bwd_apply<float>(dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, dx, ld_dx, dres, ld_dres, dgamma, dbeta, st);
    else launch_bwd_apply<__nv_bfloat16>(dy, ld_dy, x, ld_x, y, ld_y, n, c, mean, invstd, gamma, sums, relu, training, dx, ld_dx, dres, ld_dres, dgamma, dbeta, st);
  } else if (dgamma || dbeta) {
    bn_param_grad_kernel<<<(unsigned)ceil_div(c, 128), 128, 0, st>>>(sums, c, dgamma, dbeta);
  }
  GCD_LAUNCH_CHECK("gcd_bn_backward_apply");
  return GCD_OK;
}

extern "C" int32_t gcd_relu(const void* x, void* y, int64_t numel, int32_t dtype, void* stream) {
  if (numel == 0) return GCD_OK;
  const unsigned g = (unsigned)ceil_div(numel, 256);
  if (dtype == GCD_F32) relu_kernel<float><<<g, 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, numel);
  else relu_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, numel);
  GCD_LAUNCH_CHECK("gcd_relu");
  return GCD_OK;
}
extern "C" int32_t gcd_relu_backward(const void* dy, const void* y, void* dx, int64_t numel, int32_t dtype, void* stream) {
  if (numel == 0) return GCD_OK;
  const unsigned g = (unsigned)ceil_div(numel, 256);
  if (dtype == GCD_F32) relu_bwd_kernel<float><<<g, 256, 0, as_stream(stream)>>>((const float*)dy, (const float*)y, (float*)dx, numel);
  else relu_bwd_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, numel);
  GCD_LAUNCH_CHECK("gcd_relu_backward");
  return GCD_OK;
}
