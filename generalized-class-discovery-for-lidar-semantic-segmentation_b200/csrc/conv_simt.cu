// conv_simt.cu — fp32 SIMT sparse convolution (forward / dgrad / wgrad) and im2col.
//
// This is the exact-fp32 math mode (GCD_MATH_FP32_SIMT): register-tiled 64x64 output tiles on
// the FMA pipe, used for <=1e-4 parity runs and as the numerical reference for the tcgen05
// kernels in conv_tc.cu.  Forward and dgrad are the same output-stationary kernel: for a tile
// of 64 output rows, loop over the kernel offsets, gather the input rows named by the
// neighbour table (zero rows where the table holds -1, offsets with no hit in the tile are
// skipped), multiply by that offset's [Cin,Cout] slice and accumulate in registers, so every
// output row is written exactly once (no scatter-add, no atomics).
#include "common.cuh"

namespace gcd {
namespace {
constexpr int TM = 64, TN = 64, BK = 16, kConvThreads = 256;
constexpr int AS_LD = TM + 4;

template <typename Tin>
__device__ __forceinline__ void load4(const Tin* __restrict__ p, bool vec_ok, int valid, float (&v)[4]) {
  // loads up to 4 consecutive channels, `valid` of them in range
  if (sizeof(Tin) == 4 && vec_ok && valid == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = j < valid ? to_f32<Tin>(p[j]) : 0.f;
  }
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kConvThreads) conv_fwd_simt_kernel(
    const Tin* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ nbr, int kv, int64_t n_out, int c_in, int c_out,
    const float* __restrict__ w, int64_t ws_k, int64_t ws_c, int64_t ws_n, int mirror, const float* __restrict__ bias,
    Tout* __restrict__ out, int64_t ld_out, double* __restrict__ stats) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][TN];
  __shared__ int s_idx[TM];
  __shared__ float s_sum[TN], s_sq[TN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;
  const bool in_vec_ok = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const bool w_vec_ok = (ws_n == 1) && (ws_c % 4 == 0) && (ws_k % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 2, a_c4 = (tid & 3) * 4;   // A loader: row, 4-channel group within the BK chunk
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;    // B loader

  for (int k = 0; k < kv; ++k) {
    int any = 0;
    if (tid < TM) {
      const int64_t r = row0 + tid;
      int idx = -1;
      if (r < n_out) idx = nbr ? nbr[(int64_t)k * n_out + r] : (int)r;
      s_idx[tid] = idx;
      any = idx >= 0;
    }
    if (!__syncthreads_or(any)) continue;
    const int wk = mirror ? kv - 1 - k : k;
    const float* __restrict__ wkp = w + (int64_t)wk * ws_k;
    const int src = s_idx[a_row];
    for (int c0 = 0; c0 < c_in; c0 += BK) {
      {  // gather A chunk: rows of the tile x channels [c0, c0+BK)
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int c = c0 + a_c4;
        if (src >= 0 && c < c_in) load4<Tin>(in + (int64_t)src * ld_in + c, in_vec_ok, min(4, c_in - c), v);
#pragma unroll
        for (int j = 0; j < 4; ++j) As[a_c4 + j][a_row] = v[j];
      }
      {  // B chunk: W[wk][c0 + b_k][col0 + b_n4 ..]
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int c = c0 + b_k, n = col0 + b_n4;
        if (c < c_in && n < c_out) {
          const float* p = wkp + (int64_t)c * ws_c + (int64_t)n * ws_n;
          if (w_vec_ok && n + 4 <= c_out) {
            float4 t = *reinterpret_cast<const float4*>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < c_out) v[j] = p[(int64_t)j * ws_n];
          }
        }
        *reinterpret_cast<float4*>(&Bs[b_k][b_n4]) = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: bias, store, optional per-channel statistics of the fp32 result
  float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = col0 + tx * 4 + j;
    const float bj = (bias && n < c_out) ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = row0 + ty * 4 + i;
      const float v = acc[i][j] + bj;
      if (r < n_out && n < c_out) {
        out[r * ld_out + n] = from_f32<Tout>(v);
        csum[j] += v; csq[j] += v * v;
      }
    }
  }
  if (stats) {
    if (tid < TN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = csum[j] + __shfl_xor_sync(0xffffffffu, csum[j], 16);
      float q = csq[j] + __shfl_xor_sync(0xffffffffu, csq[j], 16);
      if ((tid & 16) == 0) { atomicAdd(&s_sum[tx * 4 + j], s); atomicAdd(&s_sq[tx * 4 + j], q); }
    }
    __syncthreads();
    if (tid < TN && col0 + tid < c_out) {
      atomicAdd(&stats[col0 + tid], (double)s_sum[tid]);
      atomicAdd(&stats[c_out + col0 + tid], (double)s_sq[tid]);
    }
  }
}

// dW[k][c][n] += sum over the pairs of offset k of in[pair_in[p]][c] * gout[pair_out[p]][n].
// grid.x = persistent workers striding over (offset, pair-chunk) work items, grid.y = 64x64 tile of dW.
constexpr int kWgradChunk = 2048;

template <typename Tin, typename Tg>
__global__ void __launch_bounds__(kConvThreads) conv_wgrad_simt_kernel(
    const Tin* __restrict__ in, int64_t ld_in, const Tg* __restrict__ gout, int64_t ld_g, const int32_t* __restrict__ pair_in,
    const int32_t* __restrict__ pair_out, const int32_t* __restrict__ pair_off, int64_t n_rows_identity, int kv, int c_in,
    int c_out, float* __restrict__ dw) {
  __shared__ __align__(16) float As[BK][TM];
  __shared__ __align__(16) float Gs[BK][TN];
  __shared__ int s_off[128];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_n = (c_out + TN - 1) / TN;
  const int c0 = (blockIdx.y / tiles_n) * TM, n0 = (blockIdx.y % tiles_n) * TN;
  const bool in_vec_ok = (ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const bool g_vec_ok = (ld_g % 4 == 0) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);

  // cumulative work-item counts per offset (kv <= 127)
  if (tid == 0) {
    int cum = 0;
    for (int k = 0; k < kv; ++k) {
      s_off[k] = cum;
      int64_t nk = pair_off ? (int64_t)pair_off[k + 1] - pair_off[k] : n_rows_identity;
      cum += (int)((nk + kWgradChunk - 1) / kWgradChunk);
    }
    s_off[kv] = cum;
  }
  __syncthreads();
  const int total_work = s_off[kv];
  const int l_p = tid >> 4, l_c4 = (tid & 15) * 4;  // loader: pair within the BK step, 4-channel group

  for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
    int k = 0;
    while (s_off[k + 1] <= work) ++k;
    const int64_t p_begin = (pair_off ? (int64_t)pair_off[k] : 0) + (int64_t)(work - s_off[k]) * kWgradChunk;
    const int64_t p_end_k = pair_off ? (int64_t)pair_off[k + 1] : n_rows_identity;
    const int64_t p_end = min(p_begin + kWgradChunk, p_end_k);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
      const int64_t p = p0 + l_p;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
      if (p < p_end) {
        const int64_t ri = pair_in ? pair_in[p] : p, ro = pair_out ? pair_out[p] : p;
        const int c = c0 + l_c4, n = n0 + l_c4;
        if (c < c_in) load4<Tin>(in + ri * ld_in + c, in_vec_ok, min(4, c_in - c), a);
        if (n < c_out) load4<Tg>(gout + ro * ld_g + n, g_vec_ok, min(4, c_out - n), g);
      }
      *reinterpret_cast<float4*>(&As[l_p][l_c4]) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(&Gs[l_p][l_c4]) = make_float4(g[0], g[1], g[2], g[3]);
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 gv4 = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
        const float av[4] = {av4.x, av4.y, av4.z, av4.w}, gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
      }
      __syncthreads();
    }
    float* __restrict__ dwk = dw + (int64_t)k * c_in * c_out;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (c < c_in && n < c_out && acc[i][j] != 0.f) atomicAdd(&dwk[(int64_t)c * c_out + n], acc[i][j]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t n, int c, float* __restrict__ out) {
  // block handles a slab of rows; thread t owns channel t % c of row group t / c
  const int per = 256 / c > 0 ? 256 / c : 1;
  const int ch = threadIdx.x % c, grp = threadIdx.x / c;
  if (c > 256) return;
  float s = 0.f;
  if (grp < per)
    for (int64_t r = (int64_t)blockIdx.x * per + grp; r < n; r += (int64_t)gridDim.x * per) s += to_f32<T>(x[r * ld + ch]);
  if (grp < per && s != 0.f) atomicAdd(&out[ch], s);
}

// A_col[o][k*c_in + c] = in[nbr[k][o]][c] (0 where the table holds -1); row pitch ld_out (zero padded).
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) im2col_kernel(const Tin* __restrict__ in, int64_t ld_in, int c_in, const int32_t* __restrict__ nbr,
                                                      int kv, int64_t n_out, Tout* __restrict__ out, int64_t ld_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_out * ld_out;
  if (t >= total) return;
  const int64_t o = t / ld_out;
  const int col = (int)(t - o * ld_out);
  float v = 0.f;
  if (col < kv * c_in) {
    const int k = col / c_in, c = col - k * c_in;
    const int src = nbr[(int64_t)k * n_out + o];
    if (src >= 0) v = to_f32<Tin>(in[(int64_t)src * ld_in + c]);
  }
  out[t] = from_f32<Tout>(v);
}
// The same through a shared-memory transpose, for the shapes the stem uses (c_in == 1, ld_out <= 128): a block takes 32 output
// rows; warp w reads the table entries of offsets k = w, w + 8, ... for those 32 rows (one 128-byte line per offset, where the
// element-per-thread form above reads one line per *thread*), gathers the input value and parks it in a [32][ld_out] tile that
// is then written row by row, coalesced.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) im2col_c1_kernel(const Tin* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ nbr, int kv,
                                                         int64_t n_out, Tout* __restrict__ out, int ld_out) {
  pdl_trigger(); pdl_wait();
  __shared__ float tile[32][129];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t o0 = (int64_t)blockIdx.x * 32;
  const int64_t o = o0 + lane;
  for (int k = warp; k < ld_out; k += 8) {
    float v = 0.f;
    if (k < kv && o < n_out) {
      const int src = __ldg(&nbr[(int64_t)k * n_out + o]);
      if (src >= 0) v = to_f32<Tin>(in[(int64_t)src * ld_in]);
    }
    tile[lane][k] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * ld_out; idx += 256) {
    const int r = idx / ld_out, c = idx - r * ld_out;
    if (o0 + r < n_out) out[(o0 + r) * ld_out + c] = from_f32<Tout>(tile[r][c]);
  }
}
}  // namespace

int32_t conv_forward_simt(const gcd_conv_args* a, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(a->n_out, TM), (unsigned)ceil_div(a->c_out, TN));
#define GCD_LAUNCH_FWD(TI, TO)                                                                                         \
  conv_fwd_simt_kernel<TI, TO><<<grid, kConvThreads, 0, st>>>((const TI*)a->in, a->ld_in, a->nbr, a->kv, a->n_out, a->c_in, \
                                                              a->c_out, a->w, a->w_stride_k, a->w_stride_c, a->w_stride_n, \
                                                              a->mirror, a->bias, (TO*)a->out, a->ld_out, a->stats)
  if (a->in_dtype == GCD_F32 && a->out_dtype == GCD_F32) GCD_LAUNCH_FWD(float, float);
  else if (a->in_dtype == GCD_BF16 && a->out_dtype == GCD_BF16) GCD_LAUNCH_FWD(__nv_bfloat16, __nv_bfloat16);
  else if (a->in_dtype == GCD_BF16 && a->out_dtype == GCD_F32) GCD_LAUNCH_FWD(__nv_bfloat16, float);
  else GCD_LAUNCH_FWD(float, __nv_bfloat16);
#undef GCD_LAUNCH_FWD
  GCD_LAUNCH_CHECK("gcd_conv_forward(simt)");
  return GCD_OK;
}

int32_t conv_wgrad_simt(const gcd_wgrad_args* a, cudaStream_t st) {
  const int tiles = (int)(ceil_div(a->c_in, TM) * ceil_div(a->c_out, TN));
  int64_t work_bound = ceil_div(a->n_pairs, kWgradChunk) + a->kv;
  dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(work_bound, (int64_t)kNumSMs * 8 / std::max(1, std::min(tiles, 8)))), (unsigned)tiles);
#define GCD_LAUNCH_WG(TI, TG)                                                                                           \
  conv_wgrad_simt_kernel<TI, TG><<<grid, kConvThreads, 0, st>>>((const TI*)a->in, a->ld_in, (const TG*)a->gout, a->ld_gout, \
                                                                a->pair_in, a->pair_out, a->pair_off, a->n_pairs, a->kv,  \
                                                                a->c_in, a->c_out, a->dw)
  if (a->in_dtype == GCD_F32 && a->gout_dtype == GCD_F32) GCD_LAUNCH_WG(float, float);
  else if (a->in_dtype == GCD_BF16 && a->gout_dtype == GCD_BF16) GCD_LAUNCH_WG(__nv_bfloat16, __nv_bfloat16);
  else if (a->in_dtype == GCD_BF16 && a->gout_dtype == GCD_F32) GCD_LAUNCH_WG(__nv_bfloat16, float);
  else GCD_LAUNCH_WG(float, __nv_bfloat16);
#undef GCD_LAUNCH_WG
  GCD_LAUNCH_CHECK("gcd_conv_wgrad(simt)");
  return GCD_OK;
}

int32_t colsum_f32(const void* x, int64_t ld, int64_t n, int c, int dtype, float* out, cudaStream_t st) {
  if (n == 0) return GCD_OK;
  if (c > 256) { set_error("colsum: more than 256 channels unsupported"); return GCD_ERR_UNSUPPORTED; }
  const int per = std::max(1, 256 / c);
  unsigned g = (unsigned)std::min<int64_t>(ceil_div(n, per), kNumSMs * 8);
  if (dtype == GCD_F32) colsum_kernel<float><<<g, 256, 0, st>>>((const float*)x, ld, n, c, out);
  else colsum_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)x, ld, n, c, out);
  GCD_LAUNCH_CHECK("colsum");
  return GCD_OK;
}
}  // namespace gcd

using namespace gcd;

extern "C" int32_t gcd_im2col(const void* in, int64_t ld_in, int32_t c_in, const int32_t* nbr, int32_t kv, int64_t n_out,
                              void* out, int64_t ld_out, int32_t in_dtype, int32_t out_dtype, void* stream) {
  GCD_REQUIRE(ld_out >= (int64_t)kv * c_in, "gcd_im2col: ld_out smaller than kv*c_in");
  if (n_out == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  if (c_in == 1 && ld_out <= 128 && n_out > 0) {          // the 5x5x5 stem on a one-channel input
    const dim3 grid((unsigned)ceil_div(n_out, 32)), block(256);
    cudaError_t e;
    if (in_dtype == GCD_F32 && out_dtype == GCD_F32) e = launch_pdl(im2col_c1_kernel<float, float>, grid, block, 0, st, (const float*)in, ld_in, nbr, kv, n_out, (float*)out, (int)ld_out);
    else if (in_dtype == GCD_F32 && out_dtype == GCD_BF16) e = launch_pdl(im2col_c1_kernel<float, __nv_bfloat16>, grid, block, 0, st, (const float*)in, ld_in, nbr, kv, n_out, (__nv_bfloat16*)out, (int)ld_out);
    else if (in_dtype == GCD_BF16 && out_dtype == GCD_BF16) e = launch_pdl(im2col_c1_kernel<__nv_bfloat16, __nv_bfloat16>, grid, block, 0, st, (const __nv_bfloat16*)in, ld_in, nbr, kv, n_out, (__nv_bfloat16*)out, (int)ld_out);
    else e = launch_pdl(im2col_c1_kernel<__nv_bfloat16, float>, grid, block, 0, st, (const __nv_bfloat16*)in, ld_in, nbr, kv, n_out, (float*)out, (int)ld_out);
    if (e != cudaSuccess) return cuda_fail(e, "gcd_im2col");
    return GCD_OK;
  }
  unsigned g = (unsigned)ceil_div(n_out * ld_out, 256);
  if (in_dtype == GCD_F32 && out_dtype == GCD_F32)
    im2col_kernel<float, float><<<g, 256, 0, st>>>((const float*)in, ld_in, c_in, nbr, kv, n_out, (float*)out, ld_out);
  else if (in_dtype == GCD_F32 && out_dtype == GCD_BF16)
    im2col_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)in, ld_in, c_in, nbr, kv, n_out, (__nv_bfloat16*)out, ld_out);
  else if (in_dtype == GCD_BF16 && out_dtype == GCD_BF16)
    im2col_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, ld_in, c_in, nbr, kv, n_out, (__nv_bfloat16*)out, ld_out);
  else
    im2col_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, ld_in, c_in, nbr, kv, n_out, (float*)out, ld_out);
  GCD_LAUNCH_CHECK("gcd_im2col");
  return GCD_OK;
}
