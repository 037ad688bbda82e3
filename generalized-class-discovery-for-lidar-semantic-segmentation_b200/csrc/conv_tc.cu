// conv_tc.cu — bf16 tcgen05/TMEM sparse convolution for sm_100a (forward and dgrad share one
// kernel; wgrad is the second kernel).
//
// Forward / dgrad: persistent, warp-specialised, output-stationary implicit GEMM.
//   tile      = 128 output rows x N (<= 256) output channels, fp32 accumulator in TMEM
//               (two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1)
//   iteration = (kernel offset k with at least one hit in the tile) x (64-channel slice of Cin)
//   warps 0-3 : producers.  Read the tile's column-major neighbour table once into shared
//               memory, then per iteration gather 128 input rows x 64 bf16 with 16-byte
//               cp.async (zero fill for missing neighbours) straight into the 128B-swizzled
//               K-major operand image; one thread fetches the matching pre-packed weight slice
//               with a single bulk copy on the TMA engine (cp.async.bulk -> mbarrier tx count).
//   warp 8    : one lane issues tcgen05.mma (M=128, N, K=16) per 16 channels and commits to the
//               stage's "empty" barrier; after the tile's last iteration it commits to
//               "tmem_full".
//   warps 4-7 : epilogue.  tcgen05.ld the accumulator (lane == output row), add bias, convert,
//               store the row; then release the accumulator buffer.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gcd {
namespace {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                 // bf16 per 128-byte operand row
constexpr int kRowBytes = 128;
constexpr int kABytes = kTileM * kRowBytes; // 16 KB
constexpr int kProducerThreads = 128;
constexpr int kEpilogueThreads = 128;
constexpr int kTcThreads = kProducerThreads + kEpilogueThreads + 32;
constexpr int kMaxKV = 27;
constexpr int kMaxStages = 8;
constexpr int kLag = 2;                     // cp.async groups a producer thread keeps in flight
constexpr int kTileRing = 16;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // TMEM columns between the two accumulator buffers
constexpr int kSmemBudget = 226 * 1024;

struct FwdParams {
  const __nv_bfloat16* in; int64_t ld_in;
  const int32_t* nbr; int kv; int64_t n_out;
  int c_in, c_out;                 // c_in % 16 == 0, c_out % 16 == 0
  int n_tile_cols;                 // columns per N tile (<= 256, % 16 == 0)
  int n_tiles_n;
  const uint8_t* w_packed;         // [kv][nq][c_out rows][128 B] swizzled K-major images
  int mirror;
  const float* bias;
  void* out; int64_t ld_out; int out_is_bf16;
  int stages;
};

struct SmemLayout {
  uint32_t a_off, b_off, nbr_off, bar_off, total;
  uint32_t b_bytes;
};
__host__ __device__ inline SmemLayout make_layout(int stages, int n_tile_cols) {
  SmemLayout L;
  L.b_bytes = (uint32_t)n_tile_cols * kRowBytes;
  L.a_off = 0;
  L.b_off = L.a_off + (uint32_t)stages * kABytes;
  L.nbr_off = L.b_off + (uint32_t)stages * L.b_bytes;
  L.bar_off = L.nbr_off + kMaxKV * kTileM * 4;
  L.total = L.bar_off + 512;
  return L;
}

__global__ void __launch_bounds__(kTcThreads, 1) conv_fwd_tc_kernel(const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is at least 16-byte aligned; the swizzle pattern needs 1024.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SmemLayout L = make_layout(p.stages, p.n_tile_cols);
  int32_t* s_nbr = reinterpret_cast<int32_t*>(smem + L.nbr_off);            // [kv][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full_bar = bars;                    // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;      // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* s_tmem_base = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* s_any = s_tmem_base + 1;            // [4] per-producer-warp offset masks
  int32_t* s_iters = reinterpret_cast<int32_t*>(s_any + 4);  // [kTileRing]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int nq = (p.c_in + kChunkK - 1) / kChunkK;
  const int64_t tiles_m = (p.n_out + kTileM - 1) / kTileM;
  const int64_t n_work = tiles_m * p.n_tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducerThreads); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kEpilogueThreads); }
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc<kTmemCols>(s_tmem_base); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_base;

  if (warp < 4) {
    // ===================================================================== producers
    const int t = threadIdx.x;                 // 0..127 = row of the tile for the table load
    const int chunk = lane & 7, rsub = lane >> 3;
    uint32_t st_issue = 0, ph_issue = 0;       // stage being filled, its parity for the empty barrier
    uint32_t st_arrive = 0;                    // stage whose completion is signalled next
    uint32_t in_flight = 0;                    // committed-but-not-signalled groups of this thread
    uint32_t tile_seq = 0;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x, ++tile_seq) {
      const int64_t tm = work / p.n_tiles_n;
      const int tn = (int)(work - tm * p.n_tiles_n);
      const int64_t row0 = tm * kTileM;
      named_bar_sync(1, kProducerThreads);     // previous tile's table no longer needed by anyone
      uint32_t my_mask = 0;
      {
        const int64_t r = row0 + t;
        for (int k = 0; k < p.kv; ++k) {
          int idx = -1;
          if (r < p.n_out) idx = p.nbr ? __ldg(&p.nbr[(int64_t)k * p.n_out + r]) : (int)r;
          s_nbr[k * kTileM + t] = idx;
          if (__any_sync(0xffffffffu, idx >= 0)) my_mask |= 1u << k;
        }
        if (lane == 0) s_any[warp] = my_mask;
      }
      named_bar_sync(1, kProducerThreads);
      uint32_t mask = s_any[0] | s_any[1] | s_any[2] | s_any[3];
      if (mask == 0) mask = 1;                 // degenerate tile: run offset 0 with all-zero rows
      if (t == 0) s_iters[tile_seq & (kTileRing - 1)] = __popc(mask) * nq;

      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int wk = p.mirror ? p.kv - 1 - k : k;
        for (int q = 0; q < nq; ++q) {
          mbar_wait(&empty_bar[st_issue], ph_issue ^ 1);
          const int width = min(kChunkK, p.c_in - q * kChunkK);   // channels in this slice (multiple of 16)
          if (t == 0) {
            const uint32_t bytes = (uint32_t)p.n_tile_cols * kRowBytes;
            const uint8_t* src = p.w_packed + ((int64_t)(wk * nq + q) * p.c_out + (int64_t)tn * p.n_tile_cols) * kRowBytes;
            mbar_expect_tx(&full_bar[st_issue], bytes);
            bulk_g2s(smem_u32(smem + L.b_off + st_issue * L.b_bytes), src, bytes, &full_bar[st_issue]);
          }
          const uint32_t a_stage = smem_u32(smem + L.a_off + st_issue * kABytes);
          if (chunk * 8 < width) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = j * 16 + warp * 4 + rsub;
              const int idx = s_nbr[k * kTileM + r];
              const __nv_bfloat16* src = p.in + (int64_t)(idx >= 0 ? idx : 0) * p.ld_in + q * kChunkK + chunk * 8;
              cp_async_16(a_stage + r * kRowBytes + ((chunk ^ (r & 7)) << 4), src, idx >= 0 ? 16u : 0u);
            }
          }
          cp_async_commit();
          ++in_flight;
          if (++st_issue == (uint32_t)S) { st_issue = 0; ph_issue ^= 1; }
          if (in_flight > kLag) {
            cp_async_wait<kLag>();
            fence_proxy_async();
            mbar_arrive(&full_bar[st_arrive]);
            if (++st_arrive == (uint32_t)S) st_arrive = 0;
            --in_flight;
          }
        }
      }
    }
    // drain
    cp_async_wait<0>();
    fence_proxy_async();
    while (in_flight) {
      mbar_arrive(&full_bar[st_arrive]);
      if (++st_arrive == (uint32_t)S) st_arrive = 0;
      --in_flight;
    }
  } else if (warp == 8) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_tile_cols, 0, 0);
      uint32_t st = 0, ph = 0, tile_seq = 0;
      for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x, ++tile_seq) {
        const uint32_t buf = tile_seq & 1;
        mbar_wait(&tmem_empty[buf], ((tile_seq >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kAccStride;
        int n_iters = 1;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          if (it == 0) n_iters = s_iters[tile_seq & (kTileRing - 1)];
          const int q = it % nq;
          const int ksteps = min(kChunkK, p.c_in - q * kChunkK) / 16;
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem + L.a_off + st * kABytes), 16, 1024);
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem + L.b_off + st * L.b_bytes), 16, 1024);
          for (int ks = 0; ks < ksteps; ++ks)
            mma_bf16_ss(tmem_d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (it | ks) != 0);
          mma_commit(&empty_bar[st]);
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
        }
        mma_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue
    const int ew = warp - 4;                   // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t tile_seq = 0;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x, ++tile_seq) {
      const int64_t tm = work / p.n_tiles_n;
      const int tn = (int)(work - tm * p.n_tiles_n);
      const uint32_t buf = tile_seq & 1;
      mbar_wait(&tmem_full[buf], (tile_seq >> 1) & 1);
      tc_fence_after();
      const int64_t row = tm * kTileM + ew * 32 + lane;
      const int col0 = tn * p.n_tile_cols;
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)(ew * 32) << 16);
      for (int c = 0; c < p.n_tile_cols; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c, v);
        tmem_ld_wait();
        if (row < p.n_out) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (p.bias ? p.bias[col0 + c + j] : 0.f);
          if (p.out_is_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + col0 + c;
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
            float* o = reinterpret_cast<float*>(p.out) + row * p.ld_out + col0 + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(o + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

// ------------------------------------------------------------------------ weight packing
// Image of (offset k, 64-channel slice q): rows = output channel n (c_out of them), 128 B per row,
// element (n, kk) at n*128 + (((kk>>3) ^ (n&7))<<4) + (kk&7)*2; value = B_k[q*64+kk][n].
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int kv, int c_in, int c_out, int transpose,
                                                            int mirror, __nv_bfloat16* __restrict__ packed) {
  // logical operand: K' x N' where (K', N') = (c_in, c_out) for forward, (c_out, c_in) for dgrad
  const int kdim = transpose ? c_out : c_in, ndim = transpose ? c_in : c_out;
  const int nq = (kdim + kChunkK - 1) / kChunkK;
  const int64_t total = (int64_t)kv * nq * ndim * kChunkK;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % kChunkK);
  const int n = (int)((t / kChunkK) % ndim);
  const int q = (int)((t / ((int64_t)kChunkK * ndim)) % nq);
  const int k = (int)(t / ((int64_t)kChunkK * ndim * nq));
  const int c = q * kChunkK + kk;
  float v = 0.f;
  if (c < kdim) {
    const int ksrc = (transpose && mirror) ? kv - 1 - k : k;
    v = transpose ? w[((int64_t)ksrc * c_in + n) * c_out + c] : w[((int64_t)ksrc * c_in + c) * c_out + n];
  }
  const int64_t img = ((int64_t)k * nq + q) * ndim * kChunkK;  // elements
  const int64_t off = img + (int64_t)n * kChunkK + ((((kk >> 3) ^ (n & 7)) << 3) + (kk & 7));
  packed[off] = __float2bfloat16_rn(v);
}

// All kernels of a model in one launch: block b works on descriptor d with block_start[d] <= b < block_start[d+1].
struct PackDesc {
  const float* w;
  __nv_bfloat16* dst;
  int32_t kv, c_in, c_out, transpose, mirror;
  int32_t block_start;
};
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackDesc* __restrict__ descs, int n_descs) {
  int lo = 0, hi = n_descs - 1;
  while (lo < hi) {                       // last descriptor whose block_start <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackDesc d = descs[lo];
  const int kdim = d.transpose ? d.c_out : d.c_in, ndim = d.transpose ? d.c_in : d.c_out;
  const int nq = (kdim + kChunkK - 1) / kChunkK;
  const int64_t total = (int64_t)d.kv * nq * ndim * kChunkK;
  const int64_t t = (int64_t)(blockIdx.x - d.block_start) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % kChunkK);
  const int n = (int)((t / kChunkK) % ndim);
  const int q = (int)((t / ((int64_t)kChunkK * ndim)) % nq);
  const int k = (int)(t / ((int64_t)kChunkK * ndim * nq));
  const int c = q * kChunkK + kk;
  float v = 0.f;
  if (c < kdim) {
    const int ksrc = (d.transpose && d.mirror) ? d.kv - 1 - k : k;
    v = d.transpose ? d.w[((int64_t)ksrc * d.c_in + n) * d.c_out + c] : d.w[((int64_t)ksrc * d.c_in + c) * d.c_out + n];
  }
  const int64_t img = ((int64_t)k * nq + q) * ndim * kChunkK;
  d.dst[img + (int64_t)n * kChunkK + ((((kk >> 3) ^ (n & 7)) << 3) + (kk & 7))] = __float2bfloat16_rn(v);
}
}  // namespace


// ======================================================================================= wgrad
// dW[k] (Cin x Cout, fp32) += A_k^T G_k over the pairs of offset k.  Work item = (offset k, chunk of
// pairs, 128-channel tile of Cin).  Per stage the producers gather 64 pairs: the input rows
// (their 128-channel slice) and the output-gradient rows (all Cout channels) land as 128-byte
// swizzled rows; with the *pair* index as the MMA K dimension both images are MN-major operands:
//   D[128 (Cin slice) x Cout] += A^T[128 x 16 pairs] * G[16 pairs x Cout]   (4 MMAs per stage)
// Rows of A^T beyond Cin read stale shared memory; they only produce accumulator lanes that the
// epilogue never reads.  The epilogue adds the accumulator into dW with fp32 atomics.
namespace {
constexpr int kWgPairs = 64;                         // pairs (MMA K) per stage
constexpr int kSlabBytes = kWgPairs * kRowBytes;     // 8 KB: 64 rows x 64 channels
constexpr int kWgMaxOffsets = 128;

struct WgParams {
  const __nv_bfloat16* in; int64_t ld_in;
  const __nv_bfloat16* gout; int64_t ld_g;
  const int32_t* pair_in; const int32_t* pair_out; const int32_t* pair_off;
  int64_t n_rows_identity;
  int kv, c_in, c_out;
  int m_tiles;          // ceil(c_in / 128)
  int g_slabs;          // ceil(c_out / 64)
  int chunk;            // pairs per work item (multiple of 64)
  float* dw;
  int stages;
};

__global__ void __launch_bounds__(kTcThreads, 1) conv_wgrad_tc_kernel(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const uint32_t stage_bytes = (uint32_t)(2 + p.g_slabs) * kSlabBytes;   // A: 2 slabs, G: g_slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (uint32_t)S * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* s_tmem_base = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int32_t* s_cum = reinterpret_cast<int32_t*>(s_tmem_base + 4);        // [kv + 1] cumulative chunk counts

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducerThreads); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kEpilogueThreads); }
    fence_mbar_init();
    int cum = 0;
    for (int k = 0; k < p.kv; ++k) {
      s_cum[k] = cum;
      const int64_t nk = p.pair_off ? (int64_t)p.pair_off[k + 1] - p.pair_off[k] : p.n_rows_identity;
      cum += (int)((nk + p.chunk - 1) / p.chunk);
    }
    s_cum[p.kv] = cum;
  }
  if (warp == 8) { tmem_alloc<kTmemCols>(s_tmem_base); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_base;
  const int64_t n_work = (int64_t)s_cum[p.kv] * p.m_tiles;

  // work -> (k, first pair, last pair, m tile)
  auto decode = [&](int64_t work, int& k, int64_t& p_begin, int64_t& p_end, int& mt) {
    mt = (int)(work % p.m_tiles);
    const int item = (int)(work / p.m_tiles);
    k = 0;
    while (s_cum[k + 1] <= item) ++k;
    const int64_t base = p.pair_off ? (int64_t)p.pair_off[k] : 0;
    const int64_t end_k = p.pair_off ? (int64_t)p.pair_off[k + 1] : p.n_rows_identity;
    p_begin = base + (int64_t)(item - s_cum[k]) * p.chunk;
    p_end = min(p_begin + (int64_t)p.chunk, end_k);
  };

  if (warp < 4) {
    // ===================================================================== producers
    const int chunk16 = lane & 7, rsub = lane >> 3;
    uint32_t st_issue = 0, ph_issue = 0, st_arrive = 0, in_flight = 0;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
      int k, mt; int64_t p_begin, p_end;
      decode(work, k, p_begin, p_end, mt);
      const int c_base = mt * 128;
      for (int64_t p0 = p_begin; p0 < p_end; p0 += kWgPairs) {
        int ri[4], ro[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t pp = p0 + j * 16 + warp * 4 + rsub;
          if (pp < p_end) {
            ri[j] = p.pair_in ? __ldg(&p.pair_in[pp]) : (int)pp;
            ro[j] = p.pair_out ? __ldg(&p.pair_out[pp]) : (int)pp;
          } else { ri[j] = -1; ro[j] = -1; }
        }
        mbar_wait(&empty_bar[st_issue], ph_issue ^ 1);
        const uint32_t a_stage = smem_u32(smem + st_issue * stage_bytes);
        const uint32_t g_stage = a_stage + 2 * kSlabBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = j * 16 + warp * 4 + rsub;
          const uint32_t row_off = r * kRowBytes + ((chunk16 ^ (r & 7)) << 4);
          const __nv_bfloat16* src_a = p.in + (int64_t)(ri[j] >= 0 ? ri[j] : 0) * p.ld_in + c_base + chunk16 * 8;
          const __nv_bfloat16* src_g = p.gout + (int64_t)(ro[j] >= 0 ? ro[j] : 0) * p.ld_g + chunk16 * 8;
          const uint32_t nb = ri[j] >= 0 ? 16u : 0u;
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (c_base + s * 64 + chunk16 * 8 < p.c_in) cp_async_16(a_stage + s * kSlabBytes + row_off, src_a + s * 64, nb);
          for (int s = 0; s < p.g_slabs; ++s)
            if (s * 64 + chunk16 * 8 < p.c_out) cp_async_16(g_stage + s * kSlabBytes + row_off, src_g + s * 64, nb);
        }
        cp_async_commit();
        ++in_flight;
        if (++st_issue == (uint32_t)S) { st_issue = 0; ph_issue ^= 1; }
        if (in_flight > kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async();
          mbar_arrive(&full_bar[st_arrive]);
          if (++st_arrive == (uint32_t)S) st_arrive = 0;
          --in_flight;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (in_flight) {
      mbar_arrive(&full_bar[st_arrive]);
      if (++st_arrive == (uint32_t)S) st_arrive = 0;
      --in_flight;
    }
  } else if (warp == 8) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16((uint32_t)p.c_out, 1, 1);   // both operands MN-major
      uint32_t st = 0, ph = 0, seq = 0;
      for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x, ++seq) {
        int k, mt; int64_t p_begin, p_end;
        decode(work, k, p_begin, p_end, mt);
        const uint32_t buf = seq & 1;
        mbar_wait(&tmem_empty[buf], ((seq >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kAccStride;
        const int n_stages = (int)((p_end - p_begin + kWgPairs - 1) / kWgPairs);
        for (int it = 0; it < n_stages; ++it) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + st * stage_bytes);
          const uint32_t g_addr = a_addr + 2 * kSlabBytes;
          for (int ks = 0; ks < kWgPairs / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 16 * kRowBytes, kSlabBytes, 1024);
            const uint64_t db = make_smem_desc_sw128(g_addr + ks * 16 * kRowBytes, kSlabBytes, 1024);
            mma_bf16_ss(tmem_d, da, db, idesc, (it | ks) != 0);
          }
          mma_commit(&empty_bar[st]);
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
        }
        mma_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue
    const int ew = warp - 4;
    uint32_t seq = 0;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x, ++seq) {
      int k, mt; int64_t p_begin, p_end;
      decode(work, k, p_begin, p_end, mt);
      const uint32_t buf = seq & 1;
      mbar_wait(&tmem_full[buf], (seq >> 1) & 1);
      tc_fence_after();
      const int c = mt * 128 + ew * 32 + lane;          // input channel = accumulator lane
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)(ew * 32) << 16);
      float* dst = p.dw + ((int64_t)k * p.c_in + c) * p.c_out;
      for (int n0 = 0; n0 < p.c_out; n0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + n0, v);
        tmem_ld_wait();
        if (c < p.c_in) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + n0 + j, __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}
}  // namespace

bool conv_wgrad_tc_supported(const gcd_wgrad_args* a) {
  return a->in_dtype == GCD_BF16 && a->gout_dtype == GCD_BF16 && a->c_in % 16 == 0 && a->c_out % 16 == 0 && a->c_in >= 16 &&
         a->c_out >= 16 && a->c_out <= 256 && a->kv < kWgMaxOffsets && a->ld_in % 8 == 0 && a->ld_gout % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->gout) & 15) == 0;
}

int32_t conv_wgrad_tc(const gcd_wgrad_args* a, cudaStream_t st) {
  WgParams p;
  p.in = (const __nv_bfloat16*)a->in; p.ld_in = a->ld_in; p.gout = (const __nv_bfloat16*)a->gout; p.ld_g = a->ld_gout;
  p.pair_in = a->pair_in; p.pair_out = a->pair_out; p.pair_off = a->pair_off; p.n_rows_identity = a->n_pairs;
  p.kv = a->kv; p.c_in = a->c_in; p.c_out = a->c_out; p.dw = a->dw;
  p.m_tiles = (a->c_in + 127) / 128;
  p.g_slabs = (a->c_out + 63) / 64;
  // n_pairs is an upper bound when pair lists are used (the exact count lives on the device);
  // real kernel maps fill about a quarter of the table.
  const int64_t expect = a->pair_in ? std::max<int64_t>(a->n_pairs / 4, 1) : a->n_pairs;
  int64_t chunk = ceil_div(expect, (int64_t)kNumSMs * 2);
  chunk = std::max<int64_t>(512, std::min<int64_t>(8192, ceil_div(chunk, kWgPairs) * kWgPairs));
  p.chunk = (int)chunk;
  const int stage_bytes = (2 + p.g_slabs) * kSlabBytes;
  int stages = std::min(kMaxStages, (kSmemBudget - 1024 - 1024) / stage_bytes);
  if (stages <= kLag) { set_error("conv_wgrad_tc: not enough shared memory stages"); return GCD_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_wgrad_tc_kernel)");
    attr_set = true;
  }
  const int64_t work_bound = (ceil_div(a->n_pairs, chunk) + a->kv) * p.m_tiles;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(work_bound, kNumSMs));
  conv_wgrad_tc_kernel<<<grid, kTcThreads, smem, st>>>(p);
  GCD_LAUNCH_CHECK("gcd_conv_wgrad(tcgen05)");
  return GCD_OK;
}

bool conv_forward_tc_supported(const gcd_conv_args* a) {
  return a->in_dtype == GCD_BF16 && a->c_in % 16 == 0 && a->c_out % 16 == 0 && a->c_in >= 16 && a->c_out >= 16 &&
         a->c_out <= 512 && a->kv <= kMaxKV && a->w_packed != nullptr && a->ld_in % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->in) & 15) == 0 &&
         (a->out_dtype == GCD_BF16 ? (a->ld_out % 8 == 0) : (a->ld_out % 4 == 0)) && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0;
}

int32_t conv_forward_tc(const gcd_conv_args* a, cudaStream_t st) {
  FwdParams p;
  p.in = (const __nv_bfloat16*)a->in; p.ld_in = a->ld_in; p.nbr = a->nbr; p.kv = a->kv; p.n_out = a->n_out;
  p.c_in = a->c_in; p.c_out = a->c_out;
  p.n_tiles_n = (a->c_out + 255) / 256;
  p.n_tile_cols = a->c_out / p.n_tiles_n;
  if (p.n_tile_cols % 16 != 0 || p.n_tile_cols * p.n_tiles_n != a->c_out) { set_error("conv_forward_tc: cannot split %d output channels into equal tiles", a->c_out); return GCD_ERR_UNSUPPORTED; }
  p.w_packed = (const uint8_t*)a->w_packed; p.mirror = 0;  // mirroring is baked into the packed image
  p.bias = a->bias; p.out = a->out; p.ld_out = a->ld_out; p.out_is_bf16 = a->out_dtype == GCD_BF16;
  const int stage_bytes = kABytes + p.n_tile_cols * kRowBytes;
  int stages = (kSmemBudget - 1024 - kMaxKV * kTileM * 4 - 512) / stage_bytes;
  stages = std::max(2, std::min(stages, kMaxStages));
  if (stages <= kLag) { set_error("conv_forward_tc: not enough shared memory stages"); return GCD_ERR_UNSUPPORTED; }
  p.stages = stages;
  const SmemLayout L = make_layout(stages, p.n_tile_cols);
  const size_t smem = L.total + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_fwd_tc_kernel)");
    attr_set = true;
  }
  const int64_t n_work = ceil_div(a->n_out, kTileM) * p.n_tiles_n;
  const unsigned grid = (unsigned)std::min<int64_t>(n_work, kNumSMs);
  conv_fwd_tc_kernel<<<grid, kTcThreads, smem, st>>>(p);
  GCD_LAUNCH_CHECK("gcd_conv_forward(tcgen05)");
  return GCD_OK;
}
}  // namespace gcd

using namespace gcd;

extern "C" size_t gcd_conv_packed_weight_bytes(int32_t kv, int32_t c_in, int32_t c_out) {
  // large enough for either orientation
  const int64_t a = (int64_t)kv * ((c_in + 63) / 64) * c_out * 128;
  const int64_t b = (int64_t)kv * ((c_out + 63) / 64) * c_in * 128;
  return (size_t)std::max(a, b);
}

extern "C" int32_t gcd_conv_pack_weights(const float* w, int32_t kv, int32_t c_in, int32_t c_out, int32_t transpose, int32_t mirror,
                                         void* packed, void* stream) {
  GCD_REQUIRE(w && packed && kv >= 1 && c_in >= 1 && c_out >= 1, "gcd_conv_pack_weights: bad arguments");
  const int kdim = transpose ? c_out : c_in, ndim = transpose ? c_in : c_out;
  const int64_t total = (int64_t)kv * ((kdim + 63) / 64) * ndim * 64;
  pack_weights_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, kv, c_in, c_out, transpose, mirror, (__nv_bfloat16*)packed);
  GCD_LAUNCH_CHECK("gcd_conv_pack_weights");
  return GCD_OK;
}

extern "C" int32_t gcd_conv_pack_weights_batched(const void* descs, int32_t n_descs, int32_t total_blocks, void* stream) {
  GCD_REQUIRE(descs && n_descs >= 1 && total_blocks >= 1, "gcd_conv_pack_weights_batched: bad arguments");
  static_assert(sizeof(PackDesc) == 40, "PackDesc layout is part of the C ABI (gcd_pack_desc)");
  pack_weights_batched_kernel<<<(unsigned)total_blocks, 256, 0, as_stream(stream)>>>((const PackDesc*)descs, n_descs);
  GCD_LAUNCH_CHECK("gcd_conv_pack_weights_batched");
  return GCD_OK;
}

extern "C" int32_t gcd_conv_tc_supported(int32_t c_in, int32_t c_out, int32_t kv) {
  return c_in % 16 == 0 && c_out % 16 == 0 && c_in >= 16 && c_out >= 16 && c_out <= 512 && kv <= 27;
}
