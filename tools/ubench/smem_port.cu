// Microbenchmark: do the tensor core's operand reads, cp.async (LDGSTS) row copies and bulk (TMA) weight copies share one
// shared-memory port?  DESIGN.md section 4.1 item 5 reads the forward kernel's cost model c = 256 + 1.9 N cycles per stage as
// 2 x (16 KB rows + N x 128 B weights) through 128 B/clk; this program measures the three streams alone and together, with
// no barriers between them, on every SM:
//   role M : one warp issues `iters` stages of 4 x tcgen05.mma (M=128, N, K=16, bf16, SS operands) from a ring of 3 slots
//   role G : `gw` warps issue `iters` x 32 warp-level 16-byte cp.async copies (128 rows x 128 B per "stage") from an
//            L2-resident buffer into a scratch tile laid out like the kernel's operand image; `fill` % of the rows are real,
//            the rest zero fill (src-size 0); with `compact` the zero-fill copies are not issued at all
//   role B : one thread streams `iters` bulk copies of N x 128 B (the weight slice) into a second scratch area
// If the streams were independent, running them together would take max(alone); a shared port shows up as ~sum(alone).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../generalized-class-discovery-for-lidar-semantic-segmentation_b200/csrc -o smem_port smem_port.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace gcd::ptx;

constexpr int kSlots = 3;
constexpr int kABytes = 16384;
constexpr int kMaxWarps = 10;      // warp 0: MMA, warp 1: bulk copies, warps 2..9: gather

struct Params {
  const uint8_t* rows; int n_rows;      // [n_rows][128 B], n_rows a power of two
  const uint8_t* weights;               // >= 3 x N x 128 B
  int n_cols, iters;
  int do_mma, gw, do_bulk;              // which roles run
  int fill_pct, compact;
  long long* out;                       // per CTA: [0] MMA role cycles, [1] gather role cycles (max over warps), [2] bulk role cycles, [3] whole CTA
};

__global__ void __launch_bounds__(kMaxWarps * 32, 1) k(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_bytes = (uint32_t)p.n_cols * 128u;
  uint8_t* a_slots = smem;                                   // kSlots x 16 KB
  uint8_t* b_slots = a_slots + kSlots * kABytes;             // kSlots x b_bytes
  uint8_t* g_scratch = b_slots + kSlots * b_bytes;           // 16 KB: where the gather warps write
  uint8_t* w_scratch = g_scratch + kABytes;                  // 2 x b_bytes: where the bulk copies land
  __shared__ uint64_t bar_done, bar_bulk[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_role[kMaxWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_bulk[0], 1); mbar_init(&bar_bulk[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<512>(&tmem_base_s); tmem_relinquish(); }
  for (uint32_t i = threadIdx.x; i < (kSlots * (kABytes + b_bytes)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < kMaxWarps) t_role[threadIdx.x] = 0;
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const long long t_start = clock64();

  if (warp == 0 && p.do_mma) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_cols, 0, 0);
    const uint64_t da0 = make_smem_desc_sw128(smem_u32(a_slots), 16, 1024);
    const uint64_t db0 = make_smem_desc_sw128(smem_u32(b_slots), 16, 1024);
    const bool leader = elect_one();
    int slot = 0;
    for (int it = 0; it < p.iters; ++it) {
      const uint64_t da = da0 + (uint64_t)(slot * (kABytes >> 4)), db = db0 + (uint64_t)(slot * (b_bytes >> 4));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_bf16_ss_pred(tmem, da + ks * 2, db + ks * 2, idesc, (it | ks) != 0, leader ? 1u : 0u);
      if (++slot == kSlots) slot = 0;
    }
    mma_commit_pred(&bar_done, leader ? 1u : 0u);
    mbar_wait(&bar_done, 0);                                  // all MMAs retired
    if (lane == 0) t_role[0] = clock64() - t_start;
  } else if (warp == 1 && p.do_bulk) {
    if (lane == 0) {
      uint32_t ph[2] = {0, 0};
      for (int it = 0; it < p.iters; ++it) {
        const int b = it & 1;
        if (it >= 2) { mbar_wait(&bar_bulk[b], ph[b]); ph[b] ^= 1; }      // two copies in flight
        mbar_arrive_expect_tx(&bar_bulk[b], b_bytes);
        bulk_g2s(smem_u32(w_scratch + b * b_bytes), p.weights + (size_t)(it % 3) * b_bytes, b_bytes, &bar_bulk[b]);
      }
      for (int b = 0; b < 2; ++b) if (p.iters > b) mbar_wait(&bar_bulk[b], ph[b]);
      t_role[1] = clock64() - t_start;
    }
  } else if (warp >= 2 && warp < 2 + p.gw) {
    // each gather warp fills whole "stages" (32 warp-level copies = 128 rows x 128 B), stages dealt round-robin
    const int chunk = lane & 7, rsub = lane >> 3;
    const uint32_t scratch = smem_u32(g_scratch);
    const int gw_id = warp - 2;
    uint32_t rng = 0x9e3779b9u * (uint32_t)(blockIdx.x * 16 + gw_id + 1);
    for (int it = gw_id; it < p.iters; it += p.gw) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const uint32_t row = (uint32_t)(4 * j + rsub);
        // pseudo-random source row and validity, identical for the 8 lanes of a row
        uint32_t h = (rng ^ (uint32_t)(it * 131 + 4 * j + rsub)) * 0x85ebca6bu; h ^= h >> 15; h *= 0xc2b2ae35u; h ^= h >> 16;
        const bool real = (int)(h % 100u) < p.fill_pct;
        const uint8_t* src = p.rows + (size_t)((h >> 8) & (uint32_t)(p.n_rows - 1)) * 128u + chunk * 16;
        const uint32_t dst = scratch + (row >> 3) * 1024u + (row & 7u) * 128u + (((uint32_t)chunk ^ (row & 7u)) << 4);
        if (p.compact) { if (__any_sync(0xffffffffu, real)) cp_async_16_pred(dst, src, 16u, real ? 1u : 0u); }
        else cp_async_16(dst, src, real ? 16u : 0u);
      }
      cp_async_commit();
      cp_async_wait<4>();
    }
    cp_async_wait_all();
    if (lane == 0) t_role[warp] = clock64() - t_start;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long g = 0;
    for (int w = 2; w < kMaxWarps; ++w) g = t_role[w] > g ? t_role[w] : g;
    long long* o = p.out + (size_t)blockIdx.x * 4;
    o[0] = t_role[0]; o[1] = g; o[2] = t_role[1]; o[3] = clock64() - t_start;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  const int n_sm = 148, iters = 2000, n_rows = 1 << 16;
  uint8_t *rows, *weights; long long* out;
  cudaMalloc(&rows, (size_t)n_rows * 128); cudaMemset(rows, 0x3c, (size_t)n_rows * 128);
  cudaMalloc(&weights, 3 * 256 * 128); cudaMemset(weights, 0x3c, 3 * 256 * 128);
  cudaMalloc(&out, n_sm * 4 * sizeof(long long));
  const size_t smem = 1024 + kSlots * (kABytes + 256 * 128) + kABytes + 2 * 256 * 128;    // 225 KB at N = 256
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("cycles per stage (average over %d CTAs, %d stages); tensor floor = 2 N\n", n_sm, iters);
  printf("%5s %5s %3s %5s %5s %8s | %9s %9s %9s %9s\n", "N", "mma", "gw", "bulk", "fill%", "compact", "mma", "gather", "bulk", "cta");
  struct Cfg { int mma, gw, bulk, fill, compact; };
  const Cfg cfgs[] = {{1, 0, 0, 100, 0}, {0, 8, 0, 100, 0}, {0, 8, 0, 25, 0}, {0, 8, 0, 25, 1}, {0, 0, 1, 100, 0},
                      {1, 8, 0, 100, 0}, {1, 8, 0, 25, 0}, {1, 8, 0, 25, 1}, {1, 0, 1, 100, 0}, {1, 8, 1, 100, 0}, {1, 8, 1, 25, 0}, {1, 8, 1, 25, 1},
                      {1, 4, 1, 25, 0}, {1, 4, 1, 25, 1}};
  for (int n : {32, 96, 128, 256}) {
    for (const Cfg& c : cfgs) {
      Params p{rows, n_rows, weights, n, iters, c.mma, c.gw, c.bulk, c.fill, c.compact, out};
      k<<<n_sm, kMaxWarps * 32, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      static long long h[148 * 4];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double s[4] = {0, 0, 0, 0};
      for (int b = 0; b < n_sm; ++b) for (int i = 0; i < 4; ++i) s[i] += (double)h[b * 4 + i];
      printf("%5d %5d %3d %5d %5d %8d | %9.1f %9.1f %9.1f %9.1f\n", n, c.mma, c.gw, c.bulk, c.fill, c.compact,
             s[0] / n_sm / iters, s[1] / n_sm / iters, s[2] / n_sm / iters, s[3] / n_sm / iters);
    }
  }
  return 0;
}
