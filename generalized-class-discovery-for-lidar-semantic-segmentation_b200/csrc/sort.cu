// sort.cu — stable LSD radix sort of (uint64 key, int32 payload) pairs, 8 bits per pass.
// Used only for the np.unique-ordered quantiser flavour (models/voxelizer.py:334-360), whose
// voxels are numbered by ascending key; M is the number of voxels of one scan (1e4..1e6).
#include "common.cuh"

namespace gcd {
namespace {
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 8;
constexpr int kSortTile = kSortThreads * kSortRounds;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                   int32_t* __restrict__ counts, int n_blocks) {
  __shared__ int hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = base + (int64_t)r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(int)((keys[i] >> shift) & 0xff)], 1);
  }
  __syncthreads();
  counts[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                      const int32_t* __restrict__ vals_in, int64_t n, int shift,
                                                                      const int32_t* __restrict__ offsets, int n_blocks,
                                                                      uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ int warp_cnt[kSortThreads / 32][256];
  __shared__ int running[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = offsets[(int64_t)threadIdx.x * n_blocks + blockIdx.x];
  for (int w = 0; w < kSortThreads / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + (int64_t)r * kSortThreads + threadIdx.x;
    const bool valid = i < n;
    uint64_t key = valid ? keys_in[i] : 0;
    int val = valid ? vals_in[i] : 0;
    int d = valid ? (int)((key >> shift) & 0xff) : 256 + lane;  // invalid lanes never match
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) warp_cnt[warp][d] = __popc(peers);
    __syncthreads();
    {  // thread t owns digit t: turn per-warp counts into per-warp starting positions
      int off = running[threadIdx.x];
      for (int w = 0; w < kSortThreads / 32; ++w) {
        int c = warp_cnt[w][threadIdx.x];
        warp_cnt[w][threadIdx.x] = off;
        off += c;
      }
      running[threadIdx.x] = off;
    }
    __syncthreads();
    if (valid) {
      int pos = warp_cnt[warp][d] + rank_in_warp;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    for (int w = 0; w < kSortThreads / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
  }
}
}  // namespace

size_t radix_sort_workspace_bytes(int64_t n) {
  int64_t n_blocks = ceil_div(n > 0 ? n : 1, kSortTile);
  size_t b = 0;
  b += align_up((size_t)n * sizeof(uint64_t), 256);            // alternate keys
  b += align_up((size_t)n * sizeof(int32_t), 256);             // alternate vals
  b += align_up((size_t)n_blocks * 256 * sizeof(int32_t), 256);  // counts
  b += scan_workspace_bytes(n_blocks * 256);
  return b;
}

// Sorts in place (keys, vals); key bits [0, key_bits) are significant.
int32_t radix_sort_pairs(uint64_t* keys, int32_t* vals, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
  if (n <= 1) return GCD_OK;
  if (workspace_bytes < radix_sort_workspace_bytes(n)) {
    set_error("radix_sort_pairs: workspace too small");
    return GCD_ERR_WORKSPACE;
  }
  const int n_blocks = (int)ceil_div(n, kSortTile);
  char* p = static_cast<char*>(workspace);
  uint64_t* keys_alt = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)n * sizeof(uint64_t), 256);
  int32_t* vals_alt = reinterpret_cast<int32_t*>(p);   p += align_up((size_t)n * sizeof(int32_t), 256);
  int32_t* counts = reinterpret_cast<int32_t*>(p);     p += align_up((size_t)n_blocks * 256 * sizeof(int32_t), 256);
  void* scan_ws = p;
  const size_t scan_ws_bytes = scan_workspace_bytes((int64_t)n_blocks * 256);
  int passes = (key_bits + 7) / 8;
  if (passes & 1) ++passes;  // even number of passes so the result lands in (keys, vals)
  uint64_t* kin = keys; int32_t* vin = vals; uint64_t* kout = keys_alt; int32_t* vout = vals_alt;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * 8;
    radix_hist_kernel<<<n_blocks, kSortThreads, 0, stream>>>(kin, n, shift, counts, n_blocks);
    int32_t rc = exclusive_scan_i32(counts, counts, (int64_t)n_blocks * 256, nullptr, scan_ws, scan_ws_bytes, stream);
    if (rc != GCD_OK) return rc;
    radix_scatter_kernel<<<n_blocks, kSortThreads, 0, stream>>>(kin, vin, n, shift, counts, n_blocks, kout, vout);
    uint64_t* tk = kin; kin = kout; kout = tk;
    int32_t* tv = vin; vin = vout; vout = tv;
  }
  GCD_LAUNCH_CHECK("radix_sort_pairs");
  return GCD_OK;
}
}  // namespace gcd
