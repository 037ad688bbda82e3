// tilesort.cu — tile sort of a 3x3x3 neighbour table for the tcgen05 convolution (tilesort.cuh): presence-mask keys,
// stable radix sort of the columns, permuted table.  Opt-in (GCDLSS_TILE_SORT=1, gcdlss_b200/coords.py).
#include "common.cuh"
#include "tilesort.cuh"

namespace gcd {
size_t radix_sort_workspace_bytes(int64_t n);
int32_t radix_sort_pairs(uint64_t* keys, int32_t* vals, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
namespace {
constexpr int kThreads = 256;
__global__ void __launch_bounds__(kThreads) tile_sort_key_kernel(const int32_t* __restrict__ nbr, int64_t n, unsigned long long* __restrict__ keys,
                                                                  int32_t* __restrict__ vals) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) tile_sort_key_thread(o, nbr, n, keys, vals);
}
__global__ void __launch_bounds__(kThreads) tile_sort_key8_kernel(const int32_t* __restrict__ nbr, int64_t n, unsigned long long* __restrict__ keys,
                                                                   int32_t* __restrict__ vals) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) tile_sort_key8_thread(o, nbr, n, keys, vals);
}
__global__ void __launch_bounds__(kThreads) tile_sort_permute_kernel(const int32_t* __restrict__ nbr, int64_t n, int kv,
                                                                      const int32_t* __restrict__ rows, int32_t* __restrict__ sorted) {
  tile_sort_permute_thread((int64_t)blockIdx.x * blockDim.x + threadIdx.x, nbr, n, kv, rows, sorted);
}
// One warp per 128-column tile of the sorted table: OR of the tile's (sorted) keys -> mask of offsets with a hit.
__global__ void __launch_bounds__(kThreads) tile_masks_kernel(const unsigned long long* __restrict__ keys_sorted, int64_t n, int kv,
                                                               uint32_t* __restrict__ masks) {
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tile * 128 >= n) return;
  uint32_t acc = 0;                                  // keys have at most 27 bits
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t o = tile * 128 + lane + 32 * j;
    if (o < n) acc |= (uint32_t)keys_sorted[o];
  }
  acc = __reduce_or_sync(0xffffffffu, acc);
  if (lane == 0) masks[tile] = tile_mask_from_keys(acc, kv);
}
}  // namespace
}  // namespace gcd

using namespace gcd;

extern "C" size_t gcd_tile_sort_workspace_bytes(int64_t n) {
  const int64_t nn = n > 0 ? n : 1;
  return align_up((size_t)nn * 8, 256) + radix_sort_workspace_bytes(nn);
}

extern "C" int32_t gcd_kmap_tile_sort(const int32_t* nbr, int64_t n, int32_t kv, int32_t* nbr_sorted, int32_t* out_rows,
                                      uint32_t* tile_masks, void* workspace, size_t workspace_bytes, void* stream) {
  GCD_REQUIRE(kv == 27 || kv == 8, "gcd_kmap_tile_sort: only 3x3x3 and 2x2x2 tables (kv = 27, 8) are sorted (got %d)", kv);
  GCD_REQUIRE(n >= 0 && n * (int64_t)kv < (1ll << 31), "gcd_kmap_tile_sort: table too large");
  GCD_REQUIRE(nbr && nbr_sorted && out_rows, "gcd_kmap_tile_sort: null pointer");
  if (workspace_bytes < gcd_tile_sort_workspace_bytes(n)) { set_error("gcd_kmap_tile_sort: workspace too small"); return GCD_ERR_WORKSPACE; }
  if (n == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  char* p = static_cast<char*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(p); p += align_up((size_t)n * 8, 256);
  if (kv == 27) tile_sort_key_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, st>>>(nbr, n, keys, out_rows);
  else          tile_sort_key8_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, st>>>(nbr, n, keys, out_rows);
  const int32_t rc = radix_sort_pairs(reinterpret_cast<uint64_t*>(keys), out_rows, n, kv, p, radix_sort_workspace_bytes(n), st);
  if (rc != GCD_OK) return rc;
  tile_sort_permute_kernel<<<(unsigned)ceil_div(n * kv, kThreads), kThreads, 0, st>>>(nbr, n, kv, out_rows, nbr_sorted);
  if (tile_masks)      // the sorted keys are still in the workspace: one warp per tile ORs 128 of them
    tile_masks_kernel<<<(unsigned)ceil_div(ceil_div(n, 128) * 32, kThreads), kThreads, 0, st>>>(keys, n, kv, tile_masks);
  GCD_LAUNCH_CHECK("gcd_kmap_tile_sort");
  return GCD_OK;
}
