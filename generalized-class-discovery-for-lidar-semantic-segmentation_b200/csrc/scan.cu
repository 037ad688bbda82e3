// scan.cu — device-wide exclusive scan of int32 (three small kernels; used by the compactions
// of the quantiser, the coarse-map builder and the pair-list builder).
#include "common.cuh"

namespace gcd {
namespace {
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 1024

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, writes block total.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[kScanThreads / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    int si = warp_inclusive_scan(s);
    if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) block_total = si;
  }
  __syncthreads();
  int r = inc - v + warp_sums[warp];
  *total = block_total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                                   int64_t n, int32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    sum += v[i];
  }
  int total;
  int prefix = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = prefix;
    prefix += v[i];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single block: exclusive scan of the tile sums in place, total written out.
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int32_t* __restrict__ sums, int64_t n_tiles,
                                                                  int32_t* __restrict__ total_out) {
  int carry = 0;
  for (int64_t base = 0; base < n_tiles; base += kScanThreads) {
    int64_t i = base + threadIdx.x;
    int v = i < n_tiles ? sums[i] : 0;
    int total;
    int p = block_exclusive_scan(v, &total);
    if (i < n_tiles) sums[i] = p + carry;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int32_t* __restrict__ out, int64_t n,
                                                                 const int32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  const int add = tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += add;
}

// ---- fused pair-list build (opt-in, GCD_PAIRS_FUSED=1): the flattened table [kv * n_out] is read twice and nothing but
// the pair lists is written (the three-pass version writes and re-reads a flag array and a position array of the table's
// size: 8 passes over 4 * kv * n_out bytes instead of 2).  Same tiling and the same block scan as scan_tiles_kernel, so the
// positions -- and therefore the lists -- are identical.
__global__ void __launch_bounds__(kScanThreads) pairs_count_kernel(const int32_t* __restrict__ nbr, int64_t total_entries,
                                                                    int32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) sum += (base + i < total_entries && nbr[base + i] >= 0) ? 1 : 0;
  int total;
  block_exclusive_scan(sum, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// tile_sums holds the exclusive scan of the tile counts; total_pairs the grand total.
__global__ void __launch_bounds__(kScanThreads) pairs_emit_fused_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int kv,
                                                                         const int32_t* __restrict__ tile_sums,
                                                                         const int32_t* __restrict__ total_pairs,
                                                                         int32_t* __restrict__ pair_in, int32_t* __restrict__ pair_out,
                                                                         int32_t* __restrict__ pair_off) {
  const int64_t total_entries = n_out * kv;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < total_entries) ? nbr[base + i] : -1;
    sum += v[i] >= 0 ? 1 : 0;
  }
  int total;
  int pos = block_exclusive_scan(sum, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t t = base + i;
    if (t < total_entries) {
      const int64_t k = t / n_out, o = t - k * n_out;
      if (o == 0) pair_off[k] = pos;                       // first entry of offset k: where its list starts
      if (v[i] >= 0) { pair_in[pos] = v[i]; pair_out[pos] = (int32_t)o; ++pos; }
      if (t == total_entries - 1) pair_off[kv] = *total_pairs;
    }
  }
}
// ---- single-pass pair-list build (GCD_PAIRS_FUSED=2): decoupled look-back.  One launch reads the table ONCE: a block takes a
// tile of 4096 entries by ticket (tiles are started in order, so every predecessor of a running tile is running or done),
// counts its valid entries, publishes the count, sums its predecessors' published counts backwards until it meets one that
// already knows its inclusive prefix, publishes its own, and writes its pairs.  state[tile] = status << 32 | value with status
// 0 = nothing yet, 1 = the tile's own count, 2 = inclusive prefix; state and ticket are zero on entry (one memset).
constexpr int kLbItems = 16;
constexpr int kLbTile = kScanThreads * kLbItems;   // 4096
constexpr unsigned long long kLbAggregate = 1ull << 32, kLbInclusive = 2ull << 32;

__global__ void __launch_bounds__(kScanThreads) pairs_lookback_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int kv,
                                                                       unsigned long long* state, int32_t* ticket,
                                                                       int32_t* __restrict__ pair_in, int32_t* __restrict__ pair_out,
                                                                       int32_t* __restrict__ pair_off) {
  __shared__ int s_tile, s_prefix;
  const int64_t total_entries = n_out * kv;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * kLbTile + (int64_t)threadIdx.x * kLbItems;
  int v[kLbItems];
  int sum = 0;
  if (base + kLbItems <= total_entries && (reinterpret_cast<uintptr_t>(nbr) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < kLbItems; i += 4) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(nbr + base + i));
      v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kLbItems; ++i) v[i] = (base + i < total_entries) ? __ldg(&nbr[base + i]) : -1;
  }
#pragma unroll
  for (int i = 0; i < kLbItems; ++i) sum += v[i] >= 0 ? 1 : 0;
  int total;
  int pos = block_exclusive_scan(sum, &total);
  if (threadIdx.x == 0) {
    // tile 0 knows its inclusive prefix at once
    atomicExch(&state[tile], (tile == 0 ? kLbInclusive : kLbAggregate) | (unsigned)total);
    if (tile == 0) s_prefix = 0;
  }
  if (tile > 0 && threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int prefix = 0;
    for (int look = tile - 1;; look -= 32) {
      const int t = look - lane;                       // lane 0 looks at the nearest predecessor
      unsigned long long st = kLbInclusive;            // tiles before the first: an inclusive prefix of 0
      if (t >= 0) {
        do { st = *reinterpret_cast<volatile unsigned long long*>(&state[t]); } while ((st >> 32) == 0);
      }
      const unsigned incl = __ballot_sync(0xffffffffu, (st >> 32) == 2);
      const int first = incl ? __ffs(incl) - 1 : 32;   // nearest predecessor with an inclusive prefix
      int val = lane <= first ? (int)(unsigned)st : 0; // aggregates of the nearer ones + that inclusive prefix
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
      prefix += val;
      if (incl) break;
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(&state[tile], kLbInclusive | (unsigned)(prefix + total));
      s_prefix = prefix;
    }
  }
  __syncthreads();
  pos += s_prefix;
  // entry t = k * n_out + o: the offset k of this thread's first entry by one division, then by counting
  int64_t k = base < total_entries ? base / n_out : 0, o = base - k * n_out;
#pragma unroll
  for (int i = 0; i < kLbItems; ++i) {
    if (base + i < total_entries) {
      if (o == 0) pair_off[k] = pos;                   // first entry of offset k: where its list starts
      if (v[i] >= 0) { pair_in[pos] = v[i]; pair_out[pos] = (int32_t)o; ++pos; }
      if (base + i == total_entries - 1) pair_off[kv] = pos;
      if (++o == n_out) { o = 0; ++k; }
    }
  }
}
}  // namespace

// exclusive scan of int32 in one pass (the same decoupled look-back as pairs_lookback_kernel): out[i] = sum_{j<i} in[j]
__global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(const int32_t* in, int32_t* out, int64_t n,      // in may alias out (sort.cu)
                                                                      unsigned long long* state, int32_t* ticket, int32_t* total_out) {
  __shared__ int s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * kLbTile + (int64_t)threadIdx.x * kLbItems;
  int v[kLbItems];
  int sum = 0;
  if (base + kLbItems <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < kLbItems; i += 4) {
      const int4 q = *reinterpret_cast<const int4*>(in + base + i);       // (in may alias out: plain loads)
      v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kLbItems; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
  }
#pragma unroll
  for (int i = 0; i < kLbItems; ++i) sum += v[i];
  int total;
  int pos = block_exclusive_scan(sum, &total);
  if (threadIdx.x == 0) {
    atomicExch(&state[tile], (tile == 0 ? kLbInclusive : kLbAggregate) | (unsigned)total);
    if (tile == 0) s_prefix = 0;
  }
  if (tile > 0 && threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int prefix = 0;
    for (int look = tile - 1;; look -= 32) {
      const int t = look - lane;
      unsigned long long st = kLbInclusive;
      if (t >= 0) {
        do { st = *reinterpret_cast<volatile unsigned long long*>(&state[t]); } while ((st >> 32) == 0);
      }
      const unsigned incl = __ballot_sync(0xffffffffu, (st >> 32) == 2);
      const int first = incl ? __ffs(incl) - 1 : 32;
      int val = lane <= first ? (int)(unsigned)st : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
      prefix += val;
      if (incl) break;
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(&state[tile], kLbInclusive | (unsigned)(prefix + total));
      s_prefix = prefix;
    }
  }
  __syncthreads();
  pos += s_prefix;
  const bool last_tile = base <= n - 1 && n - 1 < base + kLbItems;     // this thread holds the last element
#pragma unroll
  for (int i = 0; i < kLbItems; ++i) {
    if (base + i < n) out[base + i] = pos;
    pos += v[i];
  }
  if (last_tile && total_out) *total_out = pos;                        // v[i] = 0 beyond n: pos is the grand total
}

size_t pairs_lookback_workspace_bytes(int64_t total_entries) {
  return align_up((size_t)(ceil_div(total_entries > 0 ? total_entries : 1, kLbTile)) * 8 + 8, 256);
}

int32_t pairs_from_table_lookback(const int32_t* nbr, int64_t n_out, int kv, int32_t* pair_in, int32_t* pair_out, int32_t* pair_off,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int64_t total_entries = n_out * kv;
  const size_t need = pairs_lookback_workspace_bytes(total_entries);
  if (workspace_bytes < need) {
    set_error("pairs_from_table_lookback: workspace too small");
    return GCD_ERR_WORKSPACE;
  }
  const int64_t n_tiles = ceil_div(total_entries, kLbTile);
  if (const cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream); e != cudaSuccess) return cuda_fail(e, "gcd_pairs_from_table(memset)");
  unsigned long long* state = static_cast<unsigned long long*>(workspace);
  int32_t* ticket = reinterpret_cast<int32_t*>(state + n_tiles);
  pairs_lookback_kernel<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(nbr, n_out, kv, state, ticket, pair_in, pair_out, pair_off);
  GCD_LAUNCH_CHECK("gcd_pairs_from_table(look-back)");
  return GCD_OK;
}

int32_t pairs_from_table_fused(const int32_t* nbr, int64_t n_out, int kv, int32_t* pair_in, int32_t* pair_out, int32_t* pair_off,
                               int32_t* total_pairs, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int64_t total_entries = n_out * kv;
  if (workspace_bytes < scan_workspace_bytes(total_entries)) {
    set_error("pairs_from_table_fused: workspace too small");
    return GCD_ERR_WORKSPACE;
  }
  int32_t* sums = static_cast<int32_t*>(workspace);
  const int64_t n_tiles = ceil_div(total_entries, kScanTile);
  pairs_count_kernel<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(nbr, total_entries, sums);
  scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(sums, n_tiles, total_pairs);
  pairs_emit_fused_kernel<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(nbr, n_out, kv, sums, total_pairs, pair_in, pair_out, pair_off);
  GCD_LAUNCH_CHECK("gcd_pairs_from_table(fused)");
  return GCD_OK;
}

size_t scan_workspace_bytes(int64_t n) { return align_up((size_t)(ceil_div(n > 0 ? n : 1, kScanTile)) * sizeof(int32_t), 256); }

int32_t exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0) {
    if (total) cudaMemsetAsync(total, 0, sizeof(int32_t), stream);
    return GCD_OK;
  }
  if (workspace_bytes < scan_workspace_bytes(n)) {
    set_error("exclusive_scan_i32: workspace too small");
    return GCD_ERR_WORKSPACE;
  }
  if (option(GCD_OPT_SCAN_LOOKBACK)) {
    // one pass: scan_workspace_bytes(n) (4 bytes per 1024 elements, 256-byte granules) always covers 8 bytes per 4096 + a ticket
    const int64_t lb_tiles = ceil_div(n, kLbTile);
    const size_t need = (size_t)lb_tiles * 8 + 8;
    if (need <= workspace_bytes) {
      if (const cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream); e != cudaSuccess) return cuda_fail(e, "exclusive_scan_i32(memset)");
      unsigned long long* state = static_cast<unsigned long long*>(workspace);
      scan_lookback_kernel<<<(unsigned)lb_tiles, kScanThreads, 0, stream>>>(in, out, n, state, reinterpret_cast<int32_t*>(state + lb_tiles), total);
      GCD_LAUNCH_CHECK("exclusive_scan_i32(look-back)");
      return GCD_OK;
    }
  }
  int32_t* sums = static_cast<int32_t*>(workspace);
  const int64_t n_tiles = ceil_div(n, kScanTile);
  scan_tiles_kernel<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(in, out, n, sums);
  scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(sums, n_tiles, total);
  if (n_tiles > 1) scan_add_kernel<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(out, n, sums);
  GCD_LAUNCH_CHECK("exclusive_scan_i32");
  return GCD_OK;
}
}  // namespace gcd
