// coords.cu — point->voxel quantisation, voxel dedup, coordinate maps and kernel maps.
// All integer work; HBM/L2-bound.  Hash: open addressing, linear probing, 64-bit packed keys.
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"

namespace gcd {
size_t radix_sort_workspace_bytes(int64_t n);
int32_t radix_sort_pairs(uint64_t* keys, int32_t* vals, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);

namespace {
constexpr int kThreads = 256;
inline unsigned grid_for(int64_t n, int per_block = kThreads) { return (unsigned)ceil_div(n > 0 ? n : 1, per_block); }

// ------------------------------------------------------------------------------ quantise
template <typename T>
__global__ void __launch_bounds__(kThreads) quantize_kernel(const T* __restrict__ pts, int64_t ld, int64_t n, int dims, T q,
                                                             int round_mode, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dims) return;
  const int64_t i = t / dims;
  const int d = (int)(t - i * dims);
  T v = pts[i * ld + d] / q;  // IEEE division (no -use_fast_math, no reciprocal)
  if (sizeof(T) == 4) v = round_mode == GCD_ROUND_FLOOR ? (T)floorf((float)v) : (T)rintf((float)v);
  else                v = round_mode == GCD_ROUND_FLOOR ? (T)floor((double)v) : (T)rint((double)v);
  out[t] = (int32_t)v;
}

__global__ void __launch_bounds__(kThreads) colmin_kernel(const int32_t* __restrict__ c, int64_t n, int dims, int32_t* mins) {
  int m[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int d = 0; d < dims; ++d) m[d] = min(m[d], c[i * dims + d]);
  for (int d = 0; d < dims; ++d) {
    int v = m[d];
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v != INT_MAX) atomicMin(&mins[d], v);
  }
}
__global__ void __launch_bounds__(kThreads) sub_cols_kernel(int32_t* c, int64_t n, int dims, const int32_t* __restrict__ mins) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * dims) c[t] -= mins[t % dims];
}

// ------------------------------------------------------------------------------ hash
__global__ void __launch_bounds__(kThreads) table_clear_kernel(uint64_t* keys, int32_t* vals, int64_t cap, int32_t val_init) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cap) { keys[t] = kEmptyKey; vals[t] = val_init; }
}

__device__ __forceinline__ void load_coord(const int32_t* __restrict__ coords, int64_t i, int dims, int& b, int& x, int& y, int& z) {
  if (dims == 4) {
    int4 c = *reinterpret_cast<const int4*>(coords + i * 4);
    b = c.x; x = c.y; y = c.z; z = c.w;
  } else {
    b = 0; x = coords[i * 3]; y = coords[i * 3 + 1]; z = coords[i * 3 + 2];
  }
}

// Insert every row; vals[slot] = min row index with that key.  slot_of[i] remembers the slot.
__global__ void __launch_bounds__(kThreads) insert_min_kernel(const int32_t* __restrict__ coords, int64_t n, int dims,
                                                               uint64_t* keys, int32_t* vals, int64_t cap,
                                                               int32_t* __restrict__ slot_of, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b, x, y, z;
  load_coord(coords, i, dims, b, x, y, z);
  if (!key_in_range(b, x, y, z)) { atomicOr(status, GCD_DEV_KEY_RANGE); slot_of[i] = -1; return; }
  bool present;
  int64_t slot = table_insert(keys, cap, pack_key(b, x, y, z), &present);
  if (slot < 0) { atomicOr(status, GCD_DEV_TABLE_FULL); slot_of[i] = -1; return; }
  atomicMin(&vals[slot], (int32_t)i);
  slot_of[i] = (int32_t)slot;
}

__global__ void __launch_bounds__(kThreads) flag_first_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ vals,
                                                               int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (slot_of[i] >= 0 && vals[slot_of[i]] == (int32_t)i) ? 1 : 0;
}

// rank[] = exclusive scan of flags.  first-occurrence numbering: voxel of point i = rank[first(i)].
__global__ void __launch_bounds__(kThreads) unique_emit_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ vals,
                                                                const int32_t* __restrict__ flags, const int32_t* __restrict__ rank,
                                                                int64_t n, int64_t* __restrict__ unique_idx, int64_t* __restrict__ inverse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = slot_of[i];
  if (s < 0) { inverse[i] = -1; return; }
  inverse[i] = rank[vals[s]];
  if (flags[i]) unique_idx[rank[i]] = i;
}
// After emit: table value becomes the voxel index instead of the first point index.
__global__ void __launch_bounds__(kThreads) table_set_rank_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ flags,
                                                                   const int32_t* __restrict__ rank, int64_t n, int32_t* vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) vals[slot_of[i]] = rank[i];
}

// Sorted-order flavour helpers (the voxel count m lives on the device).
__global__ void __launch_bounds__(kThreads) gather_voxel_keys_kernel(const int64_t* __restrict__ unique_fo, const int32_t* __restrict__ slot_of,
                                                                      const uint64_t* __restrict__ keys, const int32_t* __restrict__ m_dev,
                                                                      int64_t n, uint64_t* __restrict__ vkeys, int32_t* __restrict__ vids) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (j < *m_dev) { vkeys[j] = keys[slot_of[unique_fo[j]]]; vids[j] = (int32_t)j; }
  else            { vkeys[j] = kEmptyKey; vids[j] = -1; }  // padding sorts to the end
}
__global__ void __launch_bounds__(kThreads) sorted_rank_kernel(const int32_t* __restrict__ perm, const int64_t* __restrict__ unique_fo,
                                                                const int32_t* __restrict__ m_dev, int64_t n,
                                                                int32_t* __restrict__ new_rank, int64_t* __restrict__ unique_sorted) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && j < *m_dev) { new_rank[perm[j]] = (int32_t)j; unique_sorted[j] = unique_fo[perm[j]]; }
}
__global__ void __launch_bounds__(kThreads) remap_inverse_kernel(int64_t* inverse, int64_t n, const int32_t* __restrict__ new_rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && inverse[i] >= 0) inverse[i] = new_rank[inverse[i]];
}
__global__ void __launch_bounds__(kThreads) remap_table_kernel(const uint64_t* __restrict__ keys, int32_t* vals, int64_t cap,
                                                                const int32_t* __restrict__ new_rank) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cap && keys[t] != kEmptyKey) vals[t] = new_rank[vals[t]];
}

__global__ void __launch_bounds__(kThreads) hash_build_kernel(const int32_t* __restrict__ coords, int64_t n, uint64_t* keys,
                                                               int32_t* vals, int64_t cap, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + i * 4);
  if (!key_in_range(c.x, c.y, c.z, c.w)) { atomicOr(status, GCD_DEV_KEY_RANGE); return; }
  bool present;
  int64_t slot = table_insert(keys, cap, pack_key(c.x, c.y, c.z, c.w), &present);
  if (slot < 0) { atomicOr(status, GCD_DEV_TABLE_FULL); return; }
  if (present) atomicOr(status, GCD_DEV_DUPLICATE);
  vals[slot] = (int32_t)i;
}

// ------------------------------------------------------------------------------ stride 2
__device__ __forceinline__ int floor_div_pos(int a, int s) {  // floor(a / s), s > 0
  int q = a / s;
  return (a % s != 0 && a < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(kThreads) stride2_insert_kernel(const int32_t* __restrict__ coords, int64_t n, int ts,
                                                                   uint64_t* keys, int32_t* vals, int64_t cap,
                                                                   int32_t* __restrict__ slot_of, int32_t* __restrict__ code,
                                                                   int32_t* status) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + f * 4);
  const int s = 2 * ts;
  const int px = floor_div_pos(c.y, s) * s, py = floor_div_pos(c.z, s) * s, pz = floor_div_pos(c.w, s) * s;
  code[f] = (c.y - px) / ts + 2 * ((c.z - py) / ts) + 4 * ((c.w - pz) / ts);
  if (!key_in_range(c.x, px, py, pz)) { atomicOr(status, GCD_DEV_KEY_RANGE); slot_of[f] = -1; return; }
  bool present;
  int64_t slot = table_insert(keys, cap, pack_key(c.x, px, py, pz), &present);
  if (slot < 0) { atomicOr(status, GCD_DEV_TABLE_FULL); slot_of[f] = -1; return; }
  atomicMin(&vals[slot], (int32_t)f);
  slot_of[f] = (int32_t)slot;
}

__global__ void __launch_bounds__(kThreads) stride2_emit_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ vals,
                                                                 const uint64_t* __restrict__ keys, const int32_t* __restrict__ flags,
                                                                 const int32_t* __restrict__ rank, int64_t n,
                                                                 int32_t* __restrict__ coarse, int32_t* __restrict__ parent) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int32_t s = slot_of[f];
  if (s < 0) { parent[f] = -1; return; }
  parent[f] = rank[vals[s]];
  if (flags[f]) {
    int b, x, y, z;
    unpack_key(keys[s], b, x, y, z);
    *reinterpret_cast<int4*>(coarse + (int64_t)rank[f] * 4) = make_int4(b, x, y, z);
  }
}

// ------------------------------------------------------------------------------ kernel maps
// One thread per output voxel, all K^3 offsets probed back to back (independent loads in flight);
// stores are column-major so a warp writes 128 contiguous bytes per offset.
template <int K>
__global__ void __launch_bounds__(kThreads) kmap_subm_kernel(const int32_t* __restrict__ coords, int64_t n,
                                                              const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                                              int64_t cap, int ts, int32_t* __restrict__ nbr) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + o * 4);
  constexpr int R = K / 2;
#pragma unroll 1
  for (int kz = 0; kz < K; ++kz)
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
      int res[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int x = c.y + (kx - R) * ts, y = c.z + (ky - R) * ts, z = c.w + (kz - R) * ts;
        int r = -1;
        if (kx == R && ky == R && kz == R) r = (int)o;  // centre tap is the voxel itself
        else if (key_in_range(c.x, x, y, z)) {
          int64_t slot = table_find(keys, cap, pack_key(c.x, x, y, z));
          if (slot >= 0) r = __ldg(&vals[slot]);
        }
        res[kx] = r;
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) nbr[(int64_t)((kz * K + ky) * K + kx) * n + o] = res[kx];
    }
}

// Warp-cooperative probing (the design BASELINE.json's north_star names): a cooperative group of four lanes owns one output
// voxel and probes the table one 32-byte SECTOR (four adjacent slots) at a time -- one coalesced load per probe, a ballot
// inside the group for "found" / "chain ends here" -- instead of one thread walking its chain slot by slot.  Same table, same
// result bit for bit as kmap_subm_kernel; selected by GCD_OPT_KMAP_COOP (measured against the other two searches in
// profiles/).
template <int K>
__global__ void __launch_bounds__(kThreads) kmap_subm_coop_kernel(const int32_t* __restrict__ coords, int64_t n,
                                                                   const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                                                   int64_t cap, int ts, int32_t* __restrict__ nbr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t o = t >> 2;                         // four lanes per voxel
  const int sub = threadIdx.x & 3;
  const unsigned gmask = 0xFu << (threadIdx.x & 28);  // this group's lanes within the warp
  if (o >= n) return;                               // n * 4 threads are launched in whole groups: a group leaves together
  const int4 c = *reinterpret_cast<const int4*>(coords + o * 4);
  constexpr int R = K / 2;
#pragma unroll 1
  for (int k = 0; k < K * K * K; ++k) {
    const int kx = k % K, ky = (k / K) % K, kz = k / (K * K);
    const int x = c.y + (kx - R) * ts, y = c.z + (ky - R) * ts, z = c.w + (kz - R) * ts;
    int r = -1;
    if (kx == R && ky == R && kz == R) r = (int)o;
    else if (key_in_range(c.x, x, y, z)) {
      const uint64_t key = pack_key(c.x, x, y, z);
      const int64_t home = (int64_t)(hash_key(key) & (uint64_t)(cap - 1));
      int64_t sector = home & ~int64_t(3);
      unsigned live = 0xFu << (home & 3) & 0xFu;      // first sector: slots before the home slot are not on this key's chain
      for (int64_t probes = 0; probes < cap; probes += 4) {
        const int64_t slot = sector + sub;
        const uint64_t cur = __ldg(&keys[slot]);
        const unsigned hit = (__ballot_sync(gmask, cur == key) >> (threadIdx.x & 28)) & 0xFu;
        const unsigned end = (__ballot_sync(gmask, cur == kEmptyKey) >> (threadIdx.x & 28)) & live;
        if (hit) {
          // a chain never skips an empty slot: a hit behind an empty slot of the live part would be a different chain's key,
          // which cannot equal this key (keys are unique in the table) -- so any hit is the answer
          const int src = (threadIdx.x & 28) + (__ffs(hit) - 1);
          const int v = (cur == key) ? __ldg(&vals[slot]) : 0;
          r = __shfl_sync(gmask, v, src);
          break;
        }
        if (end) break;
        sector = (sector + 4) & (cap - 1);
        live = 0xFu;
      }
    }
    if (sub == 0) nbr[(int64_t)k * n + o] = r;
  }
}

__global__ void __launch_bounds__(kThreads) fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}
__global__ void __launch_bounds__(kThreads) kmap_down2_kernel(const int32_t* __restrict__ parent, const int32_t* __restrict__ code,
                                                               int64_t n_fine, int64_t n_coarse, int32_t* __restrict__ nbr) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n_fine && parent[f] >= 0) nbr[(int64_t)code[f] * n_coarse + parent[f]] = (int32_t)f;
}
__global__ void __launch_bounds__(kThreads) kmap_up2_kernel(const int32_t* __restrict__ parent, const int32_t* __restrict__ code,
                                                             int64_t n_fine, int32_t* __restrict__ nbr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_fine * 8) return;
  const int64_t k = t / n_fine, f = t - k * n_fine;
  nbr[t] = (code[f] == (int)k) ? parent[f] : -1;
}

// ------------------------------------------------------------------------------ pair lists
__global__ void __launch_bounds__(kThreads) flag_valid_kernel(const int32_t* __restrict__ nbr, int64_t total, int32_t* __restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) flags[t] = nbr[t] >= 0;
}
__global__ void __launch_bounds__(kThreads) pairs_emit_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ pos,
                                                               const int32_t* __restrict__ total_pairs, int64_t n_out, int kv,
                                                               int32_t* __restrict__ pair_in, int32_t* __restrict__ pair_out,
                                                               int32_t* __restrict__ pair_off) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_out * kv;
  if (t >= total) return;
  const int64_t k = t / n_out, o = t - k * n_out;
  const int v = nbr[t];
  if (v >= 0) { pair_in[pos[t]] = v; pair_out[pos[t]] = (int32_t)o; }
  if (o == 0) pair_off[k] = pos[t];
  if (t == total - 1) pair_off[kv] = *total_pairs;
}
}  // namespace
}  // namespace gcd

using namespace gcd;

// =================================================================================== C ABI
extern "C" int32_t gcd_quantize_f32(const float* pts, int64_t ld, int64_t n, int32_t dims, float q, int32_t round_mode,
                                    int32_t* out, void* stream) {
  GCD_REQUIRE(n >= 0 && (dims == 3 || dims == 4) && ld >= dims, "gcd_quantize_f32: bad shape n=%lld dims=%d ld=%lld", (long long)n, dims, (long long)ld);
  GCD_REQUIRE(q > 0.f, "gcd_quantize_f32: quantisation size must be positive");
  if (n == 0) return GCD_OK;
  quantize_kernel<float><<<grid_for(n * dims), kThreads, 0, as_stream(stream)>>>(pts, ld, n, dims, q, round_mode, out);
  GCD_LAUNCH_CHECK("gcd_quantize_f32");
  return GCD_OK;
}
extern "C" int32_t gcd_quantize_f64(const double* pts, int64_t ld, int64_t n, int32_t dims, double q, int32_t round_mode,
                                    int32_t* out, void* stream) {
  GCD_REQUIRE(n >= 0 && (dims == 3 || dims == 4) && ld >= dims, "gcd_quantize_f64: bad shape");
  GCD_REQUIRE(q > 0.0, "gcd_quantize_f64: quantisation size must be positive");
  if (n == 0) return GCD_OK;
  quantize_kernel<double><<<grid_for(n * dims), kThreads, 0, as_stream(stream)>>>(pts, ld, n, dims, q, round_mode, out);
  GCD_LAUNCH_CHECK("gcd_quantize_f64");
  return GCD_OK;
}
extern "C" int32_t gcd_colmin_i32(const int32_t* coords, int64_t n, int32_t dims, int32_t* mins, void* stream) {
  GCD_REQUIRE(dims >= 1 && dims <= 4, "gcd_colmin_i32: dims must be 1..4");
  if (n == 0) return GCD_OK;
  unsigned g = (unsigned)std::min<int64_t>(ceil_div(n, kThreads), kNumSMs * 8);
  colmin_kernel<<<g, kThreads, 0, as_stream(stream)>>>(coords, n, dims, mins);
  GCD_LAUNCH_CHECK("gcd_colmin_i32");
  return GCD_OK;
}
extern "C" int32_t gcd_sub_cols_i32(int32_t* coords, int64_t n, int32_t dims, const int32_t* mins, void* stream) {
  if (n == 0) return GCD_OK;
  sub_cols_kernel<<<grid_for(n * dims), kThreads, 0, as_stream(stream)>>>(coords, n, dims, mins);
  GCD_LAUNCH_CHECK("gcd_sub_cols_i32");
  return GCD_OK;
}

extern "C" int64_t gcd_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

namespace {
struct UniqueWs { int32_t *slot_of, *flags, *rank, *new_rank, *vids; uint64_t* vkeys; int64_t* unique_tmp; void* scan; void* sort; size_t scan_bytes, sort_bytes; };
size_t carve_unique(int64_t n, char* base, UniqueWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
  int64_t nn = n > 0 ? n : 1;
  char* a = take(nn * 4); char* b = take(nn * 4); char* c = take(nn * 4); char* d = take(nn * 4); char* e = take(nn * 4);
  char* f = take(nn * 8); char* g = take(nn * 8);
  size_t sb = scan_workspace_bytes(nn), rb = radix_sort_workspace_bytes(nn);
  char* h = take(sb); char* i = take(rb);
  if (w) { w->slot_of = (int32_t*)a; w->flags = (int32_t*)b; w->rank = (int32_t*)c; w->new_rank = (int32_t*)d; w->vids = (int32_t*)e;
           w->vkeys = (uint64_t*)f; w->unique_tmp = (int64_t*)g; w->scan = h; w->sort = i; w->scan_bytes = sb; w->sort_bytes = rb; }
  return off;
}
}  // namespace

extern "C" size_t gcd_unique_workspace_bytes(int64_t n) { return carve_unique(n, nullptr, nullptr); }

extern "C" int32_t gcd_unique_rows(const int32_t* coords, int64_t n, int32_t dims, int32_t order, uint64_t* table_keys,
                                   int32_t* table_vals, int64_t cap, int64_t* unique_idx, int64_t* inverse, int32_t* m_out,
                                   void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  GCD_REQUIRE(n >= 0 && n < (1ll << 30), "gcd_unique_rows: n out of range");
  GCD_REQUIRE(dims == 3 || dims == 4, "gcd_unique_rows: dims must be 3 or 4");
  GCD_REQUIRE(order == 0 || order == 1, "gcd_unique_rows: order must be 0 or 1");
  GCD_REQUIRE(cap >= 2 * n && (cap & (cap - 1)) == 0, "gcd_unique_rows: capacity must be a power of two >= 2n");
  if (workspace_bytes < gcd_unique_workspace_bytes(n)) { set_error("gcd_unique_rows: workspace too small"); return GCD_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  UniqueWs w;
  carve_unique(n, static_cast<char*>(workspace), &w);
  table_clear_kernel<<<grid_for(cap), kThreads, 0, st>>>(table_keys, table_vals, cap, INT_MAX);
  if (n == 0) { cudaMemsetAsync(m_out, 0, sizeof(int32_t), st); GCD_LAUNCH_CHECK("gcd_unique_rows"); return GCD_OK; }
  insert_min_kernel<<<grid_for(n), kThreads, 0, st>>>(coords, n, dims, table_keys, table_vals, cap, w.slot_of, status);
  flag_first_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, table_vals, n, w.flags);
  int32_t rc = exclusive_scan_i32(w.flags, w.rank, n, m_out, w.scan, w.scan_bytes, st);
  if (rc != GCD_OK) return rc;
  int64_t* unique_fo = order == 0 ? unique_idx : w.unique_tmp;
  unique_emit_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, table_vals, w.flags, w.rank, n, unique_fo, inverse);
  table_set_rank_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, w.flags, w.rank, n, table_vals);
  if (order == 1) {
    // M is only known on the device: sort n padded entries, the padding carries the largest key.
    gather_voxel_keys_kernel<<<grid_for(n), kThreads, 0, st>>>(unique_fo, w.slot_of, table_keys, m_out, n, w.vkeys, w.vids);
    rc = radix_sort_pairs(w.vkeys, w.vids, n, 64, w.sort, w.sort_bytes, st);
    if (rc != GCD_OK) return rc;
    sorted_rank_kernel<<<grid_for(n), kThreads, 0, st>>>(w.vids, unique_fo, m_out, n, w.new_rank, unique_idx);
    remap_inverse_kernel<<<grid_for(n), kThreads, 0, st>>>(inverse, n, w.new_rank);
    remap_table_kernel<<<grid_for(cap), kThreads, 0, st>>>(table_keys, table_vals, cap, w.new_rank);
  }
  GCD_LAUNCH_CHECK("gcd_unique_rows");
  return GCD_OK;
}

extern "C" int32_t gcd_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals, int64_t cap,
                                  int32_t* status, void* stream) {
  GCD_REQUIRE(n >= 0 && n < (1ll << 30), "gcd_hash_build: n out of range");
  GCD_REQUIRE(cap >= 2 * n && (cap & (cap - 1)) == 0, "gcd_hash_build: capacity must be a power of two >= 2n");
  cudaStream_t st = as_stream(stream);
  table_clear_kernel<<<grid_for(cap), kThreads, 0, st>>>(table_keys, table_vals, cap, -1);
  if (n > 0) hash_build_kernel<<<grid_for(n), kThreads, 0, st>>>(coords, n, table_keys, table_vals, cap, status);
  GCD_LAUNCH_CHECK("gcd_hash_build");
  return GCD_OK;
}

namespace {
struct Stride2Ws { int32_t *slot_of, *flags, *rank; void* scan; size_t scan_bytes; };
size_t carve_stride2(int64_t n, char* base, Stride2Ws* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
  int64_t nn = n > 0 ? n : 1;
  char* a = take(nn * 4); char* b = take(nn * 4); char* c = take(nn * 4);
  size_t sb = scan_workspace_bytes(nn);
  char* d = take(sb);
  if (w) { w->slot_of = (int32_t*)a; w->flags = (int32_t*)b; w->rank = (int32_t*)c; w->scan = d; w->scan_bytes = sb; }
  return off;
}
}  // namespace

extern "C" size_t gcd_stride2_workspace_bytes(int64_t n) { return carve_stride2(n, nullptr, nullptr); }

extern "C" int32_t gcd_coords_stride2(const int32_t* coords, int64_t n, int32_t ts, uint64_t* coarse_keys, int32_t* coarse_vals,
                                      int64_t cap_coarse, int32_t* coarse_coords, int32_t* parent, int32_t* code, int32_t* m_out,
                                      void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  GCD_REQUIRE(n >= 0 && n < (1ll << 30), "gcd_coords_stride2: n out of range");
  GCD_REQUIRE(ts >= 1 && ts <= (1 << 14), "gcd_coords_stride2: tensor stride out of range");
  GCD_REQUIRE(cap_coarse >= 2 * n && (cap_coarse & (cap_coarse - 1)) == 0, "gcd_coords_stride2: capacity must be a power of two >= 2n");
  if (workspace_bytes < gcd_stride2_workspace_bytes(n)) { set_error("gcd_coords_stride2: workspace too small"); return GCD_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  Stride2Ws w;
  carve_stride2(n, static_cast<char*>(workspace), &w);
  table_clear_kernel<<<grid_for(cap_coarse), kThreads, 0, st>>>(coarse_keys, coarse_vals, cap_coarse, INT_MAX);
  if (n == 0) { cudaMemsetAsync(m_out, 0, sizeof(int32_t), st); GCD_LAUNCH_CHECK("gcd_coords_stride2"); return GCD_OK; }
  stride2_insert_kernel<<<grid_for(n), kThreads, 0, st>>>(coords, n, ts, coarse_keys, coarse_vals, cap_coarse, w.slot_of, code, status);
  flag_first_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, coarse_vals, n, w.flags);
  int32_t rc = exclusive_scan_i32(w.flags, w.rank, n, m_out, w.scan, w.scan_bytes, st);
  if (rc != GCD_OK) return rc;
  stride2_emit_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, coarse_vals, coarse_keys, w.flags, w.rank, n, coarse_coords, parent);
  table_set_rank_kernel<<<grid_for(n), kThreads, 0, st>>>(w.slot_of, w.flags, w.rank, n, coarse_vals);
  GCD_LAUNCH_CHECK("gcd_coords_stride2");
  return GCD_OK;
}

extern "C" int32_t gcd_kmap_subm(const int32_t* coords, int64_t n, const uint64_t* table_keys, const int32_t* table_vals,
                                 int64_t cap, int32_t kernel_size, int32_t ts, int32_t* nbr, void* stream) {
  GCD_REQUIRE(kernel_size == 3 || kernel_size == 5, "gcd_kmap_subm: kernel_size must be 3 or 5 (got %d)", kernel_size);
  GCD_REQUIRE(n >= 0 && ts >= 1, "gcd_kmap_subm: bad arguments");
  if (n == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  if (option(GCD_OPT_KMAP_COOP)) {        // warp-cooperative probing: four lanes per voxel, one sector per probe
    if (kernel_size == 3) kmap_subm_coop_kernel<3><<<grid_for(n * 4), kThreads, 0, st>>>(coords, n, table_keys, table_vals, cap, ts, nbr);
    else                  kmap_subm_coop_kernel<5><<<grid_for(n * 4), kThreads, 0, st>>>(coords, n, table_keys, table_vals, cap, ts, nbr);
  } else if (kernel_size == 3) kmap_subm_kernel<3><<<grid_for(n), kThreads, 0, st>>>(coords, n, table_keys, table_vals, cap, ts, nbr);
  else                         kmap_subm_kernel<5><<<grid_for(n), kThreads, 0, st>>>(coords, n, table_keys, table_vals, cap, ts, nbr);
  GCD_LAUNCH_CHECK("gcd_kmap_subm");
  return GCD_OK;
}

namespace gcd { namespace {
struct Affine { double m[12]; };
__global__ void __launch_bounds__(256) affine_f64_kernel(const float* __restrict__ pts, int64_t ld, int64_t n, const Affine a, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = (double)pts[i * ld], y = (double)pts[i * ld + 1], z = (double)pts[i * ld + 2];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    out[i * 3 + j] = fma(z, a.m[4 * j + 2], fma(y, a.m[4 * j + 1], x * a.m[4 * j])) + a.m[4 * j + 3];
}
} }

extern "C" int32_t gcd_affine_f64(const float* pts, int64_t ld, int64_t n, const double* m_host, double* out, void* stream) {
  GCD_REQUIRE(n >= 0 && ld >= 3 && m_host != nullptr, "gcd_affine_f64: bad arguments");
  if (n == 0) return GCD_OK;
  GCD_REQUIRE(pts && out, "gcd_affine_f64: null pointer");
  Affine a;
  for (int i = 0; i < 12; ++i) a.m[i] = m_host[i];
  affine_f64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(pts, ld, n, a, out);
  GCD_LAUNCH_CHECK("gcd_affine_f64");
  return GCD_OK;
}

extern "C" int32_t gcd_kmap_down2(const int32_t* parent, const int32_t* code, int64_t n_fine, int64_t n_coarse, int32_t* nbr,
                                  void* stream) {
  cudaStream_t st = as_stream(stream);
  if (n_coarse > 0) fill_i32_kernel<<<grid_for(n_coarse * 8), kThreads, 0, st>>>(nbr, n_coarse * 8, -1);
  if (n_fine > 0) kmap_down2_kernel<<<grid_for(n_fine), kThreads, 0, st>>>(parent, code, n_fine, n_coarse, nbr);
  GCD_LAUNCH_CHECK("gcd_kmap_down2");
  return GCD_OK;
}
extern "C" int32_t gcd_kmap_up2(const int32_t* parent, const int32_t* code, int64_t n_fine, int32_t* nbr, void* stream) {
  if (n_fine > 0) kmap_up2_kernel<<<grid_for(n_fine * 8), kThreads, 0, as_stream(stream)>>>(parent, code, n_fine, nbr);
  GCD_LAUNCH_CHECK("gcd_kmap_up2");
  return GCD_OK;
}

extern "C" size_t gcd_pairs_workspace_bytes(int64_t n_out, int32_t kv) {
  // sized for the form GCD_OPT_PAIRS_FUSED selects at the time of the query (the call checks against the form it runs)
  int64_t total = (n_out > 0 ? n_out : 1) * (int64_t)kv;
  const int form = option(GCD_OPT_PAIRS_FUSED);
  if (form == 2) return pairs_lookback_workspace_bytes(total);
  if (form == 1) return align_up(4, 256) + scan_workspace_bytes(total);
  return align_up((size_t)total * 4, 256) * 2 + align_up(4, 256) + scan_workspace_bytes(total);
}
extern "C" int32_t gcd_pairs_from_table(const int32_t* nbr, int64_t n_out, int32_t kv, int32_t* pair_in, int32_t* pair_out,
                                        int32_t* pair_off, void* workspace, size_t workspace_bytes, void* stream) {
  GCD_REQUIRE(kv >= 1 && n_out >= 0 && n_out * (int64_t)kv < (1ll << 31), "gcd_pairs_from_table: table too large for int32 pair offsets");
  if (workspace_bytes < gcd_pairs_workspace_bytes(n_out, kv)) { set_error("gcd_pairs_from_table: workspace too small"); return GCD_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  if (n_out == 0) { cudaMemsetAsync(pair_off, 0, (size_t)(kv + 1) * 4, st); return GCD_OK; }
  const int64_t total = n_out * kv;
  char* p = static_cast<char*>(workspace);
  const int form = option(GCD_OPT_PAIRS_FUSED);
  if (form == 2) return pairs_from_table_lookback(nbr, n_out, kv, pair_in, pair_out, pair_off, workspace, workspace_bytes, st);
  if (form == 1) {
    int32_t* tot1 = (int32_t*)p;
    return pairs_from_table_fused(nbr, n_out, kv, pair_in, pair_out, pair_off, tot1, p + align_up(4, 256), scan_workspace_bytes(total), st);
  }
  int32_t* flags = (int32_t*)p; p += align_up((size_t)total * 4, 256);
  int32_t* pos = (int32_t*)p;   p += align_up((size_t)total * 4, 256);
  int32_t* tot = (int32_t*)p;   p += align_up(4, 256);
  flag_valid_kernel<<<grid_for(total), kThreads, 0, st>>>(nbr, total, flags);
  int32_t rc = exclusive_scan_i32(flags, pos, total, tot, p, scan_workspace_bytes(total), st);
  if (rc != GCD_OK) return rc;
  pairs_emit_kernel<<<grid_for(total), kThreads, 0, st>>>(nbr, pos, tot, n_out, kv, pair_in, pair_out, pair_off);
  GCD_LAUNCH_CHECK("gcd_pairs_from_table");
  return GCD_OK;
}
