// gather.cu — voxel<->point movement: devoxelisation gather, CSR grouping of points by voxel,
// and segmented sum / mean / max (the gather's backward and the point->voxel encoder reduce).
#include <float.h>
#include <stdlib.h>
#include "common.cuh"
#include "gather_rows.cuh"

namespace gcd {
size_t radix_sort_workspace_bytes(int64_t n);
int32_t radix_sort_pairs(uint64_t* keys, int32_t* vals, int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
namespace {
// One warp per output row; lanes stride over channels in float4 when aligned.
__global__ void __launch_bounds__(256) rows_gather_kernel(const float* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ idx,
                                                           int64_t n_out, int c, float* __restrict__ out, int64_t ld_out, int vec) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_out) return;
  const int64_t src = idx[row];
  if (vec) {
    const float4* s = reinterpret_cast<const float4*>(in + src * ld_in);
    float4* d = reinterpret_cast<float4*>(out + row * ld_out);
    for (int j = lane; j < c / 4; j += 32) d[j] = __ldg(&s[j]);
  } else {
    for (int j = lane; j < c; j += 32) out[row * ld_out + j] = __ldg(&in[src * ld_in + j]);
  }
}

// Vectorised case: float4 elements dealt flat to threads, four independent chains per thread (gather_rows.cuh).
__global__ void __launch_bounds__(kGatherThreads) rows_gather_flat_kernel(const float4* __restrict__ in, int64_t ld_in4,
                                                                          const int64_t* __restrict__ idx, int64_t n_out, int c4,
                                                                          float4* __restrict__ out, int64_t ld_out4) {
  rows_gather_flat_thread<float4>(blockIdx.x, threadIdx.x, in, ld_in4, idx, n_out, c4, out, ld_out4);
}

__global__ void __launch_bounds__(256) csr_keys_kernel(const int64_t* __restrict__ idx, int64_t n, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = (uint64_t)idx[i]; vals[i] = (int32_t)i; }
}
__global__ void __launch_bounds__(256) csr_offsets_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t n_seg, int32_t* __restrict__ seg_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int64_t prev = i == 0 ? -1 : (int64_t)keys[i - 1];
  const int64_t cur = i == n ? n_seg : (int64_t)keys[i];
  for (int64_t s = prev + 1; s <= cur && s <= n_seg; ++s) seg_off[s] = (int32_t)i;
}

__global__ void __launch_bounds__(256) segment_reduce_kernel(const float* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ seg_off,
                                                              const int32_t* __restrict__ order, int64_t n_seg, int c, int mode,
                                                              float* __restrict__ out, int64_t ld_out) {
  const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (seg >= n_seg) return;
  const int b = seg_off[seg], e = seg_off[seg + 1];
  for (int j = lane; j < c; j += 32) {
    float acc = mode == 2 ? -FLT_MAX : 0.f;
    for (int p = b; p < e; ++p) {
      const float v = __ldg(&in[(int64_t)order[p] * ld_in + j]);
      acc = mode == 2 ? fmaxf(acc, v) : acc + v;
    }
    if (e == b) acc = 0.f;
    else if (mode == 1) acc /= (float)(e - b);
    out[seg * ld_out + j] = acc;
  }
}
}  // namespace
}  // namespace gcd

using namespace gcd;

extern "C" int32_t gcd_rows_gather(const float* in, int64_t ld_in, const int64_t* idx, int64_t n_out, int32_t c, float* out,
                                   int64_t ld_out, void* stream) {
  if (n_out == 0) return GCD_OK;
  const int vec = (c % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec && option(GCD_OPT_GATHER_FLAT))
    rows_gather_flat_kernel<<<(unsigned)ceil_div(n_out * (c / 4), kGatherThreads * kGatherUnroll), kGatherThreads, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(in), ld_in / 4, idx, n_out, c / 4, reinterpret_cast<float4*>(out), ld_out / 4);
  else
  rows_gather_kernel<<<(unsigned)ceil_div(n_out * 32, 256), 256, 0, as_stream(stream)>>>(in, ld_in, idx, n_out, c, out, ld_out, vec);
  GCD_LAUNCH_CHECK("gcd_rows_gather");
  return GCD_OK;
}

extern "C" size_t gcd_csr_workspace_bytes(int64_t n_points, int64_t n_segments) {
  (void)n_segments;
  int64_t n = n_points > 0 ? n_points : 1;
  return align_up((size_t)n * 8, 256) + radix_sort_workspace_bytes(n);
}

extern "C" int32_t gcd_csr_build(const int64_t* idx, int64_t n_points, int64_t n_segments, int32_t* seg_off, int32_t* order,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  GCD_REQUIRE(n_points >= 0 && n_points < (1ll << 31) && n_segments >= 0, "gcd_csr_build: sizes out of range");
  if (workspace_bytes < gcd_csr_workspace_bytes(n_points, n_segments)) { set_error("gcd_csr_build: workspace too small"); return GCD_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  char* p = static_cast<char*>(workspace);
  uint64_t* keys = (uint64_t*)p; p += align_up((size_t)(n_points > 0 ? n_points : 1) * 8, 256);
  if (n_points > 0) {
    csr_keys_kernel<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(idx, n_points, keys, order);
    int bits = 1;
    while ((1ll << bits) < n_segments) ++bits;
    int32_t rc = radix_sort_pairs(keys, order, n_points, bits, p, radix_sort_workspace_bytes(n_points), st);
    if (rc != GCD_OK) return rc;
  }
  csr_offsets_kernel<<<(unsigned)ceil_div(n_points + 1, 256), 256, 0, st>>>(keys, n_points, n_segments, seg_off);
  GCD_LAUNCH_CHECK("gcd_csr_build");
  return GCD_OK;
}

extern "C" int32_t gcd_segment_reduce(const float* in, int64_t ld_in, const int32_t* seg_off, const int32_t* order,
                                      int64_t n_segments, int32_t c, int32_t mode, float* out, int64_t ld_out, void* stream) {
  GCD_REQUIRE(mode >= 0 && mode <= 2, "gcd_segment_reduce: mode must be 0 (sum), 1 (mean) or 2 (max)");
  if (n_segments == 0) return GCD_OK;
  segment_reduce_kernel<<<(unsigned)ceil_div(n_segments * 32, 256), 256, 0, as_stream(stream)>>>(in, ld_in, seg_off, order, n_segments, c, mode, out, ld_out);
  GCD_LAUNCH_CHECK("gcd_segment_reduce");
  return GCD_OK;
}
