// gather_rows.cuh — per-thread body of the vectorised devoxelisation gather out[i, :] = in[idx[i], :]
// (ref models/decoder.py:416-424, modules/exp_merge_mean_teacher.py:2845-2846: voxel features indexed by the inverse map).
// The [n_out x c/4] float4 elements are dealt to threads flat, so no lane idles when c/4 is not a multiple of 32 (c = 96: a
// warp per row leaves a quarter of the lanes without work), and each thread keeps kGatherUnroll independent
// index -> row-load chains in flight.  No CUDA headers: tests/emu/ compiles the same source with g++.
#pragma once
#include <stdint.h>
#include "keys.cuh"   // GCD_DEVFN

namespace gcd {

constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 4;

// V = float4 (16 bytes).  ld_in4 / ld_out4: row pitches in units of V; c4 = channels / 4.
template <typename V>
GCD_DEVFN void rows_gather_flat_thread(int64_t block, int thread, const V* in, int64_t ld_in4, const int64_t* idx, int64_t n_out,
                                       int c4, V* out, int64_t ld_out4) {
  const int64_t total = n_out * c4;
  const int64_t e0 = block * (int64_t)(kGatherThreads * kGatherUnroll) + thread;
  if (e0 >= total) return;
  const int64_t row0 = e0 / c4;
  const int j0 = (int)(e0 - row0 * c4);
  V v[kGatherUnroll];
  int64_t dst[kGatherUnroll];
#pragma unroll
  for (int u = 0; u < kGatherUnroll; ++u) {
    const int64_t e = e0 + (int64_t)u * kGatherThreads;
    const unsigned jj = (unsigned)j0 + (unsigned)(u * kGatherThreads);     // < c4 + 768: 32-bit division
    const int64_t row = row0 + jj / (unsigned)c4;
    const int j = (int)(jj % (unsigned)c4);
    dst[u] = -1;
    if (e < total) {
      dst[u] = row * ld_out4 + j;
#if defined(__CUDA_ARCH__)
      v[u] = __ldg(in + __ldg(idx + row) * ld_in4 + j);
#else
      v[u] = in[idx[row] * ld_in4 + j];
#endif
    }
  }
#pragma unroll
  for (int u = 0; u < kGatherUnroll; ++u)
    if (dst[u] >= 0) out[dst[u]] = v[u];
}

}  // namespace gcd
