// gather_emu.cpp — host emulation of rows_gather_flat_kernel: csrc/gather_rows.cuh compiled unchanged by g++, run for
// every (block, thread) of the launch configuration gcd_rows_gather uses.
#include <stdint.h>
struct float4 { float x, y, z, w; };
#include "gather_rows.cuh"

extern "C" void emu_rows_gather_flat(const float* in, int64_t ld_in, const int64_t* idx, int64_t n_out, int32_t c, float* out, int64_t ld_out) {
  using namespace gcd;
  const int c4 = c / 4;
  const int64_t per_block = (int64_t)kGatherThreads * kGatherUnroll;
  const int64_t blocks = (n_out * c4 + per_block - 1) / per_block;
  for (int64_t b = 0; b < blocks; ++b)
    for (int t = 0; t < kGatherThreads; ++t)
      rows_gather_flat_thread<float4>(b, t, reinterpret_cast<const float4*>(in), ld_in / 4, idx, n_out, c4, reinterpret_cast<float4*>(out), ld_out / 4);
}
