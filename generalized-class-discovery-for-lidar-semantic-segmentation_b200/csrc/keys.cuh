// keys.cuh — packed 64-bit voxel keys and their hash.  No CUDA headers: the same source is compiled by nvcc for the
// device and by g++ for the host emulation of the integer kernels (tests/emu/), so the CPU suite runs the code the GPU runs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GCD_DEVFN __device__ __forceinline__
#else
#define GCD_DEVFN inline
#endif

namespace gcd {

// ---- 64-bit coordinate keys: 10 bits batch | 3 x 18 bits (coord + 2^17)
constexpr int kCoordBits = 18;
constexpr int kCoordBias = 1 << (kCoordBits - 1);
constexpr int kBatchBits = 10;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

GCD_DEVFN bool key_in_range(int b, int x, int y, int z) {
  const unsigned lim = 1u << kCoordBits;
  return (unsigned)b < (1u << kBatchBits) - 1u &&  // batch 1023 reserved: keeps every key != kEmptyKey
         (unsigned)(x + kCoordBias) < lim &&
         (unsigned)(y + kCoordBias) < lim && (unsigned)(z + kCoordBias) < lim;
}
GCD_DEVFN uint64_t pack_key(int b, int x, int y, int z) {
  return ((uint64_t)(unsigned)b << (3 * kCoordBits)) | ((uint64_t)(unsigned)(x + kCoordBias) << (2 * kCoordBits)) |
         ((uint64_t)(unsigned)(y + kCoordBias) << kCoordBits) | (uint64_t)(unsigned)(z + kCoordBias);
}
GCD_DEVFN void unpack_key(uint64_t k, int& b, int& x, int& y, int& z) {
  const uint64_t m = (1ull << kCoordBits) - 1;
  z = (int)(k & m) - kCoordBias;
  y = (int)((k >> kCoordBits) & m) - kCoordBias;
  x = (int)((k >> (2 * kCoordBits)) & m) - kCoordBias;
  b = (int)(k >> (3 * kCoordBits));
}
// murmur3 finaliser
GCD_DEVFN uint64_t hash_key(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

}  // namespace gcd
