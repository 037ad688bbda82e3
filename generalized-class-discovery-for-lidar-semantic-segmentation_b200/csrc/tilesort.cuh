// tilesort.cuh — per-thread bodies of the tile sort of a 3x3x3 (or 2x2x2) neighbour table (tilesort.cu).
//
// The forward / dgrad convolution works on tiles of 128 table columns (output rows) and visits a kernel offset only if
// some column of the tile has a neighbour there.  In scan order almost every offset has a hit in every tile (21-25 of 27
// on LiDAR sweeps, 21-34 % of the gathered rows are real); columns with the same set of present neighbours -- the ground
// plane, walls, isolated returns -- are far apart.  Sorting the columns by their 27-bit presence mask puts like with
// like: 8-12 offsets per tile and 60-70 % real rows on the same sweeps (tests/test_tile_sort_model.py), i.e. 2.2-2.6x fewer
// stages for the same result.  The bit order matters: the rarest offsets decide the order, so corners (three non-zero
// components) take the top bits, then edges, faces, and the always-present centre the lowest bit.
// The kernel then reads the permuted table and writes tile column i to output row out_rows[i].
// No CUDA headers: tests/emu/ compiles the same source with g++.
#pragma once
#include <stdint.h>
#include "keys.cuh"   // GCD_DEVFN

namespace gcd {

GCD_DEVFN int tile_sort_class(int k) {       // number of non-zero components of offset k of a 3x3x3 kernel (x fastest)
  return ((k % 3) != 1) + (((k / 3) % 3) != 1) + ((k / 9) != 1);
}
// Bit of offset k in the sort key: offsets ordered by (class, k); folds to a constant when k is one.
GCD_DEVFN int tile_sort_bit(int k) {
  const int c = tile_sort_class(k);
  int bit = 0;
  for (int j = 0; j < 27; ++j) {
    const int cj = tile_sort_class(j);
    bit += (cj < c || (cj == c && j < k)) ? 1 : 0;
  }
  return bit;
}

// keys[o] = presence mask of table column o, vals[o] = o (the pair is then sorted by key, stably).
GCD_DEVFN void tile_sort_key_thread(int64_t o, const int32_t* nbr, int64_t n, unsigned long long* keys, int32_t* vals) {
  unsigned long long key = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) key |= (unsigned long long)(nbr[(int64_t)k * n + o] >= 0 ? 1 : 0) << tile_sort_bit(k);
  keys[o] = key;
  vals[o] = (int32_t)o;
}

// 2x2x2 tables (stride-2 and transposed convolutions): bit k of the key = offset k present.  A transposed map has exactly
// one entry per column, so the sort groups the columns by child position: one offset per tile instead of all eight.
GCD_DEVFN void tile_sort_key8_thread(int64_t o, const int32_t* nbr, int64_t n, unsigned long long* keys, int32_t* vals) {
  unsigned long long key = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) key |= (unsigned long long)(nbr[(int64_t)k * n + o] >= 0 ? 1 : 0) << k;
  keys[o] = key;
  vals[o] = (int32_t)o;
}

// Offsets with at least one hit among the columns whose keys were OR-ed into ``key_or``, as a mask with bit k = offset k
// (the layout the convolution kernel walks): undoes the key's bit order for 3x3x3 tables.
GCD_DEVFN uint32_t tile_mask_from_keys(unsigned long long key_or, int kv) {
  if (kv != 27) return (uint32_t)key_or;
  uint32_t mask = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) mask |= (uint32_t)((key_or >> tile_sort_bit(k)) & 1ull) << k;
  return mask;
}

// sorted[k][i] = nbr[k][rows[i]]: writes coalesced, reads gathered (once per kernel map).
GCD_DEVFN void tile_sort_permute_thread(int64_t t, const int32_t* nbr, int64_t n, int kv, const int32_t* rows, int32_t* sorted) {
  if (t >= n * kv) return;
  const int64_t k = t / n, i = t - k * n;
  sorted[t] = nbr[k * n + rows[i]];
}

}  // namespace gcd
