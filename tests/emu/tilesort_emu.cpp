// tilesort_emu.cpp — host emulation of csrc/tilesort.cu: the per-thread bodies of csrc/tilesort.cuh compiled unchanged by
// g++; the stable radix sort between them is std::stable_sort here.
#include <algorithm>
#include <numeric>
#include <vector>
#include <stdint.h>
#include "tilesort.cuh"

extern "C" {
int32_t emu_tile_sort_bit(int32_t k) { return gcd::tile_sort_bit(k); }
uint32_t emu_tile_mask_from_keys(unsigned long long key_or, int32_t kv) { return gcd::tile_mask_from_keys(key_or, kv); }

void emu_kmap_tile_sort(const int32_t* nbr, int64_t n, int32_t kv, int32_t* nbr_sorted, int32_t* out_rows, unsigned long long* keys_out) {
  std::vector<unsigned long long> keys(n);
  std::vector<int32_t> vals(n);
  for (int64_t o = 0; o < n; ++o) {
    if (kv == 27) gcd::tile_sort_key_thread(o, nbr, n, keys.data(), vals.data());
    else gcd::tile_sort_key8_thread(o, nbr, n, keys.data(), vals.data());
  }
  std::vector<int64_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
  for (int64_t i = 0; i < n; ++i) { out_rows[i] = vals[idx[i]]; if (keys_out) keys_out[i] = keys[idx[i]]; }
  for (int64_t t = 0; t < n * kv; ++t) gcd::tile_sort_permute_thread(t, nbr, n, kv, out_rows, nbr_sorted);
}
}
