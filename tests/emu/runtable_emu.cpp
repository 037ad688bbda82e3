// runtable_emu.cpp — host emulation of csrc/runtable.cu's launches: the per-thread functions of csrc/runtable.cuh are
// compiled unchanged by g++ and run for every thread index, in the order given by `order` (nullptr = ascending).
#include "cuda_shim.h"
#include "runtable.cuh"

using namespace gcd;

extern "C" {
int64_t emu_runtable_slot_bytes() { return (int64_t)sizeof(RunSlot); }

// status bits as in include/gcdlss_b200.h: 1 key range, 2 duplicate, 4 table full
void emu_runtable_build(const int32_t* coords, int64_t n, int32_t ts, void* slots, int64_t cap, int32_t* status, const int64_t* order) {
  RunSlot* s = static_cast<RunSlot*>(slots);
  for (int64_t t = 0; t < cap; ++t) run_slot_clear(s, t, cap);
  for (int64_t j = 0; j < n; ++j) run_insert_thread(order ? order[j] : j, coords, ts, s, cap, status, 1, 2, 4);
}

void emu_kmap_subm_runs(const int32_t* coords, int64_t n, const void* slots, int64_t cap, int32_t kernel_size, int32_t ts, int32_t* nbr) {
  const RunSlot* s = static_cast<const RunSlot*>(slots);
  for (int64_t o = 0; o < n; ++o) {
    if (kernel_size == 3) kmap_runs_thread<3>(o, coords, n, s, cap, ts, nbr);
    else                  kmap_runs_thread<5>(o, coords, n, s, cap, ts, nbr);
  }
}

// slot loads a lookup of every voxel's K^3 neighbourhood performs (first probes + collision probes): the figure the
// design is about (csrc/runtable.cuh header), reported by the test
int64_t emu_count_slot_loads(const int32_t* coords, int64_t n, const void* slots, int64_t cap, int32_t kernel_size, int32_t ts) {
  const RunSlot* s = static_cast<const RunSlot*>(slots);
  const int R = kernel_size / 2;
  int64_t loads = 0;
  for (int64_t o = 0; o < n; ++o) {
    const int b = coords[o * 4], x = coords[o * 4 + 1], y = coords[o * 4 + 2], z = coords[o * 4 + 3];
    const int cell = run_floor_div(x, ts);
    const int run_lo = (cell - R) >> 2, run_hi = (cell + R) >> 2;
    for (int kz = -R; kz <= R; ++kz)
      for (int ky = -R; ky <= R; ++ky) {
        if (!key_in_range(b, 0, y + ky * ts, z + kz * ts)) continue;
        for (int run = run_lo; run <= run_hi; ++run) {
          const unsigned long long key = pack_key(b, run, y + ky * ts, z + kz * ts);
          int64_t slot = (int64_t)(hash_key(key) & (uint64_t)(cap - 1));
          for (;;) { ++loads; if (s[slot].key == key || s[slot].key == kEmptyKey) break; slot = (slot + 1) & (cap - 1); }
        }
      }
  }
  return loads;
}
}
