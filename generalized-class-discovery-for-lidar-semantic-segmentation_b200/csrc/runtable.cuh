// runtable.cuh — "run table": a voxel hash whose slots hold RUNS of four x-adjacent cells, and the stride-1 kernel-map
// search over it (same result as kmap_subm_kernel in coords.cu; ref: MinkowskiEngine's kernel-map generation behind
// every MinkowskiConvolution(kernel_size=3|5, stride=1), models/minkunet.py:62-128).
//
// Why: the point-wise table costs one scattered 8-byte key load plus one dependent 4-byte value load per neighbour
// offset (26 / 124 of them per voxel for K = 3 / 5), and a fully divergent warp load is served at one 32-byte sector per
// cycle per SM -- that rate, not DRAM or L2 bandwidth, is what kmap_subm_kernel runs at (profiles/README.md).  The K
// neighbours of one (y, z) row are x-adjacent, so a slot that stores the four cells x0 .. x0 + 3 of one (b, y, z) row
// answers them with one or two 32-byte loads (one LDG.E.256 each): 9 x 1.5 = 13.5 scattered sector loads per voxel
// for K = 3 instead of ~60, 25 x 2 = 50 for K = 5 instead of ~300.
//
// This header holds per-thread device functions only (no kernels, no launches) and needs nothing but keys.cuh and three
// atomics (CAS, exchange, or), so tests/emu/ compiles the very same source with g++ and checks it against the CPU oracle on a box without
// a GPU (threads run one after the other there, which is a legal schedule for kernels without barriers).
#pragma once
#include "keys.cuh"

namespace gcd {

// One slot = one 32-byte sector: key of (b, run, y, z) with run = floor(x / ts) >> 2, and the voxel rows of the cells
// 4 run + {0, 1, 2, 3} (-1 = no voxel).  `spare` keeps the slot at sector size; it is never read.
struct __attribute__((aligned(32))) RunSlot {
  unsigned long long key;
  int32_t v[4];
  unsigned long long spare;
};
static_assert(sizeof(RunSlot) == 32, "a run-table slot is one 32-byte sector");

GCD_DEVFN int run_floor_div(int a, int s) {  // floor(a / s), s > 0
  const int q = a / s;
  return (a % s != 0 && a < 0) ? q - 1 : q;
}

GCD_DEVFN void run_coord_load(const int32_t* coords, int64_t i, int& b, int& x, int& y, int& z) {
#if defined(__CUDA_ARCH__)
  const int4 c = *reinterpret_cast<const int4*>(coords + i * 4);
  b = c.x; x = c.y; y = c.z; z = c.w;
#else
  b = coords[i * 4]; x = coords[i * 4 + 1]; y = coords[i * 4 + 2]; z = coords[i * 4 + 3];
#endif
}

GCD_DEVFN RunSlot run_slot_load(const RunSlot* p) {
#if defined(__CUDA_ARCH__)
  RunSlot s;
  unsigned long long w1, w2;
  // one 256-bit load (LDG.E.256): the slot arrives with a single request per lane
  asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(s.key), "=l"(w1), "=l"(w2), "=l"(s.spare) : "l"(p));
  s.v[0] = (int32_t)(uint32_t)w1; s.v[1] = (int32_t)(uint32_t)(w1 >> 32);
  s.v[2] = (int32_t)(uint32_t)w2; s.v[3] = (int32_t)(uint32_t)(w2 >> 32);
  return s;
#else
  return *p;
#endif
}

GCD_DEVFN void run_slot_clear(RunSlot* slots, int64_t t, int64_t cap) {
  if (t < cap) { slots[t].key = kEmptyKey; slots[t].v[0] = slots[t].v[1] = slots[t].v[2] = slots[t].v[3] = -1; slots[t].spare = 0; }
}

// Insert voxel row i.  Status bits: key range (also: x not a multiple of ts), duplicate coordinate, table full.
GCD_DEVFN void run_insert_thread(int64_t i, const int32_t* coords, int ts, RunSlot* slots, int64_t cap, int32_t* status,
                                 int dev_key_range, int dev_duplicate, int dev_table_full) {
  int b, x, y, z;
  run_coord_load(coords, i, b, x, y, z);
  const int cell = run_floor_div(x, ts);
  if (!key_in_range(b, x, y, z) || cell * ts != x) { atomicOr(status, dev_key_range); return; }
  const unsigned long long key = pack_key(b, cell >> 2, y, z);
  int64_t slot = (int64_t)(hash_key(key) & (uint64_t)(cap - 1));
  for (int64_t probes = 0; probes < cap; ++probes) {
    unsigned long long cur = slots[slot].key;
    if (cur == kEmptyKey) {                    // try to claim; a lost race returns the winner's key
      const unsigned long long prev = atomicCAS(&slots[slot].key, (unsigned long long)kEmptyKey, key);
      cur = (prev == kEmptyKey) ? key : prev;
    }
    if (cur == key) {                          // the run's slot: publish this voxel's row in its cell
      if (atomicExch(&slots[slot].v[cell & 3], (int32_t)i) != -1) atomicOr(status, dev_duplicate);
      return;
    }
    slot = (slot + 1) & (cap - 1);
  }
  atomicOr(status, dev_table_full);
}

GCD_DEVFN int32_t run_pick(const int32_t (&v)[4], int sub) {   // v[sub] without a dynamically indexed register array
  return sub == 0 ? v[0] : sub == 1 ? v[1] : sub == 2 ? v[2] : v[3];
}

// Finish a lookup whose first probe (slot `slot`, content `s`) is already in registers: the four rows of the run,
// all -1 when the run holds no voxel.  The loop only runs on a hash collision.
GCD_DEVFN void run_resolve(const RunSlot* slots, int64_t cap, unsigned long long key, int64_t slot, RunSlot s, int32_t (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = -1;
  for (int64_t probes = 0; probes < cap; ++probes) {
    if (s.key == key) { v[0] = s.v[0]; v[1] = s.v[1]; v[2] = s.v[2]; v[3] = s.v[3]; return; }
    if (s.key == kEmptyKey) return;
    slot = (slot + 1) & (cap - 1);
    s = run_slot_load(slots + slot);
  }
}

// Kernel map of output voxel o: nbr[((kz K + ky) K + kx) n + o] = row of the voxel at (x + (kx-R) ts, y + (ky-R) ts,
// z + (kz-R) ts) or -1; the centre tap is o itself (as in kmap_subm_kernel).  K consecutive cells touch at most two
// runs of four (K <= 5), so a (ky, kz) row costs one or two slot loads; both first probes are issued before either is
// looked at (two independent 32-byte loads in flight per thread, the rest of the parallelism comes from occupancy).
template <int K>
GCD_DEVFN void kmap_runs_thread(int64_t o, const int32_t* coords, int64_t n, const RunSlot* slots, int64_t cap, int ts, int32_t* nbr) {
  static_assert(K == 3 || K == 5, "K consecutive cells must fit two runs of four");
  constexpr int R = K / 2;
  int b, x, y, z;
  run_coord_load(coords, o, b, x, y, z);
  const int cell = run_floor_div(x, ts);
  const int run_lo = (cell - R) >> 2, run_hi = (cell + R) >> 2;
  const bool two = run_hi != run_lo;
#pragma unroll 1
  for (int kz = 0; kz < K; ++kz) {
    const int zz = z + (kz - R) * ts;
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
      const int yy = y + (ky - R) * ts;
      int32_t lo[4] = {-1, -1, -1, -1}, hi[4] = {-1, -1, -1, -1};
      if (key_in_range(b, 0, yy, zz)) {
        const unsigned long long key_lo = pack_key(b, run_lo, yy, zz), key_hi = pack_key(b, run_hi, yy, zz);
        const int64_t slot_lo = (int64_t)(hash_key(key_lo) & (uint64_t)(cap - 1));
        const int64_t slot_hi = (int64_t)(hash_key(key_hi) & (uint64_t)(cap - 1));
        const RunSlot s_lo = run_slot_load(slots + slot_lo);
        const RunSlot s_hi = run_slot_load(slots + (two ? slot_hi : slot_lo));   // same address when there is one run: no second sector
        run_resolve(slots, cap, key_lo, slot_lo, s_lo, lo);
        if (two) run_resolve(slots, cap, key_hi, slot_hi, s_hi, hi);
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int cc = cell + kx - R;
        int32_t r = ((cc >> 2) == run_lo) ? run_pick(lo, cc & 3) : run_pick(hi, cc & 3);
        if (!key_in_range(b, x + (kx - R) * ts, yy, zz)) r = -1;     // kmap_subm_kernel tests the range per offset
        if (kx == R && ky == R && kz == R) r = (int32_t)o;
        nbr[(int64_t)((kz * K + ky) * K + kx) * n + o] = r;
      }
    }
  }
}

}  // namespace gcd
