// cuda_shim.h — the handful of CUDA builtins the integer kernels' per-thread functions use, for a g++ build.
// Threads of an emulated launch run one after the other (in an order the test chooses), which is a legal schedule for
// kernels that have no barriers and no warp-level primitives; atomics therefore reduce to their sequential meaning.
// Test infrastructure only (tests/test_emulated_kernels.py).
#pragma once
#include <stdint.h>

static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long expected, unsigned long long desired) {
  const unsigned long long old = *a;
  if (old == expected) *a = desired;
  return old;
}
static inline int atomicExch(int* a, int v) { const int old = *a; *a = v; return old; }
static inline int atomicOr(int* a, int v) { const int old = *a; *a = old | v; return old; }
static inline int atomicMin(int* a, int v) { const int old = *a; if (v < old) *a = v; return old; }
