// runtable.cu — kernels and C entry points of the run table (runtable.cuh): build and stride-1 kernel-map search.
// Opt-in alternative to gcd_hash_build + gcd_kmap_subm (GCDLSS_KMAP=runs, gcdlss_b200/coords.py); results are identical.
#include "common.cuh"
#include "runtable.cuh"

namespace gcd {
namespace {
constexpr int kThreads = 256;
inline unsigned grid_for(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, kThreads); }

__global__ void __launch_bounds__(kThreads) run_clear_kernel(RunSlot* slots, int64_t cap) {
  run_slot_clear(slots, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, cap);
}
__global__ void __launch_bounds__(kThreads) run_insert_kernel(const int32_t* __restrict__ coords, int64_t n, int ts, RunSlot* slots,
                                                               int64_t cap, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run_insert_thread(i, coords, ts, slots, cap, status, GCD_DEV_KEY_RANGE, GCD_DEV_DUPLICATE, GCD_DEV_TABLE_FULL);
}
template <int K>
__global__ void __launch_bounds__(kThreads) kmap_runs_kernel(const int32_t* __restrict__ coords, int64_t n, const RunSlot* __restrict__ slots,
                                                              int64_t cap, int ts, int32_t* __restrict__ nbr) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) kmap_runs_thread<K>(o, coords, n, slots, cap, ts, nbr);
}
}  // namespace
}  // namespace gcd

using namespace gcd;

extern "C" size_t gcd_runtable_slot_bytes(void) { return sizeof(RunSlot); }

extern "C" int32_t gcd_runtable_build(const int32_t* coords, int64_t n, int32_t ts, void* slots, int64_t cap, int32_t* status,
                                      void* stream) {
  GCD_REQUIRE(n >= 0 && n < (1ll << 30), "gcd_runtable_build: n out of range");
  GCD_REQUIRE(ts >= 1 && ts <= (1 << 14), "gcd_runtable_build: tensor stride out of range");
  GCD_REQUIRE(slots && status && cap >= 2 * n && cap >= 1 && (cap & (cap - 1)) == 0, "gcd_runtable_build: capacity must be a power of two >= 2n");
  GCD_REQUIRE((reinterpret_cast<uintptr_t>(slots) & 31) == 0, "gcd_runtable_build: the slot array must be 32-byte aligned");
  cudaStream_t st = as_stream(stream);
  run_clear_kernel<<<grid_for(cap), kThreads, 0, st>>>(static_cast<RunSlot*>(slots), cap);
  if (n > 0) run_insert_kernel<<<grid_for(n), kThreads, 0, st>>>(coords, n, ts, static_cast<RunSlot*>(slots), cap, status);
  GCD_LAUNCH_CHECK("gcd_runtable_build");
  return GCD_OK;
}

extern "C" int32_t gcd_kmap_subm_runs(const int32_t* coords, int64_t n, const void* slots, int64_t cap, int32_t kernel_size,
                                      int32_t ts, int32_t* nbr, void* stream) {
  GCD_REQUIRE(kernel_size == 3 || kernel_size == 5, "gcd_kmap_subm_runs: kernel_size must be 3 or 5 (got %d)", kernel_size);
  GCD_REQUIRE(n >= 0 && ts >= 1 && cap >= 1 && (cap & (cap - 1)) == 0, "gcd_kmap_subm_runs: bad arguments");
  GCD_REQUIRE((reinterpret_cast<uintptr_t>(slots) & 31) == 0, "gcd_kmap_subm_runs: the slot array must be 32-byte aligned");
  if (n == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  const RunSlot* s = static_cast<const RunSlot*>(slots);
  if (kernel_size == 3) kmap_runs_kernel<3><<<grid_for(n), kThreads, 0, st>>>(coords, n, s, cap, ts, nbr);
  else                  kmap_runs_kernel<5><<<grid_for(n), kThreads, 0, st>>>(coords, n, s, cap, ts, nbr);
  GCD_LAUNCH_CHECK("gcd_kmap_subm_runs");
  return GCD_OK;
}
