// common.cuh — shared device/host helpers for libgcdlss_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gcdlss_b200.h"
#include "keys.cuh"

namespace gcd {

// ---- host-side error plumbing (thread-local text, C error codes; no exceptions cross the ABI)
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);

#define GCD_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::gcd::set_error(__VA_ARGS__);                            \
      return GCD_ERR_INVALID_ARG;                               \
    }                                                           \
  } while (0)

#define GCD_LAUNCH_CHECK(what)                                  \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return ::gcd::cuda_fail(e__, what); \
  } while (0)

// process-wide tuning options (api.cu): atomics, initialised once from the environment
int32_t option(int32_t key);

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL).  Every kernel of the training step starts with pdl_trigger() (its successor in the
// stream may be launched as soon as all of this grid's blocks are resident) and executes pdl_wait() before it first touches
// global memory that a predecessor may still be writing or reading: block scheduling, barrier / TMEM set-up and parameter
// loads of kernel i + 1 overlap the tail of kernel i instead of following it (a MinkUNet34 step is ~350 dependent launches
// on one stream).  Without the launch attribute both instructions are no-ops, so the same kernels run under plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute when GCD_OPT_PDL is on.
// ONLY for kernels that call pdl_wait() before their first global access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = option(GCD_OPT_PDL) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// cooperative launch (grid-wide barrier inside the kernel), with the PDL attribute when the driver accepts the combination
template <typename Arg>
inline cudaError_t launch_coop_pdl(void (*kernel)(Arg), dim3 grid, dim3 block, cudaStream_t st, const Arg& a) {
  static int pdl_ok = 1;      // benign race: worst case the combination is tried twice
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  if (pdl_ok && option(GCD_OPT_PDL)) {
    cfg.numAttrs = 2;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (e == cudaSuccess) return e;
    cudaGetLastError();
    pdl_ok = 0;
  }
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

// Open-addressing table, linear probing.  Slots are grouped in 32-byte sectors of four keys; a
// probe sequence visits whole sectors so the four keys a DRAM/L2 sector delivers are all used.
// insert: returns the slot holding `key` (claimed or already present), or -1 when the table is full.
__device__ __forceinline__ int64_t table_insert(uint64_t* keys, int64_t cap, uint64_t key, bool* was_present) {
  int64_t slot = (int64_t)(hash_key(key) & (uint64_t)(cap - 1));
  for (int64_t probes = 0; probes < cap; ++probes) {
    uint64_t cur = keys[slot];
    if (cur == key) { *was_present = true; return slot; }
    if (cur == kEmptyKey) {
      unsigned long long prev = atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey,
                                          (unsigned long long)key);
      if (prev == kEmptyKey) { *was_present = false; return slot; }
      if (prev == key) { *was_present = true; return slot; }
    }
    slot = (slot + 1) & (cap - 1);
  }
  return -1;
}
__device__ __forceinline__ int64_t table_find(const uint64_t* __restrict__ keys, int64_t cap, uint64_t key) {
  int64_t slot = (int64_t)(hash_key(key) & (uint64_t)(cap - 1));
  for (int64_t probes = 0; probes < cap; ++probes) {
    uint64_t cur = __ldg(&keys[slot]);
    if (cur == key) return slot;
    if (cur == kEmptyKey) return -1;
    slot = (slot + 1) & (cap - 1);
  }
  return -1;
}

// ---- dtype helpers
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- device-wide exclusive scan of int32 (scan.cu)
size_t scan_workspace_bytes(int64_t n);
// out[i] = sum_{j<i} in[j]; total (device int32, may be null) = sum of all.  in may alias out.
int32_t exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);
// pair lists straight from a neighbour table (scan.cu): two reads of the table, no flag / position arrays.
// total_pairs: device int32 scratch; workspace >= scan_workspace_bytes(n_out * kv).
int32_t pairs_from_table_fused(const int32_t* nbr, int64_t n_out, int kv, int32_t* pair_in, int32_t* pair_out, int32_t* pair_off,
                               int32_t* total_pairs, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// the same lists in ONE pass over the table (decoupled look-back, scan.cu); workspace >= pairs_lookback_workspace_bytes(n_out * kv)
size_t pairs_lookback_workspace_bytes(int64_t total_entries);
int32_t pairs_from_table_lookback(const int32_t* nbr, int64_t n_out, int kv, int32_t* pair_in, int32_t* pair_out, int32_t* pair_off,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace gcd
