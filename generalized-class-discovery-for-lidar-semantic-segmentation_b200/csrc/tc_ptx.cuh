// tc_ptx.cuh — thin inline-PTX wrappers for sm_100a: mbarrier, cp.async, bulk copy (TMA engine),
// tcgen05 (TMEM alloc / mma / commit / ld) and the shared-memory / instruction descriptors.
#pragma once
#include <stdint.h>

namespace gcd {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// address-taking variants (the caller keeps shared-window addresses in registers)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}"
      ::"r"(bar_addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc_addr(uint32_t bar_addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// Warp-level wait: one lane polls (optionally backing off), the rest of the warp parks on __syncwarp.  The
// mbarrier unit serialises every arrival and every poll, so per-thread polling by hundreds of threads is what a
// pipeline stage ends up costing (measured: ~460 cycles per stage for the barrier traffic alone).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, unsigned sleep_ns = 0) {
  if ((threadIdx.x & 31) == 0) {
    while (!mbar_try_wait(bar, parity)) {
      if (sleep_ns) __nanosleep(sleep_ns);
    }
  }
  __syncwarp();
}
// true in exactly one (converged) lane of the warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Explicit shared-memory load: pointers derived from the 1024-byte-aligned dynamic smem base lose their address space and
// compile to generic LD.E otherwise.
__device__ __forceinline__ int lds_s32(uint32_t smem_addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
  return v;
}

// release / acquire on a shared-memory word (CTA scope): a role publishes a counter after the data it guards
__device__ __forceinline__ void st_release_shared_u32(uint32_t smem_addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared_u32(uint32_t smem_addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- async copies
// 16-byte cp.async (LDGSTS) with zero fill: src_bytes = 16 copies, 0 writes zeros.
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// predicated form (per-lane predicate, no branch)
__device__ __forceinline__ void cp_async_16_pred(uint32_t dst_smem, const void* src, uint32_t src_bytes, uint32_t pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}"
               ::"r"(dst_smem), "l"(src), "r"(src_bytes), "r"(pred) : "memory");
}
// 4-byte cp.async (cache-all; .cg only takes 16 bytes): the table warp's copies of unaligned neighbour-table slices
__device__ __forceinline__ void cp_async_4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// The calling thread's prior cp.async copies, once complete, perform one arrival on the mbarrier
// (.noinc: the arrival is part of the barrier's expected count).  This is how the gather producers
// publish a stage to the MMA warp without ever waiting for their own copies.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// predicated forms for the elected lane of a converged warp (no branch, see mma_bf16_ss_pred)
__device__ __forceinline__ void mbar_arrive_pred(uint64_t* bar, uint32_t pred) {
  asm volatile("{\n\t.reg .b64 st;\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q mbarrier.arrive.shared::cta.b64 st, [%0];\n\t}"
               ::"r"(smem_u32(bar)), "r"(pred) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_pred(uint64_t* bar, uint32_t bytes, uint32_t pred) {
  asm volatile("{\n\t.reg .b64 st;\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
               ::"r"(smem_u32(bar)), "r"(bytes), "r"(pred) : "memory");
}
__device__ __forceinline__ void bulk_g2s_pred(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint32_t pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
               "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "r"(pred) : "memory");
}

// 32-byte global store (sm_100: STG.E.256); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// 16-byte vector reduction (sm_90+): four fp32 adds in one L2 operation
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Predicated forms: every lane of the (converged) warp executes the instruction slot, only lanes with pred != 0 issue.
// Measured on B200: wrapping the issue in `if (elected) { ... } __syncwarp();` costs ~230 cycles per loop iteration
// for the divergence / reconvergence alone; the predicated form has no branch.
__device__ __forceinline__ void mma_bf16_ss_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(pred)
      : "memory");
}
// Same, descriptors given as (low word, shared high word): the high word (LBO / SBO / layout bits) never changes, only the
// 14-bit start-address field in the low word does, so the issuing warp carries two 32-bit values instead of two 64-bit ones.
__device__ __forceinline__ void mma_bf16_ss_pred_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                    uint32_t accumulate, uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void mma_commit_pred(uint64_t* bar, uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(pred)
      : "memory");
}
// All previously issued MMAs of this thread complete -> one arrival on the mbarrier.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1 (sm_100).
// lbo / sbo in bytes.  K-major operand: rows of 128 B (64 bf16 along K), 8-row groups sbo apart,
// lbo unused.  MN-major operand: 128-B rows hold 64 bf16 along M/N, consecutive rows step K,
// 8-row (K) groups sbo apart, 64-element M/N blocks lbo apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ | (a_mn_major << 15) | (b_mn_major << 16) |
         ((n >> 3) << 17) | ((128u >> 4) << 24);
}

}  // namespace ptx
}  // namespace gcd
