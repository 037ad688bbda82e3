// Microbenchmark: the cp.async (LDGSTS, 16 B, zero fill) row gather of the sparse convolution on its own —
// P producer warps fill 16 KB stages (128 rows x 64 bf16, 128B-swizzled), one warp consumes them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_gather ldgsts_gather.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity) {
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_dbg(uint32_t b, uint32_t parity, int who, int it) {
  const long long t0 = clock64();
  while (!mbar_try(b, parity)) {
    if (clock64() - t0 > 200000000ll) { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) printf("stuck: who %d it %d bar %u parity %u\n", who, it, b, parity); __trap(); }
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_pred(uint32_t b, uint32_t pred) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p mbarrier.arrive.shared::cta.b64 _, [%0];\n}" ::"r"(b), "r"(pred) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

constexpr int kStages = 6, kStageBytes = 16384;

template <int P, bool kSkip, int kLag, bool kElectPoll, int kVariant = 0>
__global__ void __launch_bounds__(32 * P + 32, 1) bench(const __nv_bfloat16* __restrict__ in, int ld, const int* __restrict__ idx, int iters,
                                                        long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  int* s_idx = (int*)(smem + kStages * kStageBytes + 256);          // [32][128] table, reused every 32 iterations
  for (int i = threadIdx.x; i < 32 * 128; i += blockDim.x) s_idx[i] = idx[(size_t)blockIdx.x * iters * 128 + i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kLag ? P : 32 * P); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < P) {
    constexpr int R = 32 / P;                 // rows per thread
    const int chunk = lane & 7, row0 = warp * 4 + (lane >> 3);
    const uint32_t a_base = smem_u32(smem) + row0 * 128 + ((chunk ^ (row0 & 7)) << 4);
    const char* col = reinterpret_cast<const char*>(in + chunk * 8);
    const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    const int* my = s_idx + row0;
    uint32_t st = 0, ph = 0;
    int r[R];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < R; ++j) r[j] = my[(it & 31) * 128 + j * 4 * P];
      if (kVariant == 0) { if (kElectPoll) { if (lane == 0) mbar_wait(empty0 + st * 8, ph ^ 1); __syncwarp(); } else mbar_wait(empty0 + st * 8, ph ^ 1); }
      const uint32_t a = a_base + st * kStageBytes;
#pragma unroll
      for (int j = 0; j < R; ++j) if (!kSkip || r[j] >= 0) cp_async_16(a + j * (4 * P * 128), col + (int64_t)max(r[j], 0) * ld * 2, r[j] >= 0 ? 16u : 0u);
      if (kLag == 0) cp_arrive(full0 + st * 8);
      else {
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= kLag) {
          asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          const uint32_t pst = (st + kStages - kLag) % kStages;
          mbar_arrive_pred(full0 + pst * 8, lane == 0 ? 1u : 0u);
        }
      }
      if (++st == kStages) { st = 0; ph ^= 1; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (kLag) {
      __syncwarp();
      for (int l = kLag; l >= 1; --l) { const uint32_t pst = (st + kStages - l) % kStages; if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + pst * 8) : "memory"); }
    }
  } else {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters && kVariant != 2; ++it) {
      mbar_wait(smem_u32(&full[st]), ph);
      mbar_arrive_pred(smem_u32(&empty[st]), lane == 0 ? 1u : 0u);       // predicated, no divergent branch
      if (++st == kStages) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// Stage ownership: producer warp w fills the WHOLE stage of iterations g == w (mod P) (32 copies per lane, one completion
// arrival per lane), so up to min(P, stages) stages are being filled concurrently and the per-stage fixed costs
// (wait, index loads, arrive, loop) are paid by one warp instead of all of them in lock-step.
template <int P, int S>
__global__ void __launch_bounds__(32 * P + 32, 1) bench_own(const __nv_bfloat16* __restrict__ in, int ld, const int* __restrict__ idx, int iters,
                                                            long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S * kStageBytes);
  uint64_t* empty = full + S;
  int* s_idx = (int*)(smem + S * kStageBytes + 256);
  for (int i = threadIdx.x; i < 32 * 128; i += blockDim.x) s_idx[i] = idx[(size_t)blockIdx.x * iters * 128 + i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 32); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < P) {
    const int chunk = lane & 7, rsub = lane >> 3;
    const uint32_t d_even = rsub * 128 + ((chunk ^ rsub) << 4), d_odd = (rsub + 4) * 128 + ((chunk ^ (rsub + 4)) << 4);
    const uint32_t a0 = smem_u32(smem);
    const char* col = reinterpret_cast<const char*>(in + chunk * 8);
    const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    const int64_t ldb = (int64_t)ld * 2;
    uint32_t st = warp % S, ph = (warp / S) & 1;
    for (int it = warp; it < iters; it += P) {
      const int* nb = s_idx + (it & 31) * 128 + rsub;
      mbar_wait_dbg(empty0 + st * 8, ph ^ 1, warp, it);
      const uint32_t a = a0 + st * kStageBytes;
#pragma unroll
      for (int jb = 0; jb < 32; jb += 8) {
        int r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = nb[(jb + j) * 4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jj = jb + j;
          cp_async_16(a + (jj >> 1) * 1024 + ((jj & 1) ? d_odd : d_even), col + (int64_t)max(r[j], 0) * ldb, r[j] >= 0 ? 16u : 0u);
        }
      }
      cp_arrive(full0 + st * 8);
      st += P % S; ph ^= (P / S) & 1;
      if (st >= (uint32_t)S) { st -= S; ph ^= 1; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_dbg(smem_u32(&full[st]), ph, 100, it);
      mbar_arrive_pred(smem_u32(&empty[st]), lane == 0 ? 1u : 0u);
      if (++st == S) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  const int n_rows = 200000, c = 128, iters = 2000, grid = 148;
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, (size_t)n_rows * c * 2));
  CK(cudaMemset(d, 0, (size_t)n_rows * c * 2));
  long long* cyc;
  CK(cudaMalloc(&cyc, grid * 8));
  const size_t smem = kStages * kStageBytes + 1024 + 256 + 32 * 128 * 4;
  for (int mode = 0; mode < 4; ++mode) {
    std::vector<int> idx((size_t)grid * iters * 128);
    uint32_t s = 12345u + mode;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (size_t i = 0; i < idx.size(); ++i) {
      int v = (int)(rnd() % n_rows);
      if (mode == 3) v = (int)(((i / 128) * 131 + (rnd() % 512)) % n_rows);
      bool valid = mode == 0 || ((mode == 1 || mode == 3) && rnd() % 100 < 30);
      idx[i] = valid ? v : -1;
    }
    int* didx;
    CK(cudaMalloc(&didx, idx.size() * 4));
    CK(cudaMemcpy(didx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
    auto run = [&](auto kern, int P) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int rep = 0; rep < 2; ++rep) { kern<<<grid, 32 * P + 32, smem>>>(d, c, didx, iters, cyc); CK(cudaDeviceSynchronize()); }
      std::vector<long long> hc(grid);
      CK(cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto v : hc) avg += (double)v; avg /= grid;
      printf("mode %d (%s), %2d producer warps: %.1f cycles per 16 KB stage per SM\n", mode,
             mode == 0 ? "all valid, random" : mode == 1 ? "30% valid, random" : mode == 2 ? "none valid" : "30% valid, clustered", P, avg / iters);
    };
    printf(" baseline:\n"); run(bench<8, false, 0, false, 0>, 8);
    printf(" stage ownership (warp w fills whole stages g = w mod P), 6 stages:\n"); run(bench_own<8, 6>, 8); run(bench_own<12, 6>, 12);
    printf(" stage ownership, 4 stages:\n"); run(bench_own<4, 4>, 4); run(bench_own<8, 4>, 8);
    printf(" stage ownership, 3 stages:\n"); run(bench_own<4, 3>, 4); run(bench_own<8, 3>, 8);
    printf(" producers never wait for a free stage (consumer still consumes):\n"); run(bench<8, false, 0, false, 1>, 8); run(bench<16, false, 0, false, 1>, 16);
    printf(" producers free-running, no consumer:\n"); run(bench<4, false, 0, false, 2>, 4); run(bench<8, false, 0, false, 2>, 8); run(bench<16, false, 0, false, 2>, 16);
    printf(" producers free-running, no consumer, skipping invalid rows:\n"); run(bench<8, true, 0, false, 2>, 8);
    CK(cudaFree(didx));
  }
  return 0;
}
