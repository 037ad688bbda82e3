// Microbenchmark: how long does the issuing lane spend per tcgen05.mma / tcgen05.commit / fence on sm_100a?
// One CTA, the MMA warp pattern of conv_tc.cu (warp-uniform loop, elected lane issues).  Prints cycles per iteration.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../generalized-class-discovery-for-lidar-semantic-segmentation_b200/csrc/tc_ptx.cuh"
using namespace gcd::ptx;

__global__ void __launch_bounds__(128, 1) k(int n_cols, int ksteps, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<512>(&tmem_base_s); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)n_cols, 0, 0);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db = make_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    const bool leader = elect_one();
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode & 4) tc_fence_after();
      if (mode & 8) {      // predicated issue, no branch
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_ss_pred(tmem, da + ks * 2, db + ks * 2, idesc, (it | ks) != 0, (leader && ks < ksteps) ? 1u : 0u);
        mma_commit_pred(&bar[0], (leader && (mode & 1)) ? 1u : 0u);
      } else {
        if (leader) {
          for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ss(tmem, da + ks * 2, db + ks * 2, idesc, (it | ks) != 0);
          if (mode & 1) mma_commit(&bar[0]);
        }
        __syncwarp();
      }
      if (mode & 2) { mbar_wait(&bar[0], ph); ph ^= 1; }      // wait for the commit (full MMA latency exposed)
    }
    const long long t1 = clock64();
    if (leader) mma_commit(&bar[1]);
    __syncwarp();
    mbar_wait(&bar[1], 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  printf("%6s %6s %28s %12s %12s\n", "N", "ksteps", "mode", "issue cyc/it", "total cyc/it");
  const char* names[16] = {"mma only", "branch: mma+commit", "-", "mma+commit+wait", "fence+mma", "fence+mma+commit", "", "", "", "predicated: mma+commit", "", "predicated: +wait"};
  for (int n : {32, 96, 256}) for (int ks : {1, 2, 4}) for (int mode : {1, 9, 11}) {
    k<<<1, 128, 64 * 1024>>>(n, ks, iters, mode, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("%6d %6d %28s %12.1f %12.1f\n", n, ks, names[mode], (double)h[0] / iters, (double)h[1] / iters);
  }
  // commit only, fence only
  for (int mode : {1, 4}) {
    k<<<1, 128, 64 * 1024>>>(32, 0, iters, mode, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%6s %6d %28s %12.1f %12.1f\n", "-", 0, mode == 1 ? "commit only" : "fence only", (double)h[0] / iters, (double)h[1] / iters);
  }
  return 0;
}
