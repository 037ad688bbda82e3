// Microbenchmark: TMA tile::gather4 (4 rows x 128 B per instruction, 128B swizzle) as the row gather of the sparse
// convolution.  One warp per CTA issues 32 gather4 per 16 KB stage; a second warp consumes the stages.
// Reports cycles per stage per SM and checks the shared-memory image of the first stage against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStages = 6, kStageBytes = 16384;

template <int P>
__global__ void __launch_bounds__(32 * P + 32, 1) bench(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int iters, int c0_slices,
                                               long long* cycles, uint8_t* dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < P) {
    constexpr int kOps = 32 / P;               // gather4 per warp per stage
    uint32_t st = 0, ph = 0;
    const int slot = warp * kOps + (lane % kOps);
    const int* my = idx + (size_t)blockIdx.x * iters * 128 + slot * 4;
    int4 r = *reinterpret_cast<const int4*>(my);
    for (int it = 0; it < iters; ++it) {
      int4 rn = r;
      if (it + 1 < iters) rn = *reinterpret_cast<const int4*>(my + (size_t)(it + 1) * 128);
      mbar_wait(&empty[st], ph ^ 1);
      if (warp == 0 && lane == 0) mbar_expect_tx(&full[st], kStageBytes);
      __syncwarp();
      if (lane < kOps) gather4(smem_u32(smem + st * kStageBytes + slot * 512), &map, (it % c0_slices) * 64, r.x, r.y, r.z, r.w, &full[st]);
      r = rn;
      if (++st == kStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == P) {
    uint32_t st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (it == 0 && dump && blockIdx.x == 0) {
        for (int i = lane; i < kStageBytes / 16; i += 32) reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<uint4*>(smem)[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == kStages) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int n_rows = 200000, c = 128, iters = 2000, grid = 148;
  std::vector<__nv_bfloat16> h((size_t)n_rows * c);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)((i * 2654435761u >> 16) % 977) * 0.01f);
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)c * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }

  long long* cyc; uint8_t* dump;
  CK(cudaMalloc(&cyc, grid * 8));
  CK(cudaMalloc(&dump, kStageBytes));
  const size_t smem = kStages * kStageBytes + 1024 + 256;

  for (int mode = 0; mode < 4; ++mode) {
    // mode 0: all rows valid, random; 1: 30% valid random; 2: all invalid (-1); 3: 30% valid, locally clustered rows
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
      for (int rep = 0; rep < 2; ++rep) {
        kern<<<grid, 32 * P + 32, smem>>>(map, didx, iters, 2, cyc, rep == 0 ? dump : nullptr);
        CK(cudaDeviceSynchronize());
      }
      std::vector<long long> hc(grid);
      CK(cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto v : hc) avg += (double)v; avg /= grid;
      printf("mode %d, %d issuing warps: %.1f cycles per 16 KB stage per SM (%.1f B/clk/SM incl. zero rows)\n", mode, P, avg / iters, 16384.0 * iters / avg);
    };
    run(bench<1>, 1); run(bench<2>, 2); run(bench<4>, 4); run(bench<16>, 16); run(bench<8>, 8);
    if (mode <= 1) {
      std::vector<uint8_t> hd(kStageBytes);
      CK(cudaMemcpy(hd.data(), dump, kStageBytes, cudaMemcpyDeviceToHost));
      // expected: row j of the tile at j*128, 16-byte chunk ch at ((ch ^ (j & 7)) << 4); block 0, iteration 0, columns 0..63
      int bad = 0;
      for (int j = 0; j < 128 && bad < 5; ++j) {
        const int src = idx[j];
        for (int ch = 0; ch < 8; ++ch) for (int e = 0; e < 8; ++e) {
          const __nv_bfloat16 got = *reinterpret_cast<const __nv_bfloat16*>(&hd[j * 128 + ((ch ^ (j & 7)) << 4) + e * 2]);
          const float want = src >= 0 ? __bfloat162float(h[(size_t)src * c + ch * 8 + e]) : 0.f;
          if (__bfloat162float(got) != want) { if (bad < 5) printf("  mismatch row %d chunk %d elem %d: got %f want %f (src %d)\n", j, ch, e, __bfloat162float(got), want, src); ++bad; }
        }
      }
      printf("  layout check: %s\n", bad ? "MISMATCH" : "ok (swizzled K-major image, zero fill for -1)");
    }
    CK(cudaFree(didx));
  }
  return 0;
}
