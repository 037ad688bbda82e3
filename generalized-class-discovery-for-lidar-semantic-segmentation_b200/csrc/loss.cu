// loss.cu — fused teacher/student consistency terms on voxel logits (loss_rows.cuh).  SURVEY 8(f) rank 3.
#include "common.cuh"
#include "loss_rows.cuh"

namespace gcd {
namespace {
__global__ void __launch_bounds__(256) consistency_rows_kernel(const float* __restrict__ logits_s, int64_t ld_s, const float* __restrict__ logits_t,
                                                               int64_t ld_t, int64_t n, int c, float threshold, float* __restrict__ sq_err,
                                                               float* __restrict__ max_prob, int64_t* __restrict__ label,
                                                               float* __restrict__ grad, int64_t ld_g) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) consistency_row_thread(i, logits_s, ld_s, logits_t, ld_t, c, threshold, sq_err, max_prob, label, grad, ld_g);
}
}  // namespace
}  // namespace gcd

using namespace gcd;

extern "C" int32_t gcd_consistency_rows(const float* logits_s, int64_t ld_s, const float* logits_t, int64_t ld_t, int64_t n, int32_t c,
                                        float threshold, float* sq_err, float* max_prob, int64_t* label, float* grad, int64_t ld_g,
                                        void* stream) {
  GCD_REQUIRE(n >= 0 && c >= 1 && c <= 4096 && ld_s >= c && ld_t >= c, "gcd_consistency_rows: bad shape n=%lld c=%d", (long long)n, c);
  GCD_REQUIRE(grad == nullptr || ld_g >= c, "gcd_consistency_rows: gradient leading dimension smaller than the class count");
  if (n == 0) return GCD_OK;
  GCD_REQUIRE(logits_s && logits_t && sq_err && max_prob && label, "gcd_consistency_rows: null pointer");
  consistency_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(logits_s, ld_s, logits_t, ld_t, n, c, threshold, sq_err,
                                                                                      max_prob, label, grad, ld_g);
  GCD_LAUNCH_CHECK("gcd_consistency_rows");
  return GCD_OK;
}
