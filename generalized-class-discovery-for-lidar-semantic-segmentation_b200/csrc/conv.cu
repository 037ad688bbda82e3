// conv.cu — C-ABI entry points of the sparse convolution: argument validation and dispatch
// between the fp32 SIMT kernels (conv_simt.cu) and the bf16 tcgen05 kernels (conv_tc.cu).
#include "common.cuh"

namespace gcd {
int32_t conv_forward_simt(const gcd_conv_args* a, cudaStream_t st);
int32_t conv_wgrad_simt(const gcd_wgrad_args* a, cudaStream_t st);
int32_t colsum_f32(const void* x, int64_t ld, int64_t n, int c, int dtype, float* out, cudaStream_t st);
int32_t conv_forward_tc(const gcd_conv_args* a, cudaStream_t st);
int32_t conv_wgrad_tc(const gcd_wgrad_args* a, cudaStream_t st);
bool conv_forward_tc_supported(const gcd_conv_args* a);
bool conv_wgrad_tc_supported(const gcd_wgrad_args* a);
}  // namespace gcd

using namespace gcd;

extern "C" int32_t gcd_conv_forward(const gcd_conv_args* a, void* stream) {
  GCD_REQUIRE(a != nullptr, "gcd_conv_forward: null args");
  GCD_REQUIRE(a->kv >= 1 && a->kv <= 125, "gcd_conv_forward: kernel volume %d out of range", a->kv);
  GCD_REQUIRE(a->c_in >= 1 && a->c_out >= 1, "gcd_conv_forward: bad channel counts");
  GCD_REQUIRE(a->n_out >= 0 && a->n_out < (1ll << 31) && a->n_in >= 0 && a->n_in < (1ll << 31), "gcd_conv_forward: row counts out of the int32 index range");
  GCD_REQUIRE(a->nbr != nullptr || (a->kv == 1 && a->n_in == a->n_out), "gcd_conv_forward: identity map needs kv == 1 and n_in == n_out");
  GCD_REQUIRE(a->ld_in >= a->c_in && a->ld_out >= a->c_out, "gcd_conv_forward: leading dimension smaller than channel count");
  GCD_REQUIRE(a->in && a->out, "gcd_conv_forward: null feature pointer");
  if (a->n_out == 0) return GCD_OK;
  cudaStream_t st = as_stream(stream);
  if (a->math_mode == GCD_MATH_BF16_TCGEN05) {
    if (!conv_forward_tc_supported(a)) { set_error("gcd_conv_forward: shape/dtype not supported by the tcgen05 path (c_in=%d c_out=%d)", a->c_in, a->c_out); return GCD_ERR_UNSUPPORTED; }
    const int32_t rc = conv_forward_tc(a, st);
    if (rc != GCD_OK || a->stats == nullptr) return rc;
    // the tcgen05 kernel has no statistics epilogue (measured: it makes the epilogue warps the pacing role on the 1x1 and
    // 2x2x2 layers); honour the argument with the streaming pass over the result
    return gcd_bn_stats(a->out, a->ld_out, a->n_out, a->c_out, a->out_dtype, a->stats, stream);
  }
  GCD_REQUIRE(a->w != nullptr, "gcd_conv_forward: fp32 weights required for the SIMT path");
  GCD_REQUIRE(a->out_rows == nullptr, "gcd_conv_forward: tile-sorted tables (out_rows) are a tcgen05-path feature");
  return conv_forward_simt(a, st);
}

extern "C" int32_t gcd_conv_wgrad(const gcd_wgrad_args* a, void* stream) {
  GCD_REQUIRE(a != nullptr, "gcd_conv_wgrad: null args");
  GCD_REQUIRE(a->kv >= 1 && a->kv <= 125, "gcd_conv_wgrad: kernel volume %d out of range", a->kv);
  GCD_REQUIRE((a->pair_in && a->pair_out && a->pair_off) || (!a->pair_in && !a->pair_out && !a->pair_off && a->kv == 1),
              "gcd_conv_wgrad: pair lists must be all given, or all NULL with kv == 1");
  GCD_REQUIRE(a->in && a->gout && a->dw, "gcd_conv_wgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  if (a->dbias) {
    int32_t rc = colsum_f32(a->gout, a->ld_gout, a->n_out, a->c_out, a->gout_dtype, a->dbias, st);
    if (rc != GCD_OK) return rc;
  }
  if (a->n_pairs == 0) return GCD_OK;
  if (a->math_mode == GCD_MATH_BF16_TCGEN05) {
    if (!conv_wgrad_tc_supported(a)) { set_error("gcd_conv_wgrad: shape/dtype not supported by the tcgen05 path (c_in=%d c_out=%d)", a->c_in, a->c_out); return GCD_ERR_UNSUPPORTED; }
    return conv_wgrad_tc(a, st);
  }
  return conv_wgrad_simt(a, st);
}
