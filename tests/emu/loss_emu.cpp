// loss_emu.cpp — host emulation of consistency_rows_kernel: csrc/loss_rows.cuh compiled unchanged by g++, one call per row.
#include "loss_rows.cuh"

extern "C" void emu_consistency_rows(const float* logits_s, int64_t ld_s, const float* logits_t, int64_t ld_t, int64_t n, int32_t c,
                                     float threshold, float* sq_err, float* max_prob, int64_t* label, float* grad, int64_t ld_g) {
  for (int64_t i = 0; i < n; ++i) gcd::consistency_row_thread(i, logits_s, ld_s, logits_t, ld_t, c, threshold, sq_err, max_prob, label, grad, ld_g);
}
