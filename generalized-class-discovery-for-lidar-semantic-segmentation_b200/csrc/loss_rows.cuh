// loss_rows.cuh — per-row body of the teacher/student consistency terms of the Stage-2 step on voxel logits
// (ref modules/exp_merge_mean_teacher.py:2832-2850: two softmaxes, F.mse_loss of the probabilities, torch.max of the
// teacher's): one pass over the two logit rows instead of six elementwise kernels forward and as many backward.
//   ps = softmax(logits_s[i]), pt = softmax(logits_t[i])
//   sq_err[i]   = sum_c (ps_c - pt_c)^2                 -> F.mse_loss(ps, pt) = sum_i sq_err[i] / (n c)
//   max_prob[i] = max_c pt_c, label[i] = first argmax   -> torch.max(prob_t, dim=1); label = -1 where max_prob < threshold
//   grad[i, c]  = d sq_err[i] / d logits_s[i, c] = 2 ps_c ((ps_c - pt_c) - sum_j ps_j (ps_j - pt_j))     (optional)
// No CUDA headers: tests/emu/ compiles the same source with g++ (expf differs in the last ulps between host and device;
// both sides are held to the torch reference with a stated tolerance).
#pragma once
#include <math.h>
#include <stdint.h>
#include "keys.cuh"   // GCD_DEVFN

namespace gcd {

GCD_DEVFN void consistency_row_thread(int64_t i, const float* logits_s, int64_t ld_s, const float* logits_t, int64_t ld_t, int c,
                                      float threshold, float* sq_err, float* max_prob, int64_t* label, float* grad, int64_t ld_g) {
  const float* s = logits_s + i * ld_s;
  const float* t = logits_t + i * ld_t;
  float ms = s[0], mt = t[0];
  int arg = 0;
  for (int j = 1; j < c; ++j) {
    ms = fmaxf(ms, s[j]);
    if (t[j] > mt) { mt = t[j]; arg = j; }        // strict: the first maximum wins, as torch.max does
  }
  float zs = 0.f, zt = 0.f;
  for (int j = 0; j < c; ++j) { zs += expf(s[j] - ms); zt += expf(t[j] - mt); }
  const float is = 1.f / zs, it = 1.f / zt;
  float err = 0.f, dot = 0.f;                      // sum_c d_c^2 and sum_c ps_c d_c, d = ps - pt
  for (int j = 0; j < c; ++j) {
    const float ps = expf(s[j] - ms) * is, d = ps - expf(t[j] - mt) * it;
    err += d * d;
    dot += ps * d;
  }
  sq_err[i] = err;
  const float mp = it;                             // exp(mt - mt) / zt
  max_prob[i] = mp;
  label[i] = (threshold > 0.f && mp < threshold) ? -1 : (int64_t)arg;
  if (grad) {
    float* g = grad + i * ld_g;
    for (int j = 0; j < c; ++j) {
      const float ps = expf(s[j] - ms) * is, d = ps - expf(t[j] - mt) * it;
      g[j] = 2.f * ps * (d - dot);
    }
  }
}

}  // namespace gcd
