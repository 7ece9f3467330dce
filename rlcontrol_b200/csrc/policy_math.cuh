// PolicyNetwork.evaluate (forwardkl_network.py:303-322, reversekl_network.py:325-344) for one row of the policy head
// [mean_raw | log_std_raw]; shared by k_policy_evaluate (kl_agent.cu) and the fused small-batch forward (small_batch.cu).
#pragma once

__device__ __forceinline__ void policy_evaluate_row(const float* __restrict__ head, const float* __restrict__ eps, int A,
                                                    float scale, float lo, float hi, float* __restrict__ action,
                                                    float* __restrict__ logp, float* __restrict__ mean_out,
                                                    float* __restrict__ mu_raw_out, float* __restrict__ log_std_out,
                                                    float* __restrict__ z_out) {
  const float LOG_SQRT_2PI = 0.9189385332046727f;
  float lp = 0.f, corr = 0.f;
  for (int d = 0; d < A; ++d) {
    const float mu = head[d];
    const float ls = fminf(fmaxf(head[A + d], lo), hi);
    const float std = expf(ls);
    const float e = eps ? eps[d] : 0.f;
    float z;
    if (A == 1) {
      z = mu + std * e;
      const float t = (z - mu);
      lp += -(t * t) / (2.f * std * std) - ls - LOG_SQRT_2PI;
    } else {  // MultivariateNormal(mean, diag_embed(std)): std is passed as the covariance (:346-351)
      z = mu + sqrtf(std) * e;
      const float t = (z - mu);
      lp += -0.5f * (t * t) / std - 0.5f * ls - LOG_SQRT_2PI;
    }
    const float a = tanhf(z);
    corr += logf(1.f - a * a + 1e-6f);
    if (action) action[d] = a * scale;
    if (mean_out) mean_out[d] = tanhf(mu) * scale;
    if (mu_raw_out) mu_raw_out[d] = mu;
    if (log_std_out) log_std_out[d] = ls;
    if (z_out) z_out[d] = z;
  }
  if (logp) logp[0] = lp - corr;
}
