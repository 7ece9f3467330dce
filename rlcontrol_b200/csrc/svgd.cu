// Soft-Q-learning actor step: the Stein variational gradient of sql_network.py:96-117 with
// the adaptive RBF kernel of utils/sql_kernel.py:7-69, on top of the critic's dQ/da.
//   grad_log_p[b,i,:] = dQ/da(s_b, fixed_i) + d/da sum_d log(1 - a_d^2 + EPS)           (:101-105)
//   dist[b,i,j]       = |fixed_i - updated_j|^2 ;  h_b = max(median_b / log(Kf), h_min)  (sql_kernel.py:33-55)
//       median = the (Kf*Ku//2 + 1)-th largest of the Kf*Ku distances (tf.nn.top_k(...)[-1])
//   kappa = exp(-dist / h) ; dkappa/dfixed = -2 (fixed_i - updated_j) / h * kappa        (:57-67)
//   action_gradients[b,j,:] = mean_i( kappa[b,i,j] grad_log_p[b,i,:] + dkappa[b,i,j,:] ) (:113-114)
// One CTA per state; the Kf*Ku distances are sorted in shared memory (bitonic) for the median.
#include <math_constants.h>

#include "common.cuh"

#define SVGD_THREADS 256
#define SVGD_MAX_PAIRS 4096
#define SVGD_MAX_A 16

__global__ void __launch_bounds__(SVGD_THREADS)
k_svgd(const float* __restrict__ fixed, const float* __restrict__ updated,
       const float* __restrict__ dqda, int B, int Kf, int Ku, int A, float h_min, float eps,
       float* __restrict__ grad_out, float* __restrict__ kappa_out, float* __restrict__ h_out) {
  extern __shared__ float sm[];
  const int P = Kf * Ku;
  int P2 = 1;
  while (P2 < P) P2 <<= 1;
  float* dist = sm;          // [P]   i-major: i * Ku + j
  float* srt = dist + P;     // [P2]  sort buffer (descending)
  float* glp = srt + P2;     // [Kf][A]  grad log p
  __shared__ float h_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* fx = fixed + (long long)b * Kf * A;
  const float* up = updated + (long long)b * Ku * A;
  for (int e = tid; e < P; e += SVGD_THREADS) {
    const int i = e / Ku, j = e - i * Ku;
    float d2 = 0.f;
    for (int d = 0; d < A; ++d) {
      const float t = fx[i * A + d] - up[j * A + d];
      d2 = __fadd_rn(d2, __fmul_rn(t, t));   // no FMA contraction: diff ** 2 then reduce_sum, as the graph does
    }
    dist[e] = d2;
    srt[e] = d2;
  }
  for (int e = P + tid; e < P2; e += SVGD_THREADS) srt[e] = -CUDART_INF_F;
  for (int e = tid; e < Kf * A; e += SVGD_THREADS) {
    const float a = fx[e];
    // d/da log(1 - a^2 + eps) = -2a / (1 - a^2 + eps)
    glp[e] = dqda[(long long)b * Kf * A + e] + (-2.f * a) / (1.f - a * a + eps);
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < P2; e += SVGD_THREADS) {
        const int x = e ^ j;
        if (x > e) {
          const bool desc = (e & k) == 0;
          const float a = srt[e], c = srt[x];
          if (desc ? (a < c) : (a > c)) { srt[e] = c; srt[x] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    const float med = srt[P / 2];                 // top_k(k = P//2 + 1)[-1]
    h_s = fmaxf(med / logf((float)Kf), h_min);
    if (h_out) h_out[b] = h_s;
  }
  __syncthreads();
  const float h = h_s;
  for (int e = tid; e < P; e += SVGD_THREADS) {
    const float kap = expf(-dist[e] / h);
    dist[e] = kap;
    if (kappa_out) kappa_out[(long long)b * P + e] = kap;
  }
  __syncthreads();
  for (int e = tid; e < Ku * A; e += SVGD_THREADS) {
    const int j = e / A, d = e - j * A;
    float acc = 0.f;
    for (int i = 0; i < Kf; ++i) {
      const float kap = dist[i * Ku + j];
      const float diff = fx[i * A + d] - up[j * A + d];
      acc += kap * glp[i * A + d] + (-2.f * diff / h) * kap;
    }
    grad_out[((long long)b * Ku + j) * A + d] = acc / (float)Kf;
  }
}

extern "C" int rlc_svgd_action_grads(rlc_handle* h, const rlc_critic* c, const float* s, int B,
                                     const float* fixed, int Kf, const float* updated, int Ku,
                                     float h_min, float eps, float* dqda_scratch, float* grad_out,
                                     float* q_fixed_out, float* kappa_out, float* h_out,
                                     void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && fixed && updated && dqda_scratch && grad_out);
  RLC_REQUIRE(B >= 0 && Kf >= 2 && Ku >= 1 && (long long)Kf * Ku <= SVGD_MAX_PAIRS && c->A <= SVGD_MAX_A);
  RLC_REQUIRE(h_min > 0.f);
  if (B == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // dQ/da on the B*Kf (state, fixed particle) rows without materialising the stacked states
  int rc = rlc_critic_grad_action_rep(h, c, s, Kf, fixed, (long long)B * Kf, dqda_scratch, q_fixed_out, st);
  if (rc) return rc;
  const int P = Kf * Ku;
  int P2 = 1;
  while (P2 < P) P2 <<= 1;
  const size_t smem = ((size_t)P + P2 + (size_t)Kf * c->A) * sizeof(float);
  k_svgd<<<B, SVGD_THREADS, smem, st>>>(fixed, updated, dqda_scratch, B, Kf, Ku, c->A, h_min, eps, grad_out,
                                        kappa_out, h_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
