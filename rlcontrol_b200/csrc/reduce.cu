// Per-state reductions over the N sampled actions (K3): one warp per state, warp-shuffle
// reductions, q rows streamed with coalesced loads.  HBM-bound: algorithmic bytes = 4*B*N read
// (+4*B*N per extra [B,N] operand) and O(B*k) written.
#include <math_constants.h>

#include "common.cuh"

#define WARPS_PER_BLOCK 4

__device__ __forceinline__ bool key_gt(float va, int ia, float vb, int ib) {
  return (va > vb) || (va == vb && ia > ib);
}

// row.argsort()[::-1][:k]: k rounds of "largest key strictly below the previous pick".
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_topk(const float* __restrict__ q, int B, int N, int k, long long* __restrict__ idx_out,
       float* __restrict__ q_sel_out, const float* __restrict__ actions, int A, int act_per_state,
       float* __restrict__ elites_out) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  float lastv = CUDART_INF_F;
  int lasti = 0x7fffffff;
  for (int t = 0; t < k; ++t) {
    float bv = -CUDART_INF_F;
    int bi = -1;
    for (int n = lane; n < N; n += 32) {
      float v = row[n];
      if (v != v) v = CUDART_INF_F;  // NaN ordered like +inf (numpy sorts NaN last ascending)
      const bool below = (v < lastv) || (v == lastv && n < lasti);
      if (below && (bi < 0 || key_gt(v, n, bv, bi))) { bv = v; bi = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || key_gt(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    lastv = bv;
    lasti = bi;
    if (lane == 0) {
      idx_out[(long long)b * k + t] = bi;
      if (q_sel_out) q_sel_out[(long long)b * k + t] = (bi >= 0) ? row[bi] : 0.f;
    }
    if (elites_out && bi >= 0) {
      const float* src = actions + ((act_per_state ? (long long)b * N : 0) + bi) * A;
      float* dst = elites_out + ((long long)b * k + t) * A;
      for (int i = lane; i < A; i += 32) dst[i] = src[i];
    }
  }
}

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_stats(const float* __restrict__ q, int B, int N, long long* __restrict__ argmax_out,
        float* __restrict__ max_out, float* __restrict__ mean_out) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  float bv = -CUDART_INF_F, sum = 0.f;
  int bi = 0x7fffffff;
  for (int n = lane; n < N; n += 32) {
    const float v = row[n];
    sum += v;
    if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    if (argmax_out) argmax_out[b] = bi;
    if (max_out) max_out[b] = bv;
    if (mean_out) mean_out[b] = sum / (float)N;
  }
}

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_lse(const float* __restrict__ q, int B, int N, float offset, float* __restrict__ v_out) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  float m = -CUDART_INF_F;
  for (int n = lane; n < N; n += 32) m = fmaxf(m, row[n]);
  m = warp_max(m);
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += expf(row[n] - m);
  s = warp_sum(s);
  if (lane == 0) v_out[b] = m + logf(s) + offset;
}

// ForwardKL (forwardkl_network.py:165-194)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_fkl(const float* __restrict__ q, const float* __restrict__ w, const float* __restrict__ logp,
      int B, int N, float alpha, float inv_btotal, float* __restrict__ loss_b,
      float* __restrict__ boltz, float* __restrict__ dlogp) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  const float* lp = logp + (long long)b * N;
  float m = -CUDART_INF_F;
  for (int n = lane; n < N; n += 32) m = fmaxf(m, __fdiv_rn(row[n], alpha));
  m = warp_max(m);
  float z = 0.f;
  for (int n = lane; n < N; n += 32) z = fmaf(expf(__fdiv_rn(row[n], alpha) - m), w[n], z);
  z = warp_sum(z);
  float acc = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float p = __fdiv_rn(expf(__fdiv_rn(row[n], alpha) - m), z);
    const float pw = p * w[n];
    acc = fmaf(pw, lp[n], acc);
    if (boltz) boltz[(long long)b * N + n] = p;
    if (dlogp) dlogp[(long long)b * N + n] = -pw * inv_btotal;
  }
  acc = warp_sum(acc);
  if (lane == 0) loss_b[b] = -acc;
}

// ReverseKL (reversekl_network.py:181-190 ; hard 197-203: alpha = 0)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_rkl(const float* __restrict__ q, const float* __restrict__ v, const float* __restrict__ w,
      const float* __restrict__ logp, int B, int N, float alpha, float inv_btotal,
      float* __restrict__ loss_b, float* __restrict__ dlogp) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  const float* lp = logp + (long long)b * N;
  const float vb = v[b];
  float acc = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float l = lp[n];
    const float pe = expf(l);
    const float adv = row[n] - vb;
    const float inner = adv - alpha * l;
    acc = fmaf(-pe * inner, w[n], acc);
    if (dlogp) dlogp[(long long)b * N + n] = (-pe * (inner - alpha)) * w[n] * inv_btotal;
  }
  acc = warp_sum(acc);
  if (lane == 0) loss_b[b] = acc;
}

static inline unsigned nblocks(int B) { return (unsigned)((B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

extern "C" int rlc_reduce_topk(rlc_handle* h, const float* q, int B, int N, int k,
                               int64_t* idx_out, float* q_sel_out, const float* actions, int A,
                               int act_mode, float* elites_out, void* stream) {
  RLC_REQUIRE(h && q && idx_out && B >= 0 && N >= 1 && k >= 1 && k <= 64 && k <= N);
  RLC_REQUIRE(!elites_out || (actions && A >= 1));
  if (B == 0) return RLC_OK;
  k_topk<<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, B, N, k, (long long*)idx_out, q_sel_out, actions, A, act_mode == RLC_ACT_PER_STATE,
      elites_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_reduce_stats(rlc_handle* h, const float* q, int B, int N, int64_t* argmax_out,
                                float* max_out, float* mean_out, void* stream) {
  RLC_REQUIRE(h && q && B >= 0 && N >= 1);
  if (B == 0) return RLC_OK;
  k_stats<<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, B, N, (long long*)argmax_out, max_out, mean_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_reduce_lse(rlc_handle* h, const float* q, int B, int N, int action_dim,
                              float* v_out, void* stream) {
  RLC_REQUIRE(h && q && v_out && B >= 0 && N >= 1);
  if (B == 0) return RLC_OK;
  const float offset = -logf((float)N) + (float)(action_dim * 0.6931471805599453);
  k_lse<<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(q, B, N, offset, v_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_reduce_fkl(rlc_handle* h, const float* q, const float* w, const float* logp,
                              int B, int N, float entropy_scale, int B_total, float* loss_b_out,
                              float* boltz_out, float* dlogp_out, void* stream) {
  RLC_REQUIRE(h && q && w && logp && loss_b_out && B >= 0 && N >= 1 && B_total >= B && B_total >= 1);
  RLC_REQUIRE(entropy_scale > 0.f);
  if (B == 0) return RLC_OK;
  k_fkl<<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, w, logp, B, N, entropy_scale, 1.f / (float)B_total, loss_b_out, boltz_out, dlogp_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_reduce_rkl(rlc_handle* h, const float* q, const float* v, const float* w,
                              const float* logp, int B, int N, float entropy_scale, int hard,
                              int B_total, float* loss_b_out, float* dlogp_out, void* stream) {
  RLC_REQUIRE(h && q && v && w && logp && loss_b_out && B >= 0 && N >= 1 && B_total >= B && B_total >= 1);
  if (B == 0) return RLC_OK;
  k_rkl<<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, v, w, logp, B, N, hard ? 0.f : entropy_scale, 1.f / (float)B_total, loss_b_out,
      dlogp_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
