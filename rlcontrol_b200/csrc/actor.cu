// Actor side of the Actor-Expert agents, fused with the sampled-action path:
//   * mixture sampling with the draws supplied as tensors (ae_network.py:461-496,
//     ae_actor_network.py:310-341, ac_actor_network.py:262-288),
//   * the whole expert step of ActorExpert.update_network (ActorExpert.py:162-181) in ONE launch:
//     sample N actions per state -> Q(s,a) through the hoisted T-mid critic -> per-state top-k
//     (argsort()[::-1][:k]) -> elite gather; the [B,N,A] action stack and the [B,N] q block are
//     optional outputs only,
//   * the mixture negative log-likelihood of the elites and its gradient wrt the mixture parameters
//     (get_lossfunc, ae_network.py:262-278).
// Sampling arithmetic is float64 like numpy's (loc + scale * N(0,1), then the fp32 cast TF applies at
// the feed), so actions are bit-identical to the reference given the same draws.
#include <math_constants.h>

#include "common.cuh"

#define AE_MAX_A 8
#define AE_MAX_M 8
#define AE_MAX_K 64
#define AE_THREADS 256

// numpy RandomState.choice(M, N, p): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
__device__ __forceinline__ int pick_component(const double* cdf, int M, double u, int equal_modal) {
  if (equal_modal) {
    const int c = (int)(u * (double)M);
    return c < M ? c : M - 1;
  }
  int c = 0;
  while (c < M - 1 && cdf[c] <= u) ++c;
  return c;
}

__device__ __forceinline__ void build_cdf(const float* alpha, int M, double* cdf) {
  double acc = 0.0;
  for (int c = 0; c < M; ++c) {
    acc += (double)alpha[c];
    cdf[c] = acc;
  }
  for (int c = 0; c < M; ++c) cdf[c] /= acc;
}

// action d of sample n of state b
__device__ __forceinline__ float ae_action(int b, int n, int d, int N, int A, int M, int comp,
                                           const float* mean, const float* sigma,
                                           const float* normal, const float* amin,
                                           const float* amax, int n_uniform, const float* uni_u) {
  const double lo = (double)amin[d], hi = (double)amax[d];
  if (n < n_uniform) return (float)(lo + (hi - lo) * (double)uni_u[((long long)b * n_uniform + n) * A + d]);
  const long long pm = ((long long)b * M + comp) * A + d;
  const double v = (double)mean[pm] + (double)sigma[pm] * (double)normal[((long long)b * N + n) * A + d];
  return (float)fmin(fmax(v, lo), hi);
}

__global__ void k_mixture_sample(const float* __restrict__ alpha, const float* __restrict__ mean,
                                 const float* __restrict__ sigma, int B, int M, int A, int N,
                                 int equal_modal, const float* __restrict__ comp_u,
                                 const float* __restrict__ normal, const float* __restrict__ amin,
                                 const float* __restrict__ amax, int n_uniform,
                                 const float* __restrict__ uni_u, float* __restrict__ actions,
                                 int* __restrict__ comp_out) {
  __shared__ double cdf[AE_MAX_M];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) build_cdf(alpha + (long long)b * M, M, cdf);
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int comp = pick_component(cdf, M, (double)comp_u[(long long)b * N + n], equal_modal);
    if (comp_out) comp_out[(long long)b * N + n] = comp;
    for (int d = 0; d < A; ++d)
      actions[((long long)b * N + n) * A + d] =
          ae_action(b, n, d, N, A, M, comp, mean, sigma, normal, amin, amax, n_uniform, uni_u);
  }
}

extern "C" int rlc_mixture_sample(rlc_handle* h, const float* alpha, const float* mean,
                                  const float* sigma, int B, int M, int A, int N, int equal_modal,
                                  const float* comp_u, const float* normal, const float* amin,
                                  const float* amax, int n_uniform, const float* uni_u,
                                  float* actions_out, int32_t* comp_out, void* stream) {
  RLC_REQUIRE(h && mean && sigma && comp_u && normal && amin && amax && actions_out);
  RLC_REQUIRE(equal_modal || alpha);
  RLC_REQUIRE(B >= 0 && N >= 1 && M >= 1 && M <= AE_MAX_M && A >= 1 && A <= 64);
  RLC_REQUIRE(n_uniform >= 0 && n_uniform <= N && (n_uniform == 0 || uni_u));
  if (B == 0) return RLC_OK;
  k_mixture_sample<<<B, 128, 0, (cudaStream_t)stream>>>(alpha ? alpha : mean, mean, sigma, B, M, A, N,
                                                        equal_modal, comp_u, normal, amin, amax,
                                                        n_uniform, uni_u, actions_out, comp_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused expert step: one CTA per state. p = hoisted state term of the T-mid critic ([B,H2]).
// ---------------------------------------------------------------------------------------------
template <int AT>
__global__ void __launch_bounds__(AE_THREADS)
k_ae_expert(const float* __restrict__ p, int B, int N, int A, int H2, int k, int M,
            const float* __restrict__ W2a, const float* __restrict__ w3,
            const float* __restrict__ b3, const float* __restrict__ alpha,
            const float* __restrict__ mean, const float* __restrict__ sigma, int equal_modal,
            const float* __restrict__ comp_u, const float* __restrict__ normal,
            const float* __restrict__ amin, const float* __restrict__ amax, int n_uniform,
            const float* __restrict__ uni_u, float* __restrict__ actions_out,
            float* __restrict__ q_out, long long* __restrict__ idx_out,
            float* __restrict__ q_sel_out, float* __restrict__ elites_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int H2P = (H2 + 3) & ~3;
  float* ps = reinterpret_cast<float*>(smraw);  // [H2P]
  float* w3s = ps + H2P;                        // [H2P]
  float* was = w3s + H2P;                       // [AT][H2P]
  float* qbuf = was + AT * H2P;                 // [N]
  float* abuf = qbuf + N;                       // [N][AT]
  __shared__ double cdf[AE_MAX_M];
  __shared__ float red_v[AE_THREADS / 32];
  __shared__ int red_i[AE_THREADS / 32];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < H2P; i += AE_THREADS) {
    ps[i] = (i < H2) ? p[(long long)b * H2 + i] : 0.f;
    w3s[i] = (i < H2) ? w3[i] : 0.f;
  }
  for (int i = tid; i < AT * H2P; i += AE_THREADS) {
    const int ai = i / H2P, j = i - ai * H2P;
    was[i] = (ai < A && j < H2) ? W2a[(long long)ai * H2 + j] : 0.f;
  }
  if (tid == 0) build_cdf(alpha + (long long)b * M, M, cdf);
  __syncthreads();
  const float bb3 = b3[0];

  // ---- sample + evaluate ----
  for (int n = tid; n < N; n += AE_THREADS) {
    const int comp = pick_component(cdf, M, (double)comp_u[(long long)b * N + n], equal_modal);
    float ar[AT];
#pragma unroll
    for (int i = 0; i < AT; ++i) {
      ar[i] = (i < A) ? ae_action(b, n, i, N, A, M, comp, mean, sigma, normal, amin, amax, n_uniform, uni_u)
                      : 0.f;
      abuf[n * AT + i] = ar[i];
      if (actions_out && i < A) actions_out[((long long)b * N + n) * A + i] = ar[i];
    }
    float q = 0.f;
    for (int j = 0; j < H2P; j += 4) {
      float4 z = *reinterpret_cast<const float4*>(ps + j);
#pragma unroll
      for (int i = 0; i < AT; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(was + i * H2P + j);
        z.x = fmaf(ar[i], w.x, z.x);
        z.y = fmaf(ar[i], w.y, z.y);
        z.z = fmaf(ar[i], w.z, z.z);
        z.w = fmaf(ar[i], w.w, z.w);
      }
      const float4 w3v = *reinterpret_cast<const float4*>(w3s + j);
      q = fmaf(w3v.x, fmaxf(z.x, 0.f), q);
      q = fmaf(w3v.y, fmaxf(z.y, 0.f), q);
      q = fmaf(w3v.z, fmaxf(z.z, 0.f), q);
      q = fmaf(w3v.w, fmaxf(z.w, 0.f), q);
    }
    q += bb3;
    qbuf[n] = q;
    if (q_out) q_out[(long long)b * N + n] = q;
  }
  __syncthreads();
  // ---- top-k tournament: descending, ties -> larger index first (reversed stable ascending sort) ----
  float lastv = CUDART_INF_F;
  int lasti = 0x7fffffff;
  for (int t = 0; t < k; ++t) {
    float bv = -CUDART_INF_F;
    int bi = -1;
    for (int n = tid; n < N; n += AE_THREADS) {
      float v = qbuf[n];
      if (v != v) v = CUDART_INF_F;
      const bool below = (v < lastv) || (v == lastv && n < lasti);
      if (below && (bi < 0 || v > bv || (v == bv && n > bi))) { bv = v; bi = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < AE_THREADS / 32; ++w) {
        const float ov = red_v[w];
        const int oi = red_i[w];
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
      }
      red_v[0] = bv;
      red_i[0] = bi;
      idx_out[(long long)b * k + t] = bi;
      if (q_sel_out) q_sel_out[(long long)b * k + t] = (bi >= 0) ? qbuf[bi] : 0.f;
    }
    __syncthreads();
    lastv = red_v[0];
    lasti = red_i[0];
    if (elites_out && tid < A && lasti >= 0)
      elites_out[((long long)b * k + t) * A + tid] = abuf[lasti * AT + tid];
    __syncthreads();
  }
}

extern "C" int rlc_ae_expert_step(rlc_handle* h, const rlc_critic* c, const float* s, int B, int N,
                                  int k, const float* alpha, const float* mean, const float* sigma,
                                  int M, int equal_modal, const float* comp_u, const float* normal,
                                  const float* amin, const float* amax, int n_uniform,
                                  const float* uni_u, float* actions_out, float* q_out,
                                  int64_t* idx_out, float* q_sel_out, float* elites_out,
                                  void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && mean && sigma && comp_u && normal && amin && amax && idx_out);
  RLC_REQUIRE(equal_modal || alpha);
  RLC_REQUIRE(B >= 0 && N >= 1 && k >= 1 && k <= AE_MAX_K && k <= N && M >= 1 && M <= AE_MAX_M);
  RLC_REQUIRE(n_uniform >= 0 && n_uniform <= N && (n_uniform == 0 || uni_u));
  if (c->topology != RLC_TMID || c->A > AE_MAX_A) return RLC_ERR_UNSUPPORTED;
  if (B == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  void* ws = nullptr;
  int rc = rlc_workspace(h, (size_t)B * c->H2 * sizeof(float), &ws);
  if (rc) return rc;
  rc = rlc_tmid_state_term(h, c, s, B, (float*)ws, st);
  if (rc) return rc;
  const ThetaView t = theta_view(RLC_TMID, c->S, c->A, c->H1, c->H2);
  const float* W2a = c->theta + t.oW2 + (int64_t)c->H1 * c->H2;
  const int H2P = (c->H2 + 3) & ~3;
#define RLC_AE_CASE(AT)                                                                          \
  {                                                                                              \
    const size_t smem = ((size_t)H2P * (2 + AT) + (size_t)N * (1 + AT)) * sizeof(float);         \
    if (smem + 4096 > h->smem_optin) return RLC_ERR_UNSUPPORTED;                                 \
    auto kern = k_ae_expert<AT>;                                                                 \
    RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<B, AE_THREADS, smem, st>>>((const float*)ws, B, N, c->A, c->H2, k, M, W2a,            \
                                      c->theta + t.ow3, c->theta + t.ob3, alpha ? alpha : mean,  \
                                      mean, sigma, equal_modal, comp_u, normal, amin, amax,      \
                                      n_uniform, uni_u, actions_out, q_out,                      \
                                      (long long*)idx_out, q_sel_out, elites_out);               \
  }
  if (c->A <= 1) RLC_AE_CASE(1)
  else if (c->A <= 2) RLC_AE_CASE(2)
  else if (c->A <= 4) RLC_AE_CASE(4)
  else RLC_AE_CASE(8)
#undef RLC_AE_CASE
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Mixture NLL of the elites and its gradient (get_lossfunc, ae_network.py:262-278; tf_normal :230-243).
// One warp per state; lanes stride over the k elites.  Pass 1 caches every elite's component densities
// and dLoss/dmix in shared memory, pass 2 reduces one mixture parameter at a time (fixed order:
// deterministic).  fp32 like the TF graph, including clip_by_value(mix, 1e-30, 1e30) (no gradient
// outside the clip range).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_mixture_nll(const float* __restrict__ alpha, const float* __restrict__ mean,
              const float* __restrict__ sigma, const float* __restrict__ y, int B, int M, int A, int k,
              int equal_modal, float inv_rows, float* __restrict__ loss, float* __restrict__ nll_out,
              float* __restrict__ dalpha, float* __restrict__ dmean, float* __restrict__ dsigma) {
  extern __shared__ float smf[];
  float* gE = smf;          // [k]     dLoss/dmix of elite e
  float* densE = smf + k;   // [k][M]  component densities
  const int b = blockIdx.x, lane = threadIdx.x;
  const float TWO_PI = 6.283185307179586f;
  const float* mu = mean + (long long)b * M * A;
  const float* sg = sigma + (long long)b * M * A;
  float lsum = 0.f;
  for (int e = lane; e < k; e += 32) {
    const float* ye = y + ((long long)b * k + e) * A;
    float mix = 0.f;
    for (int c = 0; c < M; ++c) {
      float pr = 1.f;
      for (int d = 0; d < A; ++d) {
        const float s_ = sg[c * A + d], t = ye[d] - mu[c * A + d];
        pr *= sqrtf(1.f / (TWO_PI * s_ * s_)) * expf(-(t * t) / (2.f * s_ * s_));
      }
      densE[e * M + c] = pr;
      mix += (equal_modal ? 1.f / (float)M : alpha[(long long)b * M + c]) * pr;
    }
    const bool inside = (mix >= 1e-30f) && (mix <= 1e30f);
    const float nll = -logf(fminf(fmaxf(mix, 1e-30f), 1e30f));
    lsum += nll;
    if (nll_out) nll_out[(long long)b * k + e] = nll;
    gE[e] = inside ? -inv_rows / mix : 0.f;
  }
  lsum = warp_sum(lsum);
  if (lane == 0 && loss) atomicAdd(loss, lsum * inv_rows);
  __syncwarp();
  for (int c = 0; c < M; ++c) {
    const float w = equal_modal ? 1.f / (float)M : alpha[(long long)b * M + c];
    if (dalpha) {
      float a = 0.f;
      if (!equal_modal)
        for (int e = lane; e < k; e += 32) a += gE[e] * densE[e * M + c];
      a = warp_sum(a);
      if (lane == 0) dalpha[(long long)b * M + c] = a;
    }
    for (int d = 0; d < A; ++d) {
      const float s_ = sg[c * A + d], m_ = mu[c * A + d];
      float am = 0.f, as = 0.f;
      for (int e = lane; e < k; e += 32) {
        const float t = y[((long long)b * k + e) * A + d] - m_;
        const float gd = gE[e] * w * densE[e * M + c];
        am += gd * t / (s_ * s_);
        as += gd * (t * t / (s_ * s_ * s_) - 1.f / s_);
      }
      am = warp_sum(am);
      as = warp_sum(as);
      if (lane == 0) {
        if (dmean) dmean[((long long)b * M + c) * A + d] = am;
        if (dsigma) dsigma[((long long)b * M + c) * A + d] = as;
      }
    }
  }
}

extern "C" int rlc_mixture_nll(rlc_handle* h, const float* alpha, const float* mean,
                               const float* sigma, const float* actions, int B, int M, int A, int k,
                               int equal_modal, int B_total, float* loss_out, float* nll_out,
                               float* dalpha_out, float* dmean_out, float* dsigma_out,
                               void* stream) {
  RLC_REQUIRE(h && mean && sigma && actions && (equal_modal || alpha));
  RLC_REQUIRE(B >= 0 && M >= 1 && M <= AE_MAX_M && A >= 1 && A <= 64 && k >= 1 && B_total >= B);
  const size_t smem = (size_t)k * (1 + M) * sizeof(float);
  RLC_REQUIRE(smem <= 40 * 1024);
  cudaStream_t st = (cudaStream_t)stream;
  if (loss_out) RLC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (B == 0) return RLC_OK;
  k_mixture_nll<<<B, 32, smem, st>>>(alpha ? alpha : mean, mean, sigma, actions, B, M, A, k, equal_modal,
                                     1.f / ((float)B_total * (float)k), loss_out, nll_out, dalpha_out,
                                     dmean_out, dsigma_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
