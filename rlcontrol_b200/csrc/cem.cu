// K4: the whole QT-Opt CEM (qt_opt_network.py:132-175) in one launch: one CTA per state keeps
// the hoisted state term p_b, W2a and w3 in shared memory, and for every iteration samples the
// N actions (random draws supplied as tensors), evaluates Q, selects the top-m by a block-wide
// tournament and refits the bounded diagonal mixture (utils/boundedvar_gaussian_mixture.py) in
// shared memory.  Mixture arithmetic is fp64 (as sklearn), Q arithmetic fp32 (as TF).
#include <math_constants.h>

#include "common.cuh"

#define CEM_MAX_A 8
#define CEM_MAX_K 32
#define CEM_THREADS 256

struct Gmm {
  double w[2];
  double mu[2][CEM_MAX_A];
  double var[2][CEM_MAX_A];
};

__device__ __forceinline__ double clampd(double x, double lo, double hi) {
  return fmin(fmax(x, lo), hi);
}

// sklearn _estimate_gaussian_parameters('diag') + the two np.clip lines + weights = nk/n.
__device__ void gmm_mstep(const double* X, int k, int A, int M, const double* resp /*[k][2]*/,
                          Gmm& g) {
  const double VAR_LO = 0.1353352832366127, VAR_HI = 7.38905609893065;  // exp(-2), exp(2)
  for (int c = 0; c < M; ++c) {
    double nk = 0.0;
    for (int i = 0; i < k; ++i) nk += resp[i * 2 + c];
    nk += 10.0 * 2.220446049250313e-16;
    for (int d = 0; d < A; ++d) {
      double sx = 0.0, sx2 = 0.0;
      for (int i = 0; i < k; ++i) {
        const double x = X[i * CEM_MAX_A + d], r = resp[i * 2 + c];
        sx += r * x;
        sx2 += r * x * x;
      }
      const double mean = sx / nk;
      const double cov = sx2 / nk - 2.0 * (mean * sx / nk) + mean * mean + 1e-6;
      g.mu[c][d] = clampd(mean, -2.0, 2.0);
      g.var[c][d] = clampd(cov, VAR_LO, VAR_HI);
    }
    g.w[c] = nk / (double)k;
  }
}

// sklearn E-step (diag). Writes resp = exp(log_resp); returns mean log-likelihood.
__device__ double gmm_estep(const double* X, int k, int A, int M, const Gmm& g, double* resp) {
  double ll = 0.0;
  double logdet[2], logw[2];
  for (int c = 0; c < M; ++c) {
    double ld = 0.0;
    for (int d = 0; d < A; ++d) ld += log(1.0 / sqrt(g.var[c][d]));
    logdet[c] = ld;
    logw[c] = log(g.w[c]);
  }
  for (int i = 0; i < k; ++i) {
    double wl[2];
    double m = -CUDART_INF;
    for (int c = 0; c < M; ++c) {
      double lp = 0.0;
      for (int d = 0; d < A; ++d) {
        const double pc = 1.0 / sqrt(g.var[c][d]);
        const double prec = pc * pc;
        const double x = X[i * CEM_MAX_A + d], mu = g.mu[c][d];
        lp += mu * mu * prec - 2.0 * x * mu * prec + x * x * prec;
      }
      wl[c] = -0.5 * (A * 1.8378770664093453 + lp) + logdet[c] + logw[c];
      m = fmax(m, wl[c]);
    }
    double se = 0.0;
    for (int c = 0; c < M; ++c) se += exp(wl[c] - m);
    const double norm = m + log(se);
    ll += norm;
    for (int c = 0; c < M; ++c) resp[i * 2 + c] = exp(wl[c] - norm);
    if (M == 1) resp[i * 2 + 1] = 0.0;
  }
  return ll / (double)k;
}

// Deterministic initial partition (oracle_np.farthest_point_labels).
__device__ void gmm_init_labels(const double* X, int k, int A, double* resp) {
  int far = 0;
  double best = -1.0;
  for (int i = 0; i < k; ++i) {
    double d0 = 0.0;
    for (int d = 0; d < A; ++d) {
      const double t = X[i * CEM_MAX_A + d] - X[d];
      d0 += t * t;
    }
    if (d0 > best) { best = d0; far = i; }
  }
  for (int i = 0; i < k; ++i) {
    double d0 = 0.0, d1 = 0.0;
    for (int d = 0; d < A; ++d) {
      const double t0 = X[i * CEM_MAX_A + d] - X[d];
      const double t1 = X[i * CEM_MAX_A + d] - X[far * CEM_MAX_A + d];
      d0 += t0 * t0;
      d1 += t1 * t1;
    }
    const int lab = d1 < d0;
    resp[i * 2 + 0] = lab ? 0.0 : 1.0;
    resp[i * 2 + 1] = lab ? 1.0 : 0.0;
  }
}

// BaseMixture.fit loop given initial responsibilities. Returns n_iter.
__device__ int gmm_fit(const double* X, int k, int A, int M, double* resp, double tol,
                       int max_iter, Gmm& g) {
  gmm_mstep(X, k, A, M, resp, g);
  double lower = -CUDART_INF;
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    const double prev = lower;
    lower = gmm_estep(X, k, A, M, g, resp);
    gmm_mstep(X, k, A, M, resp, g);
    if (fabs(lower - prev) < tol) break;
  }
  return it > max_iter ? max_iter : it;
}

// ---- warp-cooperative versions of the same EM (identical arithmetic and summation order, so the
// results are bit-identical to the serial functions above): the serial refit on one thread was the
// critical path of the CEM kernel (fp64 log/exp/div chains while 255 threads idled).
// Work split: M-step one lane per (component, dim); E-step one lane per (elite, component) pair, then
// one lane per elite for the normalisation.  X, resp, g and the scratch live in shared memory.
struct GmmScratch {
  double wl[CEM_MAX_K * 2];
  double norm[CEM_MAX_K];
  double logdet[2], logw[2];
};

__device__ void gmm_mstep_w(int lane, const double* X, int k, int A, int M, const double* resp, Gmm& g) {
  const double VAR_LO = 0.1353352832366127, VAR_HI = 7.38905609893065;  // exp(-2), exp(2)
  for (int p = lane; p < M * A; p += 32) {
    const int c = p / A, d = p - c * A;
    double nk = 0.0;
    for (int i = 0; i < k; ++i) nk += resp[i * 2 + c];
    nk += 10.0 * 2.220446049250313e-16;
    double sx = 0.0, sx2 = 0.0;
    for (int i = 0; i < k; ++i) {
      const double x = X[i * CEM_MAX_A + d], r = resp[i * 2 + c];
      sx += r * x;
      sx2 += r * x * x;
    }
    const double mean = sx / nk;
    const double cov = sx2 / nk - 2.0 * (mean * sx / nk) + mean * mean + 1e-6;
    g.mu[c][d] = clampd(mean, -2.0, 2.0);
    g.var[c][d] = clampd(cov, VAR_LO, VAR_HI);
    if (d == 0) g.w[c] = nk / (double)k;
  }
  __syncwarp();
}

__device__ double gmm_estep_w(int lane, const double* X, int k, int A, int M, const Gmm& g, double* resp,
                              GmmScratch& sc) {
  if (lane < M) {
    double ld = 0.0;
    for (int d = 0; d < A; ++d) ld += log(1.0 / sqrt(g.var[lane][d]));
    sc.logdet[lane] = ld;
    sc.logw[lane] = log(g.w[lane]);
  }
  __syncwarp();
  for (int p = lane; p < k * M; p += 32) {
    const int i = p / M, c = p - i * M;
    double lp = 0.0;
    for (int d = 0; d < A; ++d) {
      const double pc = 1.0 / sqrt(g.var[c][d]);
      const double prec = pc * pc;
      const double x = X[i * CEM_MAX_A + d], mu = g.mu[c][d];
      lp += mu * mu * prec - 2.0 * x * mu * prec + x * x * prec;
    }
    sc.wl[i * 2 + c] = -0.5 * (A * 1.8378770664093453 + lp) + sc.logdet[c] + sc.logw[c];
  }
  __syncwarp();
  for (int i = lane; i < k; i += 32) {
    double m = -CUDART_INF;
    for (int c = 0; c < M; ++c) m = fmax(m, sc.wl[i * 2 + c]);
    double se = 0.0;
    for (int c = 0; c < M; ++c) se += exp(sc.wl[i * 2 + c] - m);
    const double norm = m + log(se);
    sc.norm[i] = norm;
    for (int c = 0; c < M; ++c) resp[i * 2 + c] = exp(sc.wl[i * 2 + c] - norm);
    if (M == 1) resp[i * 2 + 1] = 0.0;
  }
  __syncwarp();
  double ll = 0.0;
  if (lane == 0) {
    for (int i = 0; i < k; ++i) ll += sc.norm[i];
    ll /= (double)k;
  }
  return __shfl_sync(0xffffffffu, ll, 0);
}

__device__ int gmm_fit_w(int lane, const double* X, int k, int A, int M, double* resp, double tol, int max_iter,
                         Gmm& g, GmmScratch& sc) {
  gmm_mstep_w(lane, X, k, A, M, resp, g);
  double lower = -CUDART_INF;
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    const double prev = lower;
    lower = gmm_estep_w(lane, X, k, A, M, g, resp, sc);
    gmm_mstep_w(lane, X, k, A, M, resp, g);
    if (fabs(lower - prev) < tol) break;   // `lower` is warp-uniform (broadcast from lane 0)
  }
  return it > max_iter ? max_iter : it;
}

// action of sample n in dimension d for the current iteration
__device__ __forceinline__ float cem_action(int it, long long bn, int d, int A, const float* u0,
                                            const float* noise, const float* comp_u, long long BN,
                                            const float* amin, const float* amax, const Gmm& g,
                                            int M) {
  if (it == 0) {
    const double lo = amin[d], hi = amax[d];
    return (float)(lo + (double)u0[bn * A + d] * (hi - lo));
  }
  const long long off = (long long)(it - 1) * BN + bn;
  int c = 0;
  if (M == 2) c = ((double)comp_u[off] >= g.w[0]) ? 1 : 0;
  return (float)(g.mu[c][d] + sqrt(g.var[c][d]) * (double)noise[off * A + d]);
}

template <int AT>
__global__ void __launch_bounds__(CEM_THREADS)
k_cem(const float* __restrict__ p, int B, int N, int A, int H2, int iters, int top_m, int M,
      const float* __restrict__ W2a, const float* __restrict__ w3, const float* __restrict__ b3,
      const float* __restrict__ u0, const float* __restrict__ noise,
      const float* __restrict__ comp_u, const float* __restrict__ amin,
      const float* __restrict__ amax, float* __restrict__ weights_out,
      float* __restrict__ means_out, float* __restrict__ vars_out, float* __restrict__ best_out,
      long long* __restrict__ elite_idx_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int H2P = (H2 + 3) & ~3;
  float* ps = reinterpret_cast<float*>(smraw);  // [H2P]
  float* w3s = ps + H2P;                        // [H2P]
  float* was = w3s + H2P;                       // [AT][H2P]
  float* qbuf = was + AT * H2P;                 // [N]
  __shared__ Gmm g;
  __shared__ double Xel[CEM_MAX_K * CEM_MAX_A];
  __shared__ double resp[CEM_MAX_K * 2];
  __shared__ GmmScratch gsc;
  __shared__ float red_v[CEM_THREADS / 32];
  __shared__ int red_i[CEM_THREADS / 32];
  __shared__ int sel[CEM_MAX_K];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long BN = (long long)B * N;
  for (int i = tid; i < H2P; i += CEM_THREADS) {
    ps[i] = (i < H2) ? p[(long long)b * H2 + i] : 0.f;
    w3s[i] = (i < H2) ? w3[i] : 0.f;
  }
  for (int i = tid; i < AT * H2P; i += CEM_THREADS) {
    const int ai = i / H2P, j = i - ai * H2P;
    was[i] = (ai < A && j < H2) ? W2a[(long long)ai * H2 + j] : 0.f;
  }
  __syncthreads();
  const float bb3 = b3[0];

  for (int it = 0; it < iters; ++it) {
    // ---- sample + evaluate ----
    for (int n = tid; n < N; n += CEM_THREADS) {
      const long long bn = (long long)b * N + n;
      float ar[AT];
#pragma unroll
      for (int i = 0; i < AT; ++i)
        ar[i] = (i < A) ? cem_action(it, bn, i, A, u0, noise, comp_u, BN, amin, amax, g, M) : 0.f;
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
      qbuf[n] = q + bb3;
    }
    __syncthreads();
    // ---- top-m tournament (descending, ties -> larger index) ----
    float lastv = CUDART_INF_F;
    int lasti = 0x7fffffff;
    for (int t = 0; t < top_m; ++t) {
      float bv = -CUDART_INF_F;
      int bi = -1;
      for (int n = tid; n < N; n += CEM_THREADS) {
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
        for (int w = 1; w < CEM_THREADS / 32; ++w) {
          const float ov = red_v[w];
          const int oi = red_i[w];
          if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; }
        }
        red_v[0] = bv;
        red_i[0] = bi;
        sel[t] = bi;
        if (elite_idx_out) elite_idx_out[((long long)it * B + b) * top_m + t] = bi;
      }
      __syncthreads();
      lastv = red_v[0];
      lasti = red_i[0];
      __syncthreads();
    }
    // ---- gather elites (recomputed from the draws; mixture of the previous iter still live) ----
    for (int e = tid; e < top_m * A; e += CEM_THREADS) {
      const int t = e / A, d = e - t * A;
      const int n = sel[t];
      Xel[t * CEM_MAX_A + d] =
          (n >= 0) ? (double)cem_action(it, (long long)b * N + n, d, A, u0, noise, comp_u, BN,
                                        amin, amax, g, M)
                   : 0.0;
    }
    __syncthreads();
    // ---- refit: warp 0, cooperatively ----
    if (wid == 0) {
      if (M == 1) {
        for (int i = lane; i < top_m; i += 32) { resp[i * 2] = 1.0; resp[i * 2 + 1] = 0.0; }
        __syncwarp();
        gmm_mstep_w(lane, Xel, top_m, A, 1, resp, g);
        if (lane == 0) g.w[0] = 1.0;
      } else {
        if (lane == 0) gmm_init_labels(Xel, top_m, A, resp);
        __syncwarp();
        gmm_fit_w(lane, Xel, top_m, A, 2, resp, 1e-2, 100, g, gsc);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int best = (M == 2 && g.w[1] > g.w[0]) ? 1 : 0;
    for (int c = 0; c < M; ++c) {
      weights_out[(long long)b * M + c] = (float)g.w[c];
      for (int d = 0; d < A; ++d) {
        means_out[((long long)b * M + c) * A + d] = (float)g.mu[c][d];
        vars_out[((long long)b * M + c) * A + d] = (float)g.var[c][d];
      }
    }
    if (best_out)
      for (int d = 0; d < A; ++d) best_out[(long long)b * A + d] = (float)g.mu[best][d];
  }
}

extern "C" int rlc_cem(rlc_handle* h, const rlc_critic* c, const float* s, int B, int N, int iters,
                       int top_m, int num_modal, const float* u0, const float* noise,
                       const float* comp_u, const float* amin, const float* amax,
                       float* weights_out, float* means_out, float* vars_out,
                       float* best_action_out, int64_t* elite_idx_out, void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && u0 && amin && amax && weights_out && means_out && vars_out);
  RLC_REQUIRE(B >= 0 && N >= 1 && iters >= 1 && top_m >= 1 && top_m <= CEM_MAX_K && top_m <= N);
  RLC_REQUIRE(num_modal == 1 || num_modal == 2);
  RLC_REQUIRE(iters == 1 || (noise && (num_modal == 1 || comp_u)));
  if (c->topology != RLC_TMID || c->A > CEM_MAX_A) return RLC_ERR_UNSUPPORTED;
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
#define RLC_CEM_CASE(AT)                                                                         \
  {                                                                                              \
    const size_t smem = ((size_t)H2P * (2 + AT) + (size_t)N) * sizeof(float);                    \
    if (smem + 4096 > h->smem_optin) return RLC_ERR_UNSUPPORTED;                                 \
    auto kern = k_cem<AT>;                                                                       \
    RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<B, CEM_THREADS, smem, st>>>((const float*)ws, B, N, c->A, c->H2, iters, top_m,        \
                                       num_modal, W2a, c->theta + t.ow3, c->theta + t.ob3, u0,   \
                                       noise, comp_u, amin, amax, weights_out, means_out,        \
                                       vars_out, best_action_out, (long long*)elite_idx_out);    \
  }
  if (c->A <= 1) RLC_CEM_CASE(1)
  else if (c->A <= 2) RLC_CEM_CASE(2)
  else if (c->A <= 3) RLC_CEM_CASE(3)
  else if (c->A <= 4) RLC_CEM_CASE(4)
  else if (c->A <= 6) RLC_CEM_CASE(6)
  else RLC_CEM_CASE(8)
#undef RLC_CEM_CASE
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// Standalone refit: one thread per state.
__global__ void k_gmm_refit(const float* __restrict__ X, int B, int k, int A, int M,
                            const float* __restrict__ resp0, double tol, int max_iter,
                            float* __restrict__ weights_out, float* __restrict__ means_out,
                            float* __restrict__ vars_out, int* __restrict__ n_iter_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double Xl[CEM_MAX_K * CEM_MAX_A];
  double resp[CEM_MAX_K * 2];
  Gmm g;
  for (int i = 0; i < k; ++i)
    for (int d = 0; d < A; ++d) Xl[i * CEM_MAX_A + d] = (double)X[((long long)b * k + i) * A + d];
  int nit = 0;
  if (M == 1) {
    for (int i = 0; i < k; ++i) { resp[i * 2] = 1.0; resp[i * 2 + 1] = 0.0; }
    nit = gmm_fit(Xl, k, A, 1, resp, tol, max_iter, g);
  } else {
    if (resp0) {
      for (int i = 0; i < k; ++i) {
        resp[i * 2] = (double)resp0[((long long)b * k + i) * 2];
        resp[i * 2 + 1] = (double)resp0[((long long)b * k + i) * 2 + 1];
      }
    } else {
      gmm_init_labels(Xl, k, A, resp);
    }
    nit = gmm_fit(Xl, k, A, 2, resp, tol, max_iter, g);
  }
  for (int c = 0; c < M; ++c) {
    weights_out[(long long)b * M + c] = (float)g.w[c];
    for (int d = 0; d < A; ++d) {
      means_out[((long long)b * M + c) * A + d] = (float)g.mu[c][d];
      vars_out[((long long)b * M + c) * A + d] = (float)g.var[c][d];
    }
  }
  if (n_iter_out) n_iter_out[b] = nit;
}

extern "C" int rlc_gmm_refit(rlc_handle* h, const float* X, int B, int k, int A, int num_modal,
                             const float* resp0, float tol, int max_iter, float* weights_out,
                             float* means_out, float* vars_out, int32_t* n_iter_out,
                             void* stream) {
  RLC_REQUIRE(h && X && weights_out && means_out && vars_out && B >= 0);
  RLC_REQUIRE(k >= 1 && k <= CEM_MAX_K && A >= 1 && A <= CEM_MAX_A);
  RLC_REQUIRE((num_modal == 1 || num_modal == 2) && max_iter >= 1 && tol >= 0.f);
  if (B == 0) return RLC_OK;
  k_gmm_refit<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      X, B, k, A, num_modal, resp0, (double)tol, max_iter, weights_out, means_out, vars_out,
      n_iter_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
