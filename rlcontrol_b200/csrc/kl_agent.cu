// B-row pieces of the ForwardKL / ReverseKL update_network around the sampled-action hot path
// (forwardkl_network.py:123-209, reversekl_network.py:130-217): the value and policy MLPs
// (ValueNetwork :270-290, PolicyNetwork.forward :293-322) as one generic two-hidden-layer MLP with O
// linear outputs, PolicyNetwork.evaluate with the normal draws fed in, the TD / soft-value targets
// and the log-likelihood policy gradients.  fp32 CUDA cores on the shared tiled SGEMM
// (rows_gemm.cuh); everything is stream-ordered and graph-capturable.
#include "rows_gemm.cuh"
#include "policy_math.cuh"

struct MlpView {
  int64_t oW1, ob1, oW2, ob2, oW3, ob3, numel;
};
static inline MlpView mlp_view(int in, int H1, int H2, int O) {
  MlpView t;
  t.oW1 = 0;
  t.ob1 = t.oW1 + (int64_t)in * H1;
  t.oW2 = t.ob1 + H1;
  t.ob2 = t.oW2 + (int64_t)H1 * H2;
  t.oW3 = t.ob2 + H2;
  t.ob3 = t.oW3 + (int64_t)H2 * O;
  t.numel = t.ob3 + O;
  return t;
}
static inline bool mlp_ok(const rlc_mlp* m) {
  return m && m->theta && m->in >= 1 && m->H1 >= 1 && m->H2 >= 1 && m->O >= 1 && m->in <= 4096 &&
         m->H1 <= 2048 && m->H2 <= 2048 && m->O <= 64;
}

extern "C" int64_t rlc_mlp_numel(int in, int H1, int H2, int O) {
  if (in < 1 || H1 < 1 || H2 < 1 || O < 1) return -1;
  return mlp_view(in, H1, H2, O).numel;
}
extern "C" int rlc_mlp_offsets(int in, int H1, int H2, int O, int64_t off[6]) {
  RLC_REQUIRE(off && in >= 1 && H1 >= 1 && H2 >= 1 && O >= 1);
  const MlpView t = mlp_view(in, H1, H2, O);
  off[0] = t.oW1; off[1] = t.ob1; off[2] = t.oW2; off[3] = t.ob2; off[4] = t.oW3; off[5] = t.ob3;
  return RLC_OK;
}

// The GEMM every B-row layer above is made of, exposed for parity tests and callers with their own layer structure.
extern "C" int rlc_rows_gemm(rlc_handle* h, int trans_a, int trans_b, int M, int N, int K, const float* A, int lda,
                             const float* Bm, int ldb, float* C, int ldc, const float* bias, const float* maskZ, int ldz,
                             int relu_a, float alpha, int split_k, int path, void* stream) {
  RLC_REQUIRE(h && A && Bm && C && M >= 0 && N >= 0 && K >= 0 && lda >= 1 && ldb >= 1 && ldc >= N);
  RLC_REQUIRE(path == 0 || path == 1 || path == 2);
  RLC_REQUIRE(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N) && (!maskZ || ldz >= N));
  RLC_REQUIRE(!split_k || (trans_a && !trans_b && !bias && !maskZ && ldc == N));
  cudaStream_t st = (cudaStream_t)stream;
  GemmEpi e{bias, maskZ, ldz, relu_a, alpha};
  const int prev = rlc_gemm_tc_forced();
  if (path != 0) rlc_gemm_tc_force(path == 1 ? 0 : 2);
  int rc;
  if (split_k) {
    void* ws = nullptr;
    rc = rlc_workspace(h, (size_t)SPLITK_MAX * M * N * sizeof(float) + 256, &ws);
    if (!rc) rc = gemm_splitk(h, M, N, K, A, lda, Bm, ldb, C, e, (float*)ws, st);
  } else {
    rc = gemm(h, trans_a != 0, trans_b != 0, M, N, K, A, lda, Bm, ldb, C, ldc, e, st);
  }
  rlc_gemm_tc_force(prev);
  return rc;
}

// Z1 = X W1 + b1 ; Z2 = relu(Z1) W2 + b2 ; out = relu(Z2) W3 + b3   (pre-activations kept)
static int mlp_forward_rows(rlc_handle* h, const rlc_mlp* m, const float* x, int B, float* Z1,
                            float* Z2, float* out, cudaStream_t st) {
  const MlpView t = mlp_view(m->in, m->H1, m->H2, m->O);
  const float* th = m->theta;
  GemmEpi e1{th + t.ob1, nullptr, 0, 0, 1.f};
  int rc = gemm(h, false, false, B, m->H1, m->in, x, m->in, th + t.oW1, m->H1, Z1, m->H1, e1, st);
  if (rc) return rc;
  GemmEpi e2{th + t.ob2, nullptr, 0, 1, 1.f};
  rc = gemm(h, false, false, B, m->H2, m->H1, Z1, m->H1, th + t.oW2, m->H2, Z2, m->H2, e2, st);
  if (rc) return rc;
  GemmEpi e3{th + t.ob3, nullptr, 0, 1, 1.f};
  return gemm(h, false, false, B, m->O, m->H2, Z2, m->H2, th + t.oW3, m->O, out, m->O, e3, st);
}

static size_t mlp_ws_floats(const rlc_mlp* m, long long B, bool need_act, bool need_back) {
  size_t n = 64;
  if (need_act) n += (size_t)B * ((size_t)m->H1 + m->H2) + 8;
  if (need_back) {
    const size_t maxw = (size_t)(m->H1 > m->H2 ? m->H1 : m->H2) + 1;
    const size_t maxin = (size_t)(m->in > m->H1 ? m->in : m->H1);
    n += (size_t)B * ((size_t)m->H1 + m->H2 + m->O) + 16;
    n += (size_t)SPLITK_MAX * ((maxin > (size_t)m->H2 ? maxin : (size_t)m->H2) + 1) * maxw;
    n += (size_t)COLRED_MAX_CHUNKS * maxw;
  }
  return n;
}

extern "C" int rlc_mlp_forward(rlc_handle* h, const rlc_mlp* m, const float* x, int B, float* out,
                               float* act, void* stream) {
  RLC_REQUIRE(h && mlp_ok(m) && x && out && B >= 0);
  if (B == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* Z1 = act;
  if (!Z1) {
    void* ws = nullptr;
    int rc = rlc_workspace(h, mlp_ws_floats(m, B, true, false) * sizeof(float), &ws);
    if (rc) return rc;
    Z1 = (float*)ws;
  }
  float* Z2 = Z1 + (((size_t)B * m->H1 + 3) & ~(size_t)3);
  return mlp_forward_rows(h, m, x, B, Z1, Z2, out, st);
}

extern "C" int64_t rlc_mlp_act_numel(int H1, int H2, int B) {
  if (H1 < 1 || H2 < 1 || B < 0) return -1;
  return (int64_t)((((size_t)B * H1 + 3) & ~(size_t)3) + (size_t)B * H2);
}

extern "C" int rlc_mlp_grads(rlc_handle* h, const rlc_mlp* m, const float* x, const float* act,
                             const float* dout, int B, float* grad_out, float* dx_out,
                             void* stream) {
  RLC_REQUIRE(h && mlp_ok(m) && x && dout && grad_out && B >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  const MlpView t = mlp_view(m->in, m->H1, m->H2, m->O);
  const float* th = m->theta;
  void* ws = nullptr;
  int rc = rlc_workspace(h, mlp_ws_floats(m, B, act == nullptr, true) * sizeof(float), &ws);
  if (rc) return rc;
  float* base = (float*)ws;
  auto take = [&](size_t n) {
    float* p = base;
    base += (n + 3) & ~(size_t)3;
    return p;
  };
  const float *Z1, *Z2;
  if (act) {
    Z1 = act;
    Z2 = act + (((size_t)B * m->H1 + 3) & ~(size_t)3);
  } else {
    float* z1 = take((size_t)B * m->H1);
    float* z2 = take((size_t)B * m->H2);
    float* o = take((size_t)B * m->O);
    rc = mlp_forward_rows(h, m, x, B, z1, z2, o, st);
    if (rc) return rc;
    Z1 = z1;
    Z2 = z2;
  }
  float* G2 = take((size_t)B * m->H2);
  float* G1 = take((size_t)B * m->H1);
  const size_t maxw = (size_t)(m->H1 > m->H2 ? m->H1 : m->H2) + 1;
  const size_t maxin = (size_t)(m->in > m->H1 ? m->in : m->H1);
  float* slabs = take((size_t)SPLITK_MAX * ((maxin > (size_t)m->H2 ? maxin : (size_t)m->H2) + 1) * maxw);
  float* part = take((size_t)COLRED_MAX_CHUNKS * maxw);
  GemmEpi e{nullptr, nullptr, 0, 0, 1.f};
  GemmEpi er{nullptr, nullptr, 0, 1, 1.f};
  // every layer: [gW ; gb] = [input | 1]^T G, contiguous in theta
  // output layer: gW3 = relu(Z2)^T dout ; gb3 = colsum(dout)
  rc = gemm_splitk_bias(h, m->H2, m->O, B, Z2, m->H2, dout, m->O, grad_out + t.oW3, er, slabs, part, st);
  if (rc) return rc;
  // G2 = (dout W3^T) * [Z2 > 0]
  GemmEpi em2{nullptr, Z2, m->H2, 0, 1.f};
  rc = gemm(h, false, true, B, m->H2, m->O, dout, m->O, th + t.oW3, m->O, G2, m->H2, em2, st);
  if (rc) return rc;
  rc = gemm_splitk_bias(h, m->H1, m->H2, B, Z1, m->H1, G2, m->H2, grad_out + t.oW2, er, slabs, part, st);
  if (rc) return rc;
  // G1 = (G2 W2^T) * [Z1 > 0]
  GemmEpi em1{nullptr, Z1, m->H1, 0, 1.f};
  rc = gemm(h, false, true, B, m->H1, m->H2, G2, m->H2, th + t.oW2, m->H2, G1, m->H1, em1, st);
  if (rc) return rc;
  rc = gemm_splitk_bias(h, m->in, m->H1, B, x, m->in, G1, m->H1, grad_out + t.oW1, e, slabs, part, st);
  if (rc) return rc;
  if (dx_out)  // dX = G1 W1^T
    rc = gemm(h, false, true, B, m->in, m->H1, G1, m->H1, th + t.oW1, m->H1, dx_out, m->in, e, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// PolicyNetwork.evaluate (forwardkl_network.py:303-322 / reversekl_network.py:325-344) on the
// policy head head[B,2A] = [mean_raw | log_std_raw], with the standard-normal draws eps[B,A]
// supplied (torch's normal.sample() is loc + scale * N(0,1) from the same stream).
// A == 1: Normal(mean, std).  A > 1: MultivariateNormal(mean, covariance = diag_embed(std)) exactly
// as the reference passes it (get_distribution :346-351), i.e. scale = sqrt(std).
// One thread per state.
// ---------------------------------------------------------------------------------------------
__global__ void k_policy_evaluate(const float* __restrict__ head, const float* __restrict__ eps,
                                  int B, int A, float scale, float lo, float hi,
                                  float* __restrict__ action, float* __restrict__ logp,
                                  float* __restrict__ mean_out, float* __restrict__ mu_raw_out,
                                  float* __restrict__ log_std_out, float* __restrict__ z_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long o = (long long)b * A;
  policy_evaluate_row(head + 2 * o, eps ? eps + o : nullptr, A, scale, lo, hi, action ? action + o : nullptr,
                      logp ? logp + b : nullptr, mean_out ? mean_out + o : nullptr,
                      mu_raw_out ? mu_raw_out + o : nullptr, log_std_out ? log_std_out + o : nullptr,
                      z_out ? z_out + o : nullptr);
}

extern "C" int rlc_policy_evaluate(rlc_handle* h, const float* head, const float* eps, int B, int A,
                                   float action_scale, float log_std_min, float log_std_max,
                                   float* action_out, float* logp_out, float* mean_out,
                                   float* mu_raw_out, float* log_std_out, float* z_out,
                                   void* stream) {
  RLC_REQUIRE(h && head && B >= 0 && A >= 1 && A <= 64 && log_std_min <= log_std_max);
  if (B == 0) return RLC_OK;
  k_policy_evaluate<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      head, eps, B, A, action_scale, log_std_min, log_std_max, action_out, logp_out, mean_out,
      mu_raw_out, log_std_out, z_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Regression targets of the update (forwardkl_network.py:137-150 / reversekl_network.py:144-157):
//   y_q      = r + gamma * V_target(s')
//   target_v = Q(s, a_new) - alpha * logp                       ('sac')
//            = (r - alpha * logp) + gamma * V_target(s')        ('non_sac')
//   dv       = d mean_b (v - target_v)^2 / dv = 2 (v - target_v) / B_total ; v_loss += sum (..)^2 / B_total
// ---------------------------------------------------------------------------------------------
__global__ void k_kl_targets(const float* __restrict__ r, const float* __restrict__ gamma,
                             const float* __restrict__ v_next, const float* __restrict__ q_new,
                             const float* __restrict__ logp, const float* __restrict__ v, int B,
                             float alpha, int sac, float inv_btotal, float* __restrict__ y_q,
                             float* __restrict__ dv, float* __restrict__ v_loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (b < B) {
    const float gv = gamma[b] * v_next[b];
    if (y_q) y_q[b] = r[b] + gv;
    if (dv) {
      const float tv = sac ? (q_new[b] - alpha * logp[b]) : ((r[b] - alpha * logp[b]) + gv);
      const float d = v[b] - tv;
      dv[b] = 2.f * inv_btotal * d;
      sq = d * d * inv_btotal;
    }
  }
  if (v_loss) {
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0 && sq != 0.f) atomicAdd(v_loss, sq);
  }
}

extern "C" int rlc_kl_targets(rlc_handle* h, const float* r, const float* gamma, const float* v_next,
                              const float* q_new, const float* logp, const float* v, int B,
                              int B_total, float entropy_scale, int sac, float* y_q_out,
                              float* dv_out, float* v_loss_out, void* stream) {
  RLC_REQUIRE(h && r && gamma && v_next && B >= 0 && B_total >= B);
  RLC_REQUIRE(!dv_out || (logp && v && (!sac || q_new)));
  cudaStream_t st = (cudaStream_t)stream;
  if (v_loss_out) RLC_CUDA(cudaMemsetAsync(v_loss_out, 0, sizeof(float), st));
  if (B == 0) return RLC_OK;
  k_kl_targets<<<(B + 127) / 128, 128, 0, st>>>(r, gamma, v_next, q_new, logp, v, B, entropy_scale,
                                                sac, 1.f / (float)B_total, y_q_out, dv_out,
                                                v_loss_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Gradient wrt the raw policy head [mean_raw | log_std_raw]:
//   mode 0 ('intg' / 'hard_intg'): chain dmean, dlog_std (wrt forward()'s outputs, from
//           rlc_reduce_{fkl,rkl}_policy) through torch.clamp (gradient passes where lo <= x <= hi).
//   mode 1 ('ll', reversekl_network.py:161-165): loss = mean_b( -logp_b * (q_new - v - alpha logp)_detached )
//   mode 2 ('hard_ll', :167-169):                loss = mean_b( -logp_b * (q_new - v)_detached )
//           with logp_b the density of the DETACHED sample z (normal.sample()), so only the
//           Gaussian term carries gradient:  A==1: dlogp/dmu = (z-mu)/std^2, dlogp/dls = (z-mu)^2/std^2 - 1
//                                            A>1 (covariance = std): (z-mu)/std, 0.5 (z-mu)^2/std - 0.5
// ---------------------------------------------------------------------------------------------
__global__ void k_policy_head_grad(const float* __restrict__ head, int B, int A, float lo, float hi,
                                   int mode, const float* __restrict__ dmean,
                                   const float* __restrict__ dlog_std, const float* __restrict__ z,
                                   const float* __restrict__ logp, const float* __restrict__ q_new,
                                   const float* __restrict__ v, float alpha, float inv_btotal,
                                   float* __restrict__ dhead, float* __restrict__ loss_acc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * A) return;
  const long long b = i / A;
  const int d = (int)(i - b * A);
  const float raw = head[b * 2 * A + A + d];
  const float pass = (raw >= lo && raw <= hi) ? 1.f : 0.f;
  float gm, gs;
  if (mode == 0) {
    gm = dmean[i];
    gs = dlog_std[i];
  } else {
    const float c = (mode == 1) ? (q_new[b] - v[b] - alpha * logp[b]) : (q_new[b] - v[b]);
    const float coef = -c * inv_btotal;
    const float mu = head[b * 2 * A + d];
    const float ls = fminf(fmaxf(raw, lo), hi);
    const float t = z[i] - mu;
    if (A == 1) {
      const float iv = expf(-2.f * ls);
      gm = coef * t * iv;
      gs = coef * (t * t * iv - 1.f);
    } else {
      const float iv = expf(-ls);
      gm = coef * t * iv;
      gs = coef * (0.5f * t * t * iv - 0.5f);
    }
    if (loss_acc && d == 0) atomicAdd(loss_acc, -logp[b] * c * inv_btotal);
  }
  dhead[b * 2 * A + d] = gm;
  dhead[b * 2 * A + A + d] = gs * pass;
}

extern "C" int rlc_policy_head_grad(rlc_handle* h, const float* head, int B, int A,
                                    float log_std_min, float log_std_max, int mode,
                                    const float* dmean, const float* dlog_std, const float* z,
                                    const float* logp, const float* q_new, const float* v,
                                    float entropy_scale, int B_total, float* dhead_out,
                                    float* loss_out, void* stream) {
  RLC_REQUIRE(h && head && dhead_out && B >= 0 && A >= 1 && A <= 64 && B_total >= B);
  RLC_REQUIRE(mode >= 0 && mode <= 2);
  RLC_REQUIRE(mode != 0 || (dmean && dlog_std));
  RLC_REQUIRE(mode == 0 || (z && logp && q_new && v));
  cudaStream_t st = (cudaStream_t)stream;
  if (loss_out && mode != 0) RLC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (B == 0) return RLC_OK;
  const long long n = (long long)B * A;
  k_policy_head_grad<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
      head, B, A, log_std_min, log_std_max, mode, dmean, dlog_std, z, logp, q_new, v,
      entropy_scale, 1.f / (float)B_total, dhead_out, mode != 0 ? loss_out : nullptr);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
