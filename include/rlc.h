/* rlc.h -- C-ABI of librlc.so: the B200 (sm_100a) sampled-action critic path of RLControl.
 *
 * Drop-in boundary.  Every entry point takes plain pointers and sizes, returns an
 * int status (0 = ok, negative = error, see rlc_status_string), never throws, never owns caller
 * memory and is stream-ordered on the cudaStream_t passed as `void* stream` (NULL = default
 * stream).  All data pointers are DEVICE pointers unless the name ends in `_host`.
 * The only state is the opaque rlc_handle (workspace + pre-packed tensor-core operands).
 * A handle is not thread-safe (same as one reference network object: single Python thread,
 * agents/base_agent.py:40-70).
 *
 * Each function cites the reference interface (file:line under the RLControl checkout) that a
 * maintainer would re-bind to it; INTEGRATION.md shows the ctypes stub.
 *
 * Canonical parameter vector `theta` (fp32, contiguous, caller-owned):
 *     [ W1 (in1 x H1) | b1 (H1) | W2 (in2 x H2) | b2 (H2) | w3 (H2) | b3 (1) ]   all [in,out]-major
 *   T-in  (SoftQNetwork, forwardkl_network.py:250-268):  in1 = S+A, in2 = H1
 *   T-mid (critic_network.py:77-99 family):              in1 = S,   in2 = H1+A (rows H1.. = action)
 * Gradients and Adam moments use the same layout, so the data-parallel all-reduce is one
 * contiguous buffer.
 */
#ifndef RLC_H_
#define RLC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLC_VERSION 100

/* status codes */
#define RLC_OK 0
#define RLC_ERR_INVALID (-1)  /* bad shape / null pointer / unsupported combination */
#define RLC_ERR_ARCH (-2)     /* device is not sm_100 (tensor-core path) */
#define RLC_ERR_ALLOC (-3)    /* workspace allocation failed */
#define RLC_ERR_CUDA (-4)     /* CUDA runtime error (rlc_last_cuda_error has the text) */
#define RLC_ERR_UNSUPPORTED (-5)

/* critic topology */
#define RLC_TIN 0  /* concat [s;a] at the input: SoftQNetwork, sql qf_network */
#define RLC_TMID 1 /* concat [h1;a] at layer 2: critic_network.py family      */

/* source layout of the six tensors handed to rlc_pack_theta */
#define RLC_LAYOUT_OUT_IN 0 /* torch nn.Linear weight [out,in] (forwardkl_network.py:254-256) */
#define RLC_LAYOUT_IN_OUT 1 /* tf.contrib.layers.fully_connected kernel [in,out]              */

/* action tensor layout for B x N evaluation */
#define RLC_ACT_SHARED 0    /* a[N,A] broadcast over states (quadrature grid, forwardkl_network.py:104-105) */
#define RLC_ACT_PER_STATE 1 /* a[B,N,A] (ActorExpert.py:166, qt_opt_network.py:160)                         */

/* arithmetic of the critic evaluation */
#define RLC_PREC_FP32 0 /* CUDA-core fp32 FMA, fp32 accumulate (bit-comparable to the CPU path ~1e-6) */
#define RLC_PREC_FP16 1 /* tcgen05.mma kind::f16, fp16 operands, fp32 accumulate in TMEM (default for large B*N) */
#define RLC_PREC_BF16 2 /* tcgen05.mma kind::f16, bf16 operands, fp32 accumulate in TMEM */
#define RLC_PREC_AUTO 3 /* parity-preserving choice: FP16X3 for shared-grid T-in evaluations of >= 16384 rows, else FP32 */
#define RLC_PREC_FP16X3 4 /* strict tensor mode (shared grids): both operands split into fp16 hi+lo, three tcgen05.mma per
                             K step into one fp32 TMEM accumulator (h_hi.W_hi + h_lo.W_hi + h_hi.W_lo): 22-bit operands,
                             ~1e-5 of the exact Q (fp32-class parity with forwardkl_network.py:263-268) at 3x the MMA work */

#define RLC_PREC_FP16C8 5 /* fast split mode (shared grids): the fp16 product plus its two corrections on the FP8 pipe (kind::f8f6f4,
                             twice the MMA rate): h_hi.W_hi + e4m3(2^9 h_lo).e4m3(2^-9 W_hi) + e5m2(h_hi).e4m3(W_lo): ~2e-4 max of
                             the exact Q (inside north_star's 1e-3 with margin) at 2/3 of FP16X3's tensor-pipe time */

/* Adam flavour of rlc_adam_step */
#define RLC_ADAM_TORCH 0 /* torch.optim.Adam: p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) */
#define RLC_ADAM_TF 1    /* tf.train.AdamOptimizer: p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps) */

typedef struct rlc_handle rlc_handle;

/* One critic network = dims + a pointer to its canonical parameter vector.
 * smin/smax: device pointers to S floats each, or NULL.  Only honoured for RLC_TMID and the TF
 * T-in critic (tf.clip_by_value(inputs, state_min, state_max), critic_network.py:71,
 * sql_network.py:278-279); the torch SoftQNetwork never clips. */
typedef struct rlc_critic {
  int32_t topology;
  int32_t S, A, H1, H2;
  const float* theta;
  const float* smin;
  const float* smax;
} rlc_critic;

/* ---- lifecycle ------------------------------------------------------------------------- */
int rlc_version(void);
const char* rlc_status_string(int status);
const char* rlc_last_cuda_error(void);
int rlc_create(rlc_handle** out, int device);
int rlc_destroy(rlc_handle* h);
/* number of kernels this handle has launched (bench.py's gpu_launches) */
int64_t rlc_launch_count(const rlc_handle* h);

/* ---- parameters ------------------------------------------------------------------------ */
int64_t rlc_theta_numel(int topology, int S, int A, int H1, int H2);
/* off[0..5] = element offsets of W1,b1,W2,b2,w3,b3 inside theta */
int rlc_theta_offsets(int topology, int S, int A, int H1, int H2, int64_t off[6]);
/* Replaces `linearN.weight/bias` (forwardkl_network.py:254-259) or the TF variables
 * `fully_connected_N/weights|biases` (critic_network.py:79-97): device tensors in `layout` ->
 * canonical theta. */
int rlc_pack_theta(int topology, int S, int A, int H1, int H2, int layout, const float* W1,
                   const float* b1, const float* W2, const float* b2, const float* W3,
                   const float* b3, float* theta_out, void* stream);
int rlc_unpack_theta(int topology, int S, int A, int H1, int H2, int layout, const float* theta,
                     float* W1, float* b1, float* W2, float* b2, float* W3, float* b3,
                     void* stream);
/* Tell the handle that c->theta changed (drops the cached fp16/bf16 operand pack). The step
 * functions below call it themselves. */
int rlc_invalidate_pack(rlc_handle* h, const float* theta);

/* ---- critic evaluation (rows a1, a2, a6, a7) --------------------------------------------
 * Replaces q_net(stacked_s, stacked_a) (forwardkl_network.py:160-164, reversekl_network.py:176-181)
 * and predict_q / predict_q_target(inputs, action, phase) (ae_network.py:377-399,
 * qt_opt_network.py:107-129, critic_network.py:101-123) on the state-major stack
 * row = b*N + n.  s[B,S]; a per act_mode; q_out[B,N] fp32.  The broadcast/concat is fused:
 * the stacked tensors are never materialised. */
int rlc_critic_eval(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a,
                    int N, int act_mode, int precision, float* q_out, void* stream);

/* The whole sampled-action step of ForwardKL (mode 0, forwardkl_network.py:160-194 + get_logprob :324-351) or ReverseKL
 * (mode 1, reversekl_network.py:176-203; `hard` = the hard_intg variant, v[B] = V(s)) in ONE call: Q(s_b, a_n) on the shared grid
 * [N,A] and its per-state policy reduction, loss_b[B], dL/dmean[B,A], dL/dlog_std[B,A] (gradients scaled by 1/B_total).
 * fuse != 0 with precision RLC_PREC_FP16X3 / FP16C8 (or AUTO) and B >= 8 x the SM count: the reduction runs inside the
 * evaluation kernel's epilogue (state-major tiles, online softmax over the state's action blocks) and q[B,N] never reaches
 * memory unless q_out is given.  Otherwise (and always with fuse == 0, which is ~30 us faster at cfg4: the fused variant pays
 * the 7-vs-6.9 state-group imbalance of 148 persistent CTAs) it composes rlc_critic_eval + rlc_reduce_{fkl,rkl}_policy with
 * the same results.  q_out may be NULL. */
int rlc_critic_eval_reduce_policy(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* grid, int N,
                                  const float* w, float action_scale, const float* mean, const float* log_std,
                                  const float* v, float entropy_scale, int mode, int hard, int B_total, int precision,
                                  int fuse, float* q_out, float* loss_b_out, float* dmean_out, float* dlog_std_out,
                                  void* stream);

/* B x N evaluation *and* dQ/da of a T-mid critic without materialising the stack: the AE+
 * ascent (ae_plus_network.py:310-343, ActorExpert_Plus.py:130) evaluates action_grads on B*N rows
 * whose states repeat N times.  dqda_out[B,N,A]; q_out[B,N] or NULL. */
int rlc_tmid_eval_grad(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a,
                       int N, int act_mode, float* q_out, float* dqda_out, void* stream);
/* Diagnostic: synchronises `stream` and returns the code raised by a bounded mbarrier wait inside
 * the tcgen05 kernel (0 = none, <0 = CUDA error). The flag is cleared on read. */
int rlc_umma_last_error(rlc_handle* h, void* stream);
/* Diagnostic: which stated arithmetic the tensor path uses for this critic shape and action layout
 * (tests pick the matching oracle restatement, oracle_np.tin_eval_rounded(head=...)):
 *  -1  shape not supported by the tensor path
 *   0  "ss"     q = b3 + sum_j w3_j relu(h1.r(W2[:,j]) + b2_j), h1 = r(relu(r([s;a;1]).r(W1')))
 *   1  "folded" same h1; output head folded into layer 2's operands:
 *               q = b3 + 2^-k sum_j sign(w3_j) relu(h1.r(2^k|w3_j|W2[:,j]) + r(2^k|w3_j|b2_j))
 *   3  "grid"   shared action grids: hoisted layer 1, h1 = relu(r(r(b1+W1s s_b) + r(W1a a_n))),
 *               then the folded head (K1-grid kernel)
 * r() = rounding to the operand type (fp16/bf16); accumulation is fp32.  Env knobs (debug):
 * RLC_UMMA_MODE=ss selects 0; RLC_UMMA_GRID=0 disables 3. */
int rlc_umma_mode(const rlc_critic* c, int act_mode);
/* Same for an explicit precision; adds
 *   4  "grid3"  RLC_PREC_FP16X3, shared grids: h = fl32(PS[b] + PA[n]) from fp32 tables, relu, h = h_hi + h_lo (fp16 each);
 *               W' = fl32(2^k |w3_j| W2[:,j]) = W_hi + W_lo (fp16 each, 2^k also normalises max|W2| to [2^8,2^9));
 *               z = h_hi.W_hi + h_lo.W_hi + h_hi.W_lo, then the folded head.
 *   5  "grid3c8" RLC_PREC_FP16C8: same h, W' and hi parts; the corrections use 8-bit operands (see RLC_PREC_FP16C8). */
int rlc_umma_mode_prec(const rlc_critic* c, int act_mode, int precision);

/* ---- per-state reductions (rows a3, a4, a8, a9, a10) ------------------------------------ */
/* row.argsort()[::-1][:k] (ActorExpert.py:177, qt_opt_network.py:166): descending, ties -> larger
 * index first.  idx_out[B,k] int64; q_sel_out[B,k] or NULL;
 * optional elite gather (ActorExpert.py:178): actions[B,N,A] (or [N,A] if act_mode shared) ->
 * elites_out[B,k,A], pass NULL to skip. 1 <= k <= 64, k <= N. */
int rlc_reduce_topk(rlc_handle* h, const float* q, int B, int N, int k, int64_t* idx_out,
                    float* q_sel_out, const float* actions, int A, int act_mode,
                    float* elites_out, void* stream);
/* np.argmax / np.max / np.mean over axis 1 (optimal_q_network.py:156-159, ActorCritic.py:139,158).
 * any output may be NULL. argmax = first maximal index. */
int rlc_reduce_stats(rlc_handle* h, const float* q, int B, int N, int64_t* argmax_out,
                     float* max_out, float* mean_out, void* stream);
/* SQL soft value (sql_network.py:76-84): logsumexp_n q - log N + A log 2. */
int rlc_reduce_lse(rlc_handle* h, const float* q, int B, int N, int action_dim, float* v_out,
                   void* stream);
/* ForwardKL grid reduction (forwardkl_network.py:165-194). w[N], logp[B,N].
 * loss_b_out[B] = -sum_n w_n p_n logp; boltz_out[B,N] (p, may be NULL);
 * dlogp_out[B,N] = d(mean_b loss_b)/dlogp = -w_n p_n / B_total (may be NULL).
 * B_total: the global batch the mean is taken over (== B on one GPU, B*world when states are
 * sharded). */
int rlc_reduce_fkl(rlc_handle* h, const float* q, const float* w, const float* logp, int B,
                   int N, float entropy_scale, int B_total, float* loss_b_out, float* boltz_out,
                   float* dlogp_out, void* stream);
/* ReverseKL grid reduction (reversekl_network.py:181-190; hard: 197-203). v[B]. */
int rlc_reduce_rkl(rlc_handle* h, const float* q, const float* v, const float* w,
                   const float* logp, int B, int N, float entropy_scale, int hard, int B_total,
                   float* loss_b_out, float* dlogp_out, void* stream);

/* The same two reductions with PolicyNetwork.get_logprob (forwardkl_network.py:324-351; row a5)
 * evaluated in place from the policy head outputs mean[B,A], log_std[B,A] on the shared grid
 * grid[N,A] (tanh-Gaussian; for A > 1 std is the MVN covariance exactly as the reference passes it,
 * :350).  No [B,N] log-density tensor is read; the gradient of mean_b loss_b comes back as
 * dmean_out[B,A], dlog_std_out[B,A] (may be NULL); logp_out[B,N] is optional (NULL to skip).
 * 1 <= A <= 8. */
int rlc_reduce_fkl_policy(rlc_handle* h, const float* q, const float* w, const float* grid, int A,
                          float action_scale, const float* mean, const float* log_std, int B, int N,
                          float entropy_scale, int B_total, float* loss_b_out, float* dmean_out,
                          float* dlog_std_out, float* logp_out, void* stream);
int rlc_reduce_rkl_policy(rlc_handle* h, const float* q, const float* v, const float* w,
                          const float* grid, int A, float action_scale, const float* mean,
                          const float* log_std, int B, int N, float entropy_scale, int hard,
                          int B_total, float* loss_b_out, float* dmean_out, float* dlog_std_out,
                          float* logp_out, void* stream);

/* ---- CEM (rows a11, a12) ----------------------------------------------------------------
 * iterate_cem_multidim (qt_opt_network.py:132-175) + BoundedVarGaussianMixture refit
 * (utils/boundedvar_gaussian_mixture.py:10-75), all iterations in one launch, random draws
 * supplied: u0[B,N,A] in [0,1) (iter 0: a = amin + u0*(amax-amin)); noise[(iters-1),B,N,A] ~N(0,1);
 * comp_u[(iters-1),B,N] in [0,1) picks the mixture component.  T-mid critics only
 * (QT-Opt's critic).  amin/amax: A floats (device).  Outputs: weights[B,M], means[B,M,A],
 * vars[B,M,A], best_action[B,A] (= means[argmax weights], qt_opt_network.py:180),
 * elite_idx[iters,B,top_m] int64 (may be NULL). 1 <= num_modal <= 2, top_m <= 32, A <= 8. */
int rlc_cem(rlc_handle* h, const rlc_critic* c, const float* s, int B, int N, int iters,
            int top_m, int num_modal, const float* u0, const float* noise, const float* comp_u,
            const float* amin, const float* amax, float* weights_out, float* means_out,
            float* vars_out, float* best_action_out, int64_t* elite_idx_out, void* stream);
/* The refit alone: X[B,k,A] elites -> mixture; resp0[B,k,M] initial responsibilities or NULL
 * (NULL = deterministic farthest-point split). */
int rlc_gmm_refit(rlc_handle* h, const float* X, int B, int k, int A, int num_modal,
                  const float* resp0, float tol, int max_iter, float* weights_out,
                  float* means_out, float* vars_out, int32_t* n_iter_out, void* stream);

/* ---- backward (rows a14, a15, a16) -------------------------------------------------------- */
/* tf.gradients(q, action) on R stacked rows (ae_network.py:117,352-358; critic_network.py:58,
 * 169-183; sql_network.py:101-107).  s[R,S], a[R,A] -> dqda_out[R,A]; q_out[R] or NULL. */
int rlc_critic_grad_action(rlc_handle* h, const rlc_critic* c, const float* s, const float* a,
                           int R, float* dqda_out, float* q_out, void* stream);
/* Gradient of mean((y - Q(s,a))^2) over B rows wrt theta (critic_network.py:54-55,
 * forwardkl_network.py:133-140,199-201).  grad_out[numel theta]; loss_out[1]; q_out[B] or NULL.
 * B_total as in rlc_reduce_fkl (mean over the global batch; sum-allreduce the result). */
int rlc_critic_grads(rlc_handle* h, const rlc_critic* c, const float* s, const float* a,
                     const float* y, int B, int B_total, float* grad_out, float* loss_out,
                     float* q_out, void* stream);
/* One Adam step on a flat vector; step is the 1-based count.  When target != NULL also applies
 * the soft update target += tau (theta_new - target) (critic_network.py:29,197-198). */
int rlc_adam_step(rlc_handle* h, float* theta, const float* grad, float* m, float* v,
                  int64_t n, int step, float lr, float beta1, float beta2, float eps,
                  int variant, float* target, float tau, void* stream);
/* CUDA-graph-safe variant: the 1-based step count lives on the device in state_dev[0] (a 16-byte
 * device buffer, zero-initialised by the caller; words 1..2 are scratch for the derived factors) and
 * is incremented by the call, so a captured update advances on every replay. */
int rlc_adam_step_dev(rlc_handle* h, float* theta, const float* grad, float* m, float* v,
                      int64_t n, int32_t* state_dev, float lr, float beta1, float beta2, float eps,
                      int variant, float* target, float tau, void* stream);
int rlc_soft_update(rlc_handle* h, float* target, const float* online, int64_t n, float tau,
                    void* stream);

/* ---- B-row networks around the hot path (FKL / RKL update_network) -------------------------
 * Generic MLP  x[B,in] -> relu(FC(in,H1)) -> relu(FC(H1,H2)) -> FC(H2,O)  on B rows:
 * ValueNetwork (forwardkl_network.py:270-290, reversekl_network.py:290-312; O = 1) and
 * PolicyNetwork.forward (forwardkl_network.py:293-322; O = 2A, columns [mean_linear | log_std_linear],
 * the clamp of log_std is applied by the consumers below).
 * theta: [ W1 (in x H1) | b1 | W2 (H1 x H2) | b2 | W3 (H2 x O) | b3 (O) ]  all [in,out]-major. */
typedef struct rlc_mlp {
  int32_t in, H1, H2, O;
  const float* theta;
} rlc_mlp;
int64_t rlc_mlp_numel(int in, int H1, int H2, int O);
int rlc_mlp_offsets(int in, int H1, int H2, int O, int64_t off[6]);
/* floats of the caller-owned activation buffer `act` ([B,H1] | [B,H2] pre-activations) */
int64_t rlc_mlp_act_numel(int H1, int H2, int B);
/* out[B,O].  act: optional activation buffer to keep for rlc_mlp_grads (NULL = scratch). */
int rlc_mlp_forward(rlc_handle* h, const rlc_mlp* m, const float* x, int B, float* out, float* act,
                    void* stream);
/* Gradient wrt theta given dout[B,O] = dLoss/dout (loss.backward() of the value / policy losses,
 * forwardkl_network.py:199-209).  act from rlc_mlp_forward on the same x/theta, or NULL to
 * recompute.  grad_out[numel]; dx_out[B,in] or NULL. */
int rlc_mlp_grads(rlc_handle* h, const rlc_mlp* m, const float* x, const float* act,
                  const float* dout, int B, float* grad_out, float* dx_out, void* stream);

/* The GEMM the B-row layers above (torch.nn.Linear forward/backward on the minibatch, forwardkl_network.py:263-268,
 * 283-290,296-301 and their autograd) and rlc_critic_grads / rlc_critic_grad_action are built from, for parity tests and
 * callers with their own layer structure:  C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] (+ bias[N]), A := relu(A) at load when
 * relu_a, C *= (maskZ > 0) when maskZ (same shape as C, leading dimension ldz); row-major with leading dimensions;
 * trans_a: A is stored [K,M]; trans_b: B is stored [N,K].  split_k != 0 (weight gradients X^T G: trans_a, !trans_b, no
 * bias/mask, ldc == N) splits the long K = batch dimension over the grid and sums the slabs in a fixed order.
 * path: 0 = dispatcher (tcgen05 3xTF32 kernel for dense launches, fp32 CUDA-core tile kernel for small ones), 1 = CUDA
 * cores, 2 = tensor cores whatever the shape.  Both paths are fp32-class (<= ~1e-6 relative). */
int rlc_rows_gemm(rlc_handle* h, int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                  int ldb, float* C, int ldc, const float* bias, const float* maskZ, int ldz, int relu_a, float alpha,
                  int split_k, int path, void* stream);
/* Diagnostic: pin the dispatcher of every B-row GEMM issued by THIS host thread (rlc_critic_grads, rlc_mlp_*, ...):
 * -1 = default, 0 = CUDA cores only, 2 = tensor cores whatever the shape.  Returns the previous setting. */
int rlc_rows_gemm_force(int mode);
/* Same for the T-mid evaluation of large B x N stacks (rlc_critic_eval, RLC_TMID): -1 = default (tcgen05 tiles when the
 * stack fills the machine), 0 = fp32 CUDA cores only, 2 = tensor cores whenever the shape is supported. */
int rlc_tmid_tc_force(int mode);

/* PolicyNetwork.evaluate (forwardkl_network.py:303-322, reversekl_network.py:325-344) on the policy
 * head head[B,2A] = [mean_raw | log_std_raw]; eps[B,A] = the N(0,1) draws behind normal.sample()
 * (NULL = zeros, i.e. the mean action).  A == 1: Normal(mean, std); A > 1: MultivariateNormal with
 * covariance diag_embed(std), as the reference passes it (:346-351).
 * Outputs (any may be NULL): action[B,A] = tanh(z) * action_scale, logp[B], mean_out[B,A] =
 * tanh(mean) * action_scale, mu_raw_out[B,A] (forward()'s mean, contiguous), log_std_out[B,A]
 * (clamped), z_out[B,A]. */
int rlc_policy_evaluate(rlc_handle* h, const float* head, const float* eps, int B, int A,
                        float action_scale, float log_std_min, float log_std_max, float* action_out,
                        float* logp_out, float* mean_out, float* mu_raw_out, float* log_std_out,
                        float* z_out, void* stream);
/* Regression targets (forwardkl_network.py:137-150): y_q = r + gamma V_targ(s');
 * target_v = Q(s,a_new) - alpha logp ('sac', sac=1) or (r - alpha logp) + gamma V_targ(s') ('non_sac');
 * dv[B] = 2 (v - target_v) / B_total; v_loss[1] = mean (v - target_v)^2 (this rank's share).
 * dv_out/v_loss_out may be NULL (then q_new/logp/v are unused). */
int rlc_kl_targets(rlc_handle* h, const float* r, const float* gamma, const float* v_next,
                   const float* q_new, const float* logp, const float* v, int B, int B_total,
                   float entropy_scale, int sac, float* y_q_out, float* dv_out, float* v_loss_out,
                   void* stream);
/* Gradient wrt the raw policy head.  mode 0: chain dmean/dlog_std[B,A] (from
 * rlc_reduce_{fkl,rkl}_policy) through the log_std clamp.  mode 1 ('ll', reversekl_network.py:161-165)
 * and mode 2 ('hard_ll', :167-169): likelihood-ratio losses on the detached sample z[B,A] with
 * logp[B], q_new[B] = Q(s, a_new), v[B]; loss_out[1] receives the policy loss. */
int rlc_policy_head_grad(rlc_handle* h, const float* head, int B, int A, float log_std_min,
                         float log_std_max, int mode, const float* dmean, const float* dlog_std,
                         const float* z, const float* logp, const float* q_new, const float* v,
                         float entropy_scale, int B_total, float* dhead_out, float* loss_out,
                         void* stream);

/* ---- Actor-Expert actor side fused with the sampled-action path ----------------
 * sample_action (ae_network.py:461-496, ae_actor_network.py:310-341) with the draws supplied:
 * comp_u[B,N] in [0,1) picks the component as numpy's RandomState.choice(M, N, p=alpha) does
 * (cdf = cumsum(p)/sum; searchsorted(cdf, u, 'right'); equal_modal != 0: floor(u*M), alpha may be NULL),
 * normal[B,N,A] ~ N(0,1): action = clip(mean[idx] + sigma[idx] * normal, amin, amax) in float64, cast to
 * fp32.  The first n_uniform samples of every state are amin + (amax-amin) * uni_u[B,n_uniform,A]
 * (use_uniform_sampling, :489-493).  alpha[B,M], mean/sigma[B,M,A], M <= 8.  comp_out[B,N] int32 or NULL. */
int rlc_mixture_sample(rlc_handle* h, const float* alpha, const float* mean, const float* sigma, int B,
                       int M, int A, int N, int equal_modal, const float* comp_u, const float* normal,
                       const float* amin, const float* amax, int n_uniform, const float* uni_u,
                       float* actions_out, int32_t* comp_out, void* stream);
/* The expert step of ActorExpert.update_network (ActorExpert.py:162-181) in one launch: sample as above,
 * predict_q on the B*N stack through the (hoisted) T-mid critic, per-state argsort()[::-1][:k], elite
 * gather.  idx_out[B,k] int64; optional: actions_out[B,N,A], q_out[B,N], q_sel_out[B,k],
 * elites_out[B,k,A].  T-mid critics, A <= 8, k <= 64. */
int rlc_ae_expert_step(rlc_handle* h, const rlc_critic* c, const float* s, int B, int N, int k,
                       const float* alpha, const float* mean, const float* sigma, int M,
                       int equal_modal, const float* comp_u, const float* normal, const float* amin,
                       const float* amax, int n_uniform, const float* uni_u, float* actions_out,
                       float* q_out, int64_t* idx_out, float* q_sel_out, float* elites_out,
                       void* stream);
/* get_lossfunc (ae_network.py:262-278) on the elites actions[B,k,A] (state b's mixture against its k
 * elites; the reference repeats the state k times, ActorExpert.py:179-182):
 * loss[1] = mean over B_total*k rows of -log(clip(sum_m w_m prod_a N(y_a; mean, sigma), 1e-30, 1e30)),
 * w = alpha or 1/M; gradients of that loss wrt alpha[B,M], mean[B,M,A], sigma[B,M,A] (any output may be
 * NULL); nll_out[B,k] per-row values. */
int rlc_mixture_nll(rlc_handle* h, const float* alpha, const float* mean, const float* sigma,
                    const float* actions, int B, int M, int A, int k, int equal_modal, int B_total,
                    float* loss_out, float* nll_out, float* dalpha_out, float* dmean_out,
                    float* dsigma_out, void* stream);

/* ---- Soft-Q-learning SVGD actor step -----------------------------------------
 * action_gradients of sql_network.py:96-117 with adaptive_isotropic_gaussian_kernel
 * (utils/sql_kernel.py:7-69): fixed[B,Kf,A] / updated[B,Ku,A] = the two halves of the policy's particles
 * (:176-179); grad_log_p = dQ/da(s, fixed) + d/da sum log(1 - a^2 + eps) (:101-105); bandwidth
 * h = max(median/log(Kf), h_min), median = (Kf*Ku//2+1)-th largest squared distance;
 * grad_out[B,Ku,A] = mean_i(kappa grad_log_p + dkappa/dfixed) -- the grad_ys the reference back-propagates
 * through the policy network (:120).  dqda_scratch[B,Kf,A] is caller-owned scratch (holds dQ/da on return).
 * Optional: q_fixed_out[B,Kf], kappa_out[B,Kf,Ku], h_out[B].  Kf*Ku <= 4096, A <= 16. */
int rlc_svgd_action_grads(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* fixed,
                          int Kf, const float* updated, int Ku, float h_min, float eps,
                          float* dqda_scratch, float* grad_out, float* q_fixed_out, float* kappa_out,
                          float* h_out, void* stream);

/* ---- replay minibatch gather (row a17) ---------------------------------------------------
 * map(np.array, zip(*batch)) of ReplayBuffer.sample_batch (utils/replaybuffer.py:32-37) over a
 * device-resident struct-of-arrays ring: state[cap,S], action[cap,A], reward[cap],
 * next_state[cap,S], gamma[cap].  idx[B] = physical slots (int64).  Outputs are dense [B,...]. */
int rlc_replay_gather(rlc_handle* h, const float* state, const float* action,
                      const float* reward, const float* next_state, const float* gamma,
                      int64_t cap, int S, int A, const int64_t* idx, int B, float* s_out,
                      float* a_out, float* r_out, float* s2_out, float* g_out, void* stream);
/* Scatter n new transitions (host staging already copied to device, rows [n,...]) into ring
 * slots slot[n] (ReplayBuffer.add, utils/replaybuffer.py:25-27). */
int rlc_replay_scatter(rlc_handle* h, float* state, float* action, float* reward,
                       float* next_state, float* gamma, int64_t cap, int S, int A,
                       const int64_t* slot, int n, const float* s_in, const float* a_in,
                       const float* r_in, const float* s2_in, const float* g_in, void* stream);

/* Record layout of the same ring (the HBM-first variant of row a17; ReplayBuffer(layout="record")): ONE array
 * rec[cap, stride] of fixed-stride records [state S | action A | reward | next_state S | gamma | pad], with
 * stride = rlc_replay_rec_stride(S, A) = 2S+A+2 rounded up to 16 floats, so every record starts on a 64-byte
 * boundary and a random transition is a single contiguous read (utils/replaybuffer.py:25-37 stores one
 * Transition tuple per entry: same grouping).  Arguments otherwise as rlc_replay_gather / rlc_replay_scatter;
 * out-of-range slots are skipped.  rlc_replay_rec_stride is host-only and returns RLC_ERR_INVALID (< 0) on bad dims. */
int rlc_replay_rec_stride(int S, int A);
int rlc_replay_gather_rec(rlc_handle* h, const float* rec, int64_t cap, int stride, int S, int A,
                          const int64_t* idx, int B, float* s_out, float* a_out, float* r_out,
                          float* s2_out, float* g_out, void* stream);
int rlc_replay_scatter_rec(rlc_handle* h, float* rec, int64_t cap, int stride, int S, int A,
                           const int64_t* slot, int n, const float* s_in, const float* a_in,
                           const float* r_in, const float* s2_in, const float* g_in, void* stream);

/* Device-side minibatch index sampling: k distinct uniform indices in [0, n) from a
 * counter-based Philox stream keyed by (seed, counter) -- a pure function of its arguments, so a run is
 * reproducible, but NOT the numpy stream of RandomAccessQueue.sample_n_k (custom_collections.py:107-131; the
 * host sampler in rlcontrol_b200/replaybuffer.py reproduces that one and stays the default).
 * idx_out[k] = logical FIFO indices, slot_out[k] = ring slots (head + idx) % cap; either may be NULL.
 * Requires k <= 4096 and 3k < n (the reference's own condition for its rejection scheme), else
 * RLC_ERR_UNSUPPORTED. */
int rlc_replay_sample(rlc_handle* h, int64_t n, int k, uint64_t seed, uint64_t counter, int64_t head,
                      int64_t cap, int64_t* idx_out, int64_t* slot_out, void* stream);

/* ---- device-resident agent/environment loop ------------------------------------
 * Replaces the per-step Python of Experiment.run_episode_train / run_episode_eval (experiment.py:101-214),
 * BaseAgent.update (agents/base_agent.py:52-70) and the environments behind create_environment
 * (environments/environments.py:16-37) for the two environments the shipped configs name on this path:
 * gym 0.18.0 `Pendulum-v0` (requirements.txt:9; restated, gym is not vendored -> parity unpinned) and the
 * Bimodal1DEnv* bandits (environments.py:158-764; pinned on the reference classes).  Environment state, replay
 * ring and the cursors live in HBM; all randomness is drawn on the host from the reference's own streams and fed
 * in as tensors, a chunk of steps ahead.  Every call is a few threads of scalar work meant to sit inside a
 * captured graph between the policy forward pass and the update. */
#define RLC_ENV_PENDULUM 0
#define RLC_ENV_BIMODAL1D 1
#define RLC_ENV_MAX_S 8
typedef struct rlc_env {
  int kind, S, A;
  int episode_limit; /* EPISODE_STEPS_LIMIT (environments.py:52-59; 200 for Pendulum-v0, 1 for the bandits) */
  double p[8];       /* BIMODAL1D: maxima1, maxima2, stddev1, stddev2, height1, height2 */
} rlc_env;

/* Reset E environments from rows cursor[0]..cursor[0]+E-1 of reset_feed[feed_rows,2] (host-drawn internal states:
 * Pendulum (theta, thetadot) ~ U(-[pi,1],[pi,1]) from the env's np_random; bandits start at 0), write their
 * observations obs[E,S], zero ep_step[E] / ep_ret[E] / ep_done[E] (the last two may be NULL) and advance the
 * cursor by E (cursor may be NULL: rows 0..E-1). */
int rlc_env_reset(rlc_handle* h, const rlc_env* env, int E, const double* reset_feed, int64_t feed_rows,
                  int64_t* cursor, double* env_state, int* ep_step, double* ep_ret, int* ep_done, float* obs,
                  void* stream);
/* One evaluation step of E independent episodes (run_episode_eval, experiment.py:196-214): env.step(action[e]),
 * ep_ret += reward, ep_step += 1, obs <- next observation; an episode freezes once done or at the step limit. */
int rlc_env_step_eval(rlc_handle* h, const rlc_env* env, int E, double* env_state, int* ep_step, double* ep_ret,
                      int* ep_done, float* obs, const float* action, void* stream);
/* Append (ep_ret[E], ep_step[E]) as row cursor[0] of ret_log[log_rows,E] / steps_log[log_rows,E]; cursor[0] += 1. */
int rlc_eval_store(rlc_handle* h, int E, const double* ep_ret, const int* ep_step, int64_t* cursor,
                   int64_t log_rows, double* ret_log, int* steps_log, void* stream);
/* One training step of one environment: env.step(action[A]) from (env_state, obs[S]); the transition
 * (obs, action, reward, obs', done ? 0 : gamma) is appended to the replay ring (count = cur[1], head = cur[2];
 * FIFO eviction when full) unless the step was cut by the episode limit (experiment.py:127-134; the bandits are
 * exempt); reward_log[k] / flag_log[k] (bit0 done, bit1 truncated) with k = cur[0]; obs <- next observation, or on
 * done the observation of reset_feed row cur[3]++; cur[0]++, cur[4]++ (total steps).  cur = int64[8]. */
int rlc_env_step_train(rlc_handle* h, const rlc_env* env, int64_t* cur, double* env_state, int* ep_step,
                       float* obs, const float* action, const double* reset_feed, int64_t reset_rows,
                       float* rb_state, float* rb_action, float* rb_reward, float* rb_next_state, float* rb_gamma,
                       int64_t cap, float gamma, int64_t log_rows, double* reward_log, int* flag_log, int64_t rb_pitch,
                       void* stream);
/* rb_pitch (rlc_env_step_train, rlc_loop_step): 0 = the five rb_* pointers are struct-of-arrays rings; > 0 = they are the field
 * views of ONE record ring (rlc_replay_rec_stride floats per 64-byte-aligned record, rlc_replay_*_rec) and share this row
 * pitch in floats -- the default layout of ReplayBuffer and of the device-resident loop. */
/* Stage row k = (cur[0]-1) mod feed_rows (the row rlc_env_step_train just logged) of the host-drawn feeds into the fixed buffers a captured update reads: eps_act[A] <-
 * eps_act_feed[k] (sample_action's N(0,1) draws), eps_upd[B,A] <- eps_upd_feed[k] (pi.evaluate's draws inside
 * update_network), slots[B] <- (cur[2] + idx_feed[k,b]) % cap (RandomAccessQueue.sample_n_k's logical indices as
 * ring slots).  eps_upd_feed / idx_feed may be NULL (steps before learning starts). */
int rlc_loop_stage(rlc_handle* h, const int64_t* cur, int B, int A, int64_t feed_rows, const float* eps_act_feed,
                   const float* eps_upd_feed, const int* idx_feed, int64_t cap, float* eps_act, float* eps_upd,
                   int64_t* slots, void* stream);
/* rlc_env_step_train + rlc_loop_stage + rlc_replay_gather as ONE launch (the head of a captured training step):
 * env.step and the replay append, then this step's feeds are staged (eps_act[A], eps_upd[B,A]) and the minibatch
 * idx_feed[k,:] is gathered from the ring into s_out[B,S], a_out[B,A], r_out[B], s2_out[B,S], g_out[B].
 * idx_feed == NULL (steps before learning starts): only eps_act is staged. */
int rlc_loop_step(rlc_handle* h, const rlc_env* env, int64_t* cur, double* env_state, int* ep_step, float* obs,
                  const float* action, const double* reset_feed, int64_t reset_rows, float* rb_state, float* rb_action,
                  float* rb_reward, float* rb_next_state, float* rb_gamma, int64_t cap, float gamma, int64_t ring_rows,
                  double* reward_log, int* flag_log, int B, const float* eps_act_feed, const float* eps_upd_feed,
                  const int* idx_feed, float* eps_act, float* eps_upd, float* s_out, float* a_out, float* r_out,
                  float* s2_out, float* g_out, int64_t rb_pitch, void* stream);

/* ---- small-minibatch fast path of the ForwardKL / ReverseKL update (cfg1 / cfg5) -------------------
 * forwardkl_network.py:123-209 / reversekl_network.py:130-217 at B <= 64 rows: every B-row forward pass in ONE
 * launch (rlc_sb_forward) and every backward pass + Adam (+ Polyak) in ONE launch (rlc_sb_update), around the
 * B x N grid evaluation (rlc_critic_eval) and its reduction (rlc_reduce_{fkl,rkl}_policy).  theta layouts are those
 * of rlc_mlp / rlc_critic (T-in): [W1 (in x H1) | b1 | W2 | b2 | W3 (H2 x O) | b3]. */
#define RLC_SB_MAX_NETS 8
#define RLC_SB_MAX_B 64
#define RLC_SB_ROLE_DOUT 0 /* dLoss/dout given */
#define RLC_SB_ROLE_V 1    /* value regression:  target = (r - alpha logp) + gamma v_next  |  sac: q_new - alpha logp */
#define RLC_SB_ROLE_Q 2    /* critic regression: y = r + gamma v_next */
#define RLC_SB_ROLE_PI 3   /* policy head: (dmean, dlog_std) through the log_std clamp; loss = mean_b loss_b */
typedef struct rlc_sb_net {
  const float* theta;
  int inp, H1, H2, O;
  const float* x0; /* input row r = [x0[r / x0_div, 0:n0] | x1[r % x1_mod, 0:n1]], n0 + n1 == inp */
  int n0;
  const float* x1;
  int n1;
  int rows;           /* rows of this pass; 0 = the call's B.  With x0_div = x1_mod = N and rows = B*N this is the
                       * un-materialised B x N stack of row a1/a2 (x0 = s[B,S], x1 = grid[N,A], out = q[B,N]) */
  int x0_div, x1_mod; /* 0 = identity (row r of x0 / x1) */
  float* h1;          /* out, optional: post-ReLU activations [rows,H1] / [rows,H2] kept for rlc_sb_update */
  float* h2;
  float* out;         /* out [B,O] */
  float* w3_snapshot; /* out, optional [H2*O]: W3 as seen by this forward pass (rlc_sb_update reads it) */
  int* adam_state;    /* optional int32[4]: ++step and derive the bias-correction factors, as rlc_adam_step_dev */
  float lr, beta1, beta2;
  int adam_variant;
  int policy;         /* 1: O == 2A and PolicyNetwork.evaluate runs as the epilogue (as rlc_policy_evaluate) */
  const float* eps;   /* [B,A] or NULL (mean action) */
  float action_scale, log_std_min, log_std_max;
  float *action, *logp, *mean, *mu_raw, *log_std, *z; /* any may be NULL */
} rlc_sb_net;
typedef struct rlc_sb_train {
  float *theta, *m, *v; /* updated in place (torch/TF Adam as rlc_adam_step_dev) */
  const int* adam_state; /* factors written by the rlc_sb_forward of the same update */
  float beta1, beta2, eps;
  float* target;         /* optional Polyak target: target += tau (theta_new - target) */
  float tau;
  int inp, H1, H2, O;
  const float* x0;
  int n0;
  const float* x1;
  int n1;
  const float *h1, *h2, *out, *w3_snapshot; /* from rlc_sb_forward on the same x / theta */
  int role;
  const float* dout;                        /* ROLE_DOUT: [B,O] */
  const float *r, *gamma, *v_next, *logp, *q_new; /* ROLE_V / ROLE_Q operands, [B] each */
  const float *dmean, *dlog_std, *loss_b;   /* ROLE_PI: [B,A], [B,A], [B] (rlc_reduce_*_policy outputs) */
  float log_std_min, log_std_max, entropy_scale;
  int sac;
  float* loss_out;                          /* optional scalar */
} rlc_sb_train;
/* All forward passes of nets[0..n_nets) in one launch: B <= RLC_SB_MAX_B rows each unless a net sets `rows`
 * (<= 65536). */
int rlc_sb_forward(rlc_handle* h, const rlc_sb_net* nets, int n_nets, int B, void* stream);
/* All backward passes + optimiser steps, one launch; gradients are means over B_total rows.  Parameters change in
 * place; no gradient is written out (single-GPU path: the data-parallel update uses rlc_mlp_grads + all-reduce). */
int rlc_sb_update(rlc_handle* h, const rlc_sb_train* nets, int n_nets, int B, int B_total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLC_H_ */
