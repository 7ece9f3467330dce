// N2: device-resident agent/environment loop.  The reference runs one OS process per sweep
// INDEX and, per environment step, crosses Python -> gym -> replay list -> torch (experiment.py:101-161,
// agents/base_agent.py:52-70).  Here the environment state, the replay ring and the cursors that order them
// live in HBM, so one training step (env.step -> ReplayBuffer.add -> sample_batch -> update_network ->
// update_target_network -> sample_action) is ONE captured CUDA graph with no host round trip; everything
// random (normal draws, minibatch indices, reset states) is drawn on the host FROM THE REFERENCE'S OWN STREAMS,
// a chunk of steps ahead, and fed in as tensors.  These kernels are the glue of that graph: a few threads of
// fp64 scalar work each (latency-bound; nothing to tile).
//
// Environments (kind):
//   RLC_ENV_PENDULUM   gym 0.18.0 (requirements.txt:9) classic_control/pendulum.py `Pendulum-v0` behind TimeLimit(200):
//                      max_speed 8, max_torque 2, dt .05, g 10, m = l = 1; float64 arithmetic (numpy < 2 promotes the
//                      float32 action scalar to float64).  gym is not vendored in the reference: restated from the
//                      published source, parity unpinned.
//   RLC_ENV_BIMODAL1D  environments/environments.py:158-764 -- the seven one-step bandits; `reward_func` evaluates
//                      -0.5*((a-m)/sd)**2 in float32 (a is a float32 ARRAY) and math.exp in double.  Pinned on the
//                      reference classes (tests/golden/bimodal_env.npz).
#include "common.cuh"

#define PI_D 3.141592653589793

struct EnvOut {
  double s0, s1;   // next internal state
  double reward;
  int done;        // environment's own terminal flag (before the episode step limit)
};

__device__ __forceinline__ double angle_normalize(double x) {
  // ((x + pi) % (2 pi)) - pi with Python's sign-of-divisor modulo
  double m = fmod(x + PI_D, 2.0 * PI_D);
  if (m < 0.0) m += 2.0 * PI_D;
  return m - PI_D;
}

__device__ __forceinline__ EnvOut env_dynamics(const rlc_env& env, double s0, double s1, const float* a) {
  EnvOut o;
  if (env.kind == RLC_ENV_PENDULUM) {
    const double th = s0, thdot = s1;
    double u = (double)a[0];
    u = fmin(fmax(u, -2.0), 2.0);
    const double an = angle_normalize(th);
    const double costs = an * an + .1 * (thdot * thdot) + .001 * (u * u);
    double newthdot = thdot + (-3.0 * 10.0 / (2.0 * 1.0) * sin(th + PI_D) + 3.0 / (1.0 * 1.0) * u) * .05;
    const double newth = th + newthdot * .05;
    newthdot = fmin(fmax(newthdot, -8.0), 8.0);
    o.s0 = newth;
    o.s1 = newthdot;
    o.reward = -costs;
    o.done = 0;
  } else {  // RLC_ENV_BIMODAL1D: state + action, always terminal
    const float af = a[0];
    const float m1 = (float)env.p[0], m2 = (float)env.p[1], sd1 = (float)env.p[2], sd2 = (float)env.p[3];
    const float t1 = __fdiv_rn(__fsub_rn(af, m1), sd1), t2 = __fdiv_rn(__fsub_rn(af, m2), sd2);
    const float e1 = __fmul_rn(-0.5f, __fmul_rn(t1, t1)), e2 = __fmul_rn(-0.5f, __fmul_rn(t2, t2));
    o.reward = env.p[4] * exp((double)e1) + env.p[5] * exp((double)e2);
    o.s0 = s0 + (double)af;
    o.s1 = 0.0;
    o.done = 1;
  }
  return o;
}

__device__ __forceinline__ void env_observe(const rlc_env& env, double s0, double s1, float* obs) {
  if (env.kind == RLC_ENV_PENDULUM) {
    obs[0] = (float)cos(s0);
    obs[1] = (float)sin(s0);
    obs[2] = (float)s1;
  } else {
    obs[0] = (float)s0;
  }
}

// reset E environments from consecutive rows of the host-drawn feed (row = cursor[0] + e); clears the per-episode
// accumulators.  The last thread's-eye view: cursor[0] += E.
__global__ void k_env_reset(rlc_env env, int E, const double* __restrict__ feed, long long feed_rows,
                            long long* __restrict__ cursor, double* __restrict__ state,
                            int* __restrict__ ep_step, double* __restrict__ ep_ret, int* __restrict__ ep_done,
                            float* __restrict__ obs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const long long base = cursor ? cursor[0] : 0;
  if (e < E) {
    long long row = base + e;
    if (row >= feed_rows) row = feed_rows - 1;  // host sizes the feed; never read past it
    const double s0 = feed[row * 2], s1 = feed[row * 2 + 1];
    state[e * 2] = s0;
    state[e * 2 + 1] = s1;
    ep_step[e] = 0;
    if (ep_ret) ep_ret[e] = 0.0;
    if (ep_done) ep_done[e] = 0;
    env_observe(env, s0, s1, obs + (long long)e * env.S);
  }
  __syncthreads();
  if (cursor && e == 0) cursor[0] = base + E;   // single CTA (E <= 1024)
}

// Evaluation step of E independent episodes (run_episode_eval, experiment.py:196-214): frozen once done.
__global__ void k_env_step_eval(rlc_env env, int E, double* __restrict__ state, int* __restrict__ ep_step,
                                double* __restrict__ ep_ret, int* __restrict__ ep_done, float* __restrict__ obs,
                                const float* __restrict__ action) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || ep_done[e]) return;
  const EnvOut o = env_dynamics(env, state[e * 2], state[e * 2 + 1], action + (long long)e * env.A);
  const int n = ep_step[e] + 1;
  ep_step[e] = n;
  ep_ret[e] += o.reward;
  state[e * 2] = o.s0;
  state[e * 2 + 1] = o.s1;
  env_observe(env, o.s0, o.s1, obs + (long long)e * env.S);
  if (o.done || n >= env.episode_limit) ep_done[e] = 1;
}

__global__ void k_eval_store(int E, const double* __restrict__ ep_ret, const int* __restrict__ ep_step,
                             long long* __restrict__ cursor, long long log_rows, double* __restrict__ ret_log,
                             int* __restrict__ steps_log) {
  const int e = threadIdx.x;
  const long long row = cursor[0];
  if (e < E && row < log_rows) {
    ret_log[row * E + e] = ep_ret[e];
    steps_log[row * E + e] = ep_step[e];
  }
  __syncthreads();
  if (e == 0) cursor[0] = row + 1;
}

// One training step of ONE environment (run_episode_train body, experiment.py:118-142 + BaseAgent.update,
// base_agent.py:52-58): env.step(action) -> episode bookkeeping -> ReplayBuffer.add unless the step was cut by the
// episode limit -> next observation (or the next episode's reset observation) into obs.
// cur: [0] step row k in this chunk's feeds/logs  [1] replay count  [2] replay head  [3] reset-feed row
//      [4] total steps taken
__device__ __forceinline__ void env_step_train_body(const rlc_env& env, long long* __restrict__ cur, double* __restrict__ state,
                                 int* __restrict__ ep_step, float* __restrict__ obs,
                                 const float* __restrict__ action, const double* __restrict__ reset_feed,
                                 long long reset_rows, float* __restrict__ rb_state, float* __restrict__ rb_action,
                                 float* __restrict__ rb_reward, float* __restrict__ rb_next,
                                 float* __restrict__ rb_gamma, long long cap, float gamma, long long log_rows,
                                 double* __restrict__ reward_log, int* __restrict__ flag_log, long long rb_pitch) {
  // rb_pitch: 0 = five struct-of-arrays rings (row pitch S, A, 1, S, 1); > 0 = the RECORD ring (replay.cu): one array of
  // fixed-stride 64-byte-aligned records, the five pointers are its field views and share this row pitch
  const long long ps = rb_pitch ? rb_pitch : env.S, pa = rb_pitch ? rb_pitch : env.A, p1 = rb_pitch ? rb_pitch : 1;
  const long long k = cur[0] % log_rows;
  const EnvOut o = env_dynamics(env, state[0], state[1], action);
  const int n = ep_step[0] + 1;
  const int done = o.done || n >= env.episode_limit;
  // experiment.py:127-134: a step that ends BY the limit is not learned from, except in the one-step bandits
  const int truncated = (env.kind != RLC_ENV_BIMODAL1D) && done && n == env.episode_limit;
  float obs_n[RLC_ENV_MAX_S];
  env_observe(env, o.s0, o.s1, obs_n);
  if (!truncated) {
    long long count = cur[1], head = cur[2], slot;
    if (count < cap) {
      slot = (head + count) % cap;
      cur[1] = count + 1;
    } else {  // RandomAccessQueue.append with maxlen evicts the oldest (custom_collections.py:85-88)
      slot = head;
      cur[2] = (head + 1) % cap;
    }
    for (int i = 0; i < env.S; ++i) {
      rb_state[slot * ps + i] = obs[i];
      rb_next[slot * ps + i] = obs_n[i];
    }
    for (int i = 0; i < env.A; ++i) rb_action[slot * pa + i] = action[i];
    rb_reward[slot * p1] = (float)o.reward;
    rb_gamma[slot * p1] = done ? 0.f : gamma;   // is_terminal -> transition gamma 0.0 (base_agent.py:54-57)
  }
  reward_log[k] = o.reward;
  flag_log[k] = done | (truncated << 1);
  if (done) {
    long long row = cur[3];
    cur[3] = row + 1;
    if (row >= reset_rows) row = reset_rows - 1;
    const double s0 = reset_feed[row * 2], s1 = reset_feed[row * 2 + 1];
    state[0] = s0;
    state[1] = s1;
    ep_step[0] = 0;
    env_observe(env, s0, s1, obs);
  } else {
    state[0] = o.s0;
    state[1] = o.s1;
    ep_step[0] = n;
    for (int i = 0; i < env.S; ++i) obs[i] = obs_n[i];
  }
  cur[0] = (k + 1) % log_rows;   // feeds and logs are a ring: the host fills one half while the other is in flight
  cur[4] += 1;
}

__global__ void k_env_step_train(rlc_env env, long long* __restrict__ cur, double* __restrict__ state,
                                 int* __restrict__ ep_step, float* __restrict__ obs,
                                 const float* __restrict__ action, const double* __restrict__ reset_feed,
                                 long long reset_rows, float* __restrict__ rb_state, float* __restrict__ rb_action,
                                 float* __restrict__ rb_reward, float* __restrict__ rb_next,
                                 float* __restrict__ rb_gamma, long long cap, float gamma, long long log_rows,
                                 double* __restrict__ reward_log, int* __restrict__ flag_log, long long rb_pitch) {
  if (threadIdx.x != 0) return;
  env_step_train_body(env, cur, state, ep_step, obs, action, reset_feed, reset_rows, rb_state, rb_action, rb_reward, rb_next,
                      rb_gamma, cap, gamma, log_rows, reward_log, flag_log, rb_pitch);
}

// The head of a training step as ONE launch (one CTA): env.step + replay append by thread 0, then -- behind a block
// barrier, which also orders thread 0's global writes for the rest of the CTA -- the feeds of this step are staged and
// the minibatch is gathered from the ring (the row just appended may be among the sampled ones).  Three launches of a
// few microseconds each became one; inside a captured step every launch boundary is ~2 us of dependent latency.
struct LoopStepArgs {
  rlc_env env;
  long long* cur;
  double* state;
  int* ep_step;
  float* obs;
  const float* action;
  const double* reset_feed;
  long long reset_rows;
  float *rb_state, *rb_action, *rb_reward, *rb_next, *rb_gamma;
  long long rb_pitch;        // 0 = struct-of-arrays rings, > 0 = record ring row pitch (floats)
  long long cap;
  float gamma;
  long long ring_rows;
  double* reward_log;
  int* flag_log;
  int B;
  const float *eps_act_feed, *eps_upd_feed;
  const int* idx_feed;
  float *eps_act, *eps_upd;
  float *s_out, *a_out, *r_out, *s2_out, *g_out;
};

__global__ void __launch_bounds__(256) k_loop_step(const __grid_constant__ LoopStepArgs p) {
  const int tid = threadIdx.x;
  if (tid == 0)
    env_step_train_body(p.env, p.cur, p.state, p.ep_step, p.obs, p.action, p.reset_feed, p.reset_rows, p.rb_state,
                        p.rb_action, p.rb_reward, p.rb_next, p.rb_gamma, p.cap, p.gamma, p.ring_rows, p.reward_log,
                        p.flag_log, p.rb_pitch);
  __threadfence_block();
  __syncthreads();
  const int S = p.env.S, A = p.env.A, B = p.B;
  const long long k = (p.cur[0] % p.ring_rows + p.ring_rows - 1) % p.ring_rows;   // the row thread 0 just logged
  const long long head = p.cur[2], count = p.cur[1];
  if (tid < A) p.eps_act[tid] = p.eps_act_feed[k * A + tid];
  if (!p.idx_feed) return;                   // steps before learning starts: nothing to sample
  for (int t = tid; t < B * A; t += blockDim.x) p.eps_upd[t] = p.eps_upd_feed[k * B * A + t];
  const int W = 2 * S + A + 2;               // floats per gathered transition
  const long long ps = p.rb_pitch ? p.rb_pitch : S, pa = p.rb_pitch ? p.rb_pitch : A, p1 = p.rb_pitch ? p.rb_pitch : 1;
  for (int t = tid; t < B * W; t += blockDim.x) {
    const int b = t / W, f = t % W;
    long long i = p.idx_feed[k * B + b];
    if (i < 0) i = 0;
    if (count > 0 && i >= count) i = count - 1;
    const long long slot = (head + i) % p.cap;
    if (f < S) p.s_out[b * S + f] = p.rb_state[slot * ps + f];
    else if (f < 2 * S) p.s2_out[b * S + (f - S)] = p.rb_next[slot * ps + (f - S)];
    else if (f < 2 * S + A) p.a_out[b * A + (f - 2 * S)] = p.rb_action[slot * pa + (f - 2 * S)];
    else if (f == 2 * S + A) p.r_out[b] = p.rb_reward[slot * p1];
    else p.g_out[b] = p.rb_gamma[slot * p1];
  }
}

// Stage step k = cur[0]-1 of the host-drawn feeds into the fixed buffers the captured update reads: the N(0,1)
// draws of sample_action ([K,A]) and of the update's pi.evaluate ([K,B,A]), and the minibatch's logical FIFO
// indices ([K,B], RandomAccessQueue.sample_n_k) turned into ring slots (head + i) % cap.
__global__ void k_loop_stage(const long long* __restrict__ cur, int B, int A, long long feed_rows,
                             const float* __restrict__ eps_act_feed, const float* __restrict__ eps_upd_feed,
                             const int* __restrict__ idx_feed, long long cap, float* __restrict__ eps_act,
                             float* __restrict__ eps_upd, long long* __restrict__ slots) {
  const long long k = (cur[0] % feed_rows + feed_rows - 1) % feed_rows;   // the row k_env_step_train just logged
  const long long head = cur[2], count = cur[1];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < A) eps_act[t] = eps_act_feed[k * A + t];
  if (eps_upd_feed && t < B * A) eps_upd[t] = eps_upd_feed[k * B * A + t];
  if (idx_feed && t < B) {
    long long i = idx_feed[k * B + t];
    if (i < 0) i = 0;
    if (count > 0 && i >= count) i = count - 1;
    slots[t] = (head + i) % cap;
  }
}

static inline bool env_ok(const rlc_env* env) {
  if (!env || env->S < 1 || env->S > RLC_ENV_MAX_S || env->A < 1 || env->episode_limit < 1) return false;
  if (env->kind == RLC_ENV_PENDULUM) return env->S == 3 && env->A == 1;
  if (env->kind == RLC_ENV_BIMODAL1D) return env->S == 1 && env->A == 1 && env->p[2] > 0 && env->p[3] > 0;
  return false;
}

extern "C" int rlc_env_reset(rlc_handle* h, const rlc_env* env, int E, const double* reset_feed,
                             int64_t feed_rows, int64_t* cursor, double* env_state, int* ep_step,
                             double* ep_ret, int* ep_done, float* obs, void* stream) {
  RLC_REQUIRE(h && env_ok(env) && E >= 1 && E <= 1024 && reset_feed && feed_rows >= 1 && env_state && ep_step && obs);
  k_env_reset<<<1, 1024, 0, (cudaStream_t)stream>>>(*env, E, reset_feed, feed_rows, (long long*)cursor, env_state,
                                                    ep_step, ep_ret, ep_done, obs);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_env_step_eval(rlc_handle* h, const rlc_env* env, int E, double* env_state, int* ep_step,
                                 double* ep_ret, int* ep_done, float* obs, const float* action, void* stream) {
  RLC_REQUIRE(h && env_ok(env) && E >= 1 && env_state && ep_step && ep_ret && ep_done && obs && action);
  k_env_step_eval<<<(E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*env, E, env_state, ep_step, ep_ret, ep_done,
                                                                     obs, action);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_eval_store(rlc_handle* h, int E, const double* ep_ret, const int* ep_step, int64_t* cursor,
                              int64_t log_rows, double* ret_log, int* steps_log, void* stream) {
  RLC_REQUIRE(h && E >= 1 && E <= 1024 && ep_ret && ep_step && cursor && log_rows >= 1 && ret_log && steps_log);
  k_eval_store<<<1, 1024, 0, (cudaStream_t)stream>>>(E, ep_ret, ep_step, (long long*)cursor, log_rows, ret_log,
                                                     steps_log);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_env_step_train(rlc_handle* h, const rlc_env* env, int64_t* cur, double* env_state,
                                  int* ep_step, float* obs, const float* action, const double* reset_feed,
                                  int64_t reset_rows, float* rb_state, float* rb_action, float* rb_reward,
                                  float* rb_next_state, float* rb_gamma, int64_t cap, float gamma,
                                  int64_t log_rows, double* reward_log, int* flag_log, int64_t rb_pitch, void* stream) {
  RLC_REQUIRE(h && env_ok(env) && cur && env_state && ep_step && obs && action && reset_feed && reset_rows >= 1);
  RLC_REQUIRE(rb_state && rb_action && rb_reward && rb_next_state && rb_gamma && cap >= 1 && log_rows >= 1 &&
              reward_log && flag_log && rb_pitch >= 0);
  k_env_step_train<<<1, 32, 0, (cudaStream_t)stream>>>(*env, (long long*)cur, env_state, ep_step, obs, action,
                                                       reset_feed, reset_rows, rb_state, rb_action, rb_reward,
                                                       rb_next_state, rb_gamma, cap, gamma, log_rows, reward_log,
                                                       flag_log, (long long)rb_pitch);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_loop_stage(rlc_handle* h, const int64_t* cur, int B, int A, int64_t feed_rows,
                              const float* eps_act_feed, const float* eps_upd_feed, const int* idx_feed,
                              int64_t cap, float* eps_act, float* eps_upd, int64_t* slots, void* stream) {
  RLC_REQUIRE(h && cur && B >= 0 && A >= 1 && feed_rows >= 1 && eps_act_feed && eps_act && cap >= 1);
  RLC_REQUIRE((!eps_upd_feed || eps_upd) && (!idx_feed || slots));
  const int n = (B * A > A ? B * A : A);
  k_loop_stage<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const long long*)cur, B, A, feed_rows, eps_act_feed,
                                                                  eps_upd_feed, idx_feed, cap, eps_act, eps_upd,
                                                                  (long long*)slots);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_loop_step(rlc_handle* h, const rlc_env* env, int64_t* cur, double* env_state, int* ep_step, float* obs,
                             const float* action, const double* reset_feed, int64_t reset_rows, float* rb_state,
                             float* rb_action, float* rb_reward, float* rb_next_state, float* rb_gamma, int64_t cap,
                             float gamma, int64_t ring_rows, double* reward_log, int* flag_log, int B,
                             const float* eps_act_feed, const float* eps_upd_feed, const int* idx_feed, float* eps_act,
                             float* eps_upd, float* s_out, float* a_out, float* r_out, float* s2_out, float* g_out,
                             int64_t rb_pitch, void* stream) {
  RLC_REQUIRE(h && env_ok(env) && cur && env_state && ep_step && obs && action && reset_feed && reset_rows >= 1);
  RLC_REQUIRE(rb_state && rb_action && rb_reward && rb_next_state && rb_gamma && cap >= 1 && ring_rows >= 1 && reward_log &&
              flag_log && eps_act_feed && eps_act && B >= 0 && rb_pitch >= 0);
  RLC_REQUIRE(!idx_feed || (eps_upd_feed && eps_upd && s_out && a_out && r_out && s2_out && g_out && B >= 1));
  LoopStepArgs p;
  p.env = *env; p.cur = (long long*)cur; p.state = env_state; p.ep_step = ep_step; p.obs = obs; p.action = action;
  p.reset_feed = reset_feed; p.reset_rows = reset_rows; p.rb_state = rb_state; p.rb_action = rb_action;
  p.rb_reward = rb_reward; p.rb_next = rb_next_state; p.rb_gamma = rb_gamma; p.cap = cap; p.gamma = gamma;
  p.rb_pitch = rb_pitch;
  p.ring_rows = ring_rows; p.reward_log = reward_log; p.flag_log = flag_log; p.B = B; p.eps_act_feed = eps_act_feed;
  p.eps_upd_feed = eps_upd_feed; p.idx_feed = idx_feed; p.eps_act = eps_act; p.eps_upd = eps_upd; p.s_out = s_out;
  p.a_out = a_out; p.r_out = r_out; p.s2_out = s2_out; p.g_out = g_out;
  k_loop_step<<<1, 256, 0, (cudaStream_t)stream>>>(p);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
