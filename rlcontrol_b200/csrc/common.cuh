// Shared internals of librlc.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rlc.h"

#define RLC_MAX_PACKS 8
#define RLC_MAX_RETIRED 64

// Pre-packed fp16/bf16 tensor-core operands of one theta (see critic_umma.cu).
struct rlc_pack {
  const float* theta;  // key
  int topology, S, A, H1, H2, prec;
  int ch;              // layer-1 chunk width the pack was built for (W1 rows are split per chunk)
  void* dev;           // packed blob
  size_t bytes;
  bool valid;
};

struct rlc_handle {
  int device;
  int sm_major, sm_minor, num_sms;
  size_t smem_optin;
  int64_t launches;
  // grow-only scratch
  void* ws;
  size_t ws_bytes;
  rlc_pack packs[RLC_MAX_PACKS];
  int pack_rr;
  int* err_flag;  // device int raised by bounded waits in the tcgen05 kernel
  // Blocks replaced by a larger one (workspace, operand packs) are RETIRED, not freed: a CUDA graph captured earlier
  // has their addresses baked in and keeps replaying into them.  Freed in rlc_destroy.
  void* retired[RLC_MAX_RETIRED];
  int n_retired;
};

// Replace a device block: the old one is kept alive for captured graphs (falls back to synchronise + free when the
// retirement list is full).
int rlc_retire_block(rlc_handle* h, void* old_block);

extern thread_local char g_rlc_cuda_err[512];

static inline int rlc_cuda_fail(cudaError_t e, const char* file, int line) {
  snprintf(g_rlc_cuda_err, sizeof(g_rlc_cuda_err), "%s:%d: %s (%s)", file, line,
           cudaGetErrorString(e), cudaGetErrorName(e));
  return RLC_ERR_CUDA;
}

#define RLC_CUDA(expr)                                             \
  do {                                                             \
    cudaError_t e__ = (expr);                                      \
    if (e__ != cudaSuccess) return rlc_cuda_fail(e__, __FILE__, __LINE__); \
  } while (0)

#define RLC_LAUNCH_CHECK(h)                                        \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return rlc_cuda_fail(e__, __FILE__, __LINE__); \
    (h)->launches++;                                               \
  } while (0)

#define RLC_REQUIRE(cond)              \
  do {                                 \
    if (!(cond)) return RLC_ERR_INVALID; \
  } while (0)

// Scratch of at least `bytes` (256-B aligned). Grow-only; growing synchronises the device, so
// steady-state steps never allocate.
int rlc_workspace(rlc_handle* h, size_t bytes, void** out);

struct ThetaView {
  int in1, in2;
  int64_t oW1, ob1, oW2, ob2, ow3, ob3, numel;
};

static inline __host__ __device__ ThetaView theta_view(int topology, int S, int A, int H1, int H2) {
  ThetaView t;
  t.in1 = (topology == RLC_TIN) ? S + A : S;
  t.in2 = (topology == RLC_TIN) ? H1 : H1 + A;
  t.oW1 = 0;
  t.ob1 = t.oW1 + (int64_t)t.in1 * H1;
  t.oW2 = t.ob1 + H1;
  t.ob2 = t.oW2 + (int64_t)t.in2 * H2;
  t.ow3 = t.ob2 + H2;
  t.ob3 = t.ow3 + H2;
  t.numel = t.ob3 + 1;
  return t;
}

static inline bool critic_ok(const rlc_critic* c) {
  return c && c->theta && (c->topology == RLC_TIN || c->topology == RLC_TMID) && c->S >= 1 &&
         c->A >= 1 && c->H1 >= 1 && c->H2 >= 1 && c->S <= 4096 && c->A <= 256 &&
         c->H1 <= 2048 && c->H2 <= 2048 && ((c->smin == nullptr) == (c->smax == nullptr));
}

// ---- entry points implemented across the .cu files (internal linkage across TUs) ----------
int rlc_eval_fp32(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                  int act_mode, float* q_out, cudaStream_t st);
int rlc_eval_umma(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                  int act_mode, int prec, float* q_out, cudaStream_t st);
bool rlc_umma_supported(const rlc_handle* h, const rlc_critic* c, int B, int N);
bool rlc_umma3_supported(const rlc_handle* h, const rlc_critic* c, int prec);   // split modes (FP16X3 / FP16C8), shared grids
// p[B,H2] = relu(clip(s) W1 + b1) W2[:H1] + b2 for a T-mid critic (state-only hoisted term)
// scratch (optional, rlc_tmid_state_scratch_floats(c, B) floats): lets dense state batches run as tensor-core GEMMs
int rlc_tmid_state_term(rlc_handle* h, const rlc_critic* c, const float* s, int B, float* p_out,
                        cudaStream_t st, float* scratch = nullptr);
size_t rlc_tmid_state_scratch_floats(const rlc_critic* c, int B);

// T-mid rows of a B x N stack on the tensor cores (tmid_rows_tc.cu); p = state terms [B,H2]
bool rlc_tmid_tc_ok(const rlc_handle* h, const rlc_critic* c, long long R, int N);
int rlc_tmid_rows_tc(rlc_handle* h, const rlc_critic* c, const float* p, const float* a, int act_per_state, int B, int N,
                     float* q_out, cudaStream_t st);

// U[k][n] = atanh(a_nk / scale), J[n] = sum_k log(1 - (a_nk/scale)^2 + 1e-6): grid-only terms of the policy log-density (reduce.cu)
int rlc_launch_grid_logterms(rlc_handle* h, const float* grid, int N, int A, float action_scale, float* U, float* J,
                             cudaStream_t st);
// fused evaluation + per-state policy reduction on the split tensor kernels (critic_umma_grid3.cuh); RLC_ERR_UNSUPPORTED when
// the shape does not qualify (the caller then composes rlc_critic_eval + rlc_reduce_*_policy)
struct rlc_fuse_args {
  int mode;                   // 1 = ForwardKL, 2 = ReverseKL
  int A;
  const float *w, *grid, *mean, *log_std, *v;
  float action_scale, alpha;
  int B_total;
  float *loss_b, *dmean, *dlog_std;
};
int rlc_eval_umma_grid3_fused(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N, int prec,
                              float* q_out, const rlc_fuse_args* f, cudaStream_t st);

// dQ/da on R stacked rows, state row r/rep (critic_fp32.cu)
int rlc_critic_grad_action_rep(rlc_handle* h, const rlc_critic* c, const float* s, int rep,
                               const float* a, long long R, float* dqda_out, float* q_out,
                               cudaStream_t st);

// warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
