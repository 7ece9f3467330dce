// Handle lifecycle, workspace and the evaluation dispatcher of librlc.so.
#include <new>

#include "common.cuh"

thread_local char g_rlc_cuda_err[512] = {0};

extern "C" int rlc_version(void) { return RLC_VERSION; }

extern "C" const char* rlc_status_string(int status) {
  switch (status) {
    case RLC_OK: return "ok";
    case RLC_ERR_INVALID: return "invalid argument (shape, null pointer or unsupported combination)";
    case RLC_ERR_ARCH: return "device is not sm_100: the tensor-core path needs a B200";
    case RLC_ERR_ALLOC: return "workspace allocation failed";
    case RLC_ERR_CUDA: return "CUDA runtime error";
    case RLC_ERR_UNSUPPORTED: return "shape not supported by this kernel";
    default: return "unknown status";
  }
}

extern "C" const char* rlc_last_cuda_error(void) { return g_rlc_cuda_err; }

extern "C" int rlc_create(rlc_handle** out, int device) {
  RLC_REQUIRE(out);
  *out = nullptr;
  int count = 0;
  RLC_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return RLC_ERR_INVALID;
  cudaDeviceProp prop;
  RLC_CUDA(cudaGetDeviceProperties(&prop, device));
  rlc_handle* h = new (std::nothrow) rlc_handle();
  if (!h) return RLC_ERR_ALLOC;
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->sm_major = prop.major;
  h->sm_minor = prop.minor;
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  RLC_CUDA(cudaSetDevice(device));
  if (cudaMalloc((void**)&h->err_flag, 256) != cudaSuccess) {
    (void)cudaGetLastError();
    delete h;
    return RLC_ERR_ALLOC;
  }
  RLC_CUDA(cudaMemset(h->err_flag, 0, 256));
  *out = h;
  return RLC_OK;
}

extern "C" int rlc_destroy(rlc_handle* h) {
  if (!h) return RLC_OK;
  if (h->ws) cudaFree(h->ws);
  if (h->err_flag) cudaFree(h->err_flag);
  for (int i = 0; i < RLC_MAX_PACKS; ++i)
    if (h->packs[i].dev) cudaFree(h->packs[i].dev);
  for (int i = 0; i < h->n_retired; ++i) cudaFree(h->retired[i]);
  delete h;
  return RLC_OK;
}

extern "C" int64_t rlc_launch_count(const rlc_handle* h) { return h ? h->launches : 0; }

int rlc_retire_block(rlc_handle* h, void* old_block) {
  if (!old_block) return RLC_OK;
  if (h->n_retired < RLC_MAX_RETIRED) {
    h->retired[h->n_retired++] = old_block;
    return RLC_OK;
  }
  RLC_CUDA(cudaDeviceSynchronize());     // list full: nothing older than 64 growths is expected to be replayed
  cudaFree(old_block);
  return RLC_OK;
}

// Scratch of at least `bytes`.  Grow-only and geometric (>= 1.5x), so a handle retires a handful of blocks at most.
// The outgrown block stays allocated: captured graphs (kl_networks, steps, device_loop) that ran on this handle before
// the growth keep their baked-in pointer valid; eager calls move on to the new block.
int rlc_workspace(rlc_handle* h, size_t bytes, void** out) {
  if (bytes > h->ws_bytes) {
    const int rc = rlc_retire_block(h, h->ws);
    if (rc) return rc;
    h->ws = nullptr;
    size_t want = bytes + (bytes >> 2) + 4096;
    if (want < h->ws_bytes + (h->ws_bytes >> 1)) want = h->ws_bytes + (h->ws_bytes >> 1);
    h->ws_bytes = 0;
    cudaError_t e = cudaMalloc(&h->ws, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return RLC_ERR_ALLOC;
    }
    h->ws_bytes = want;
  }
  *out = h->ws;
  return RLC_OK;
}

extern "C" int64_t rlc_theta_numel(int topology, int S, int A, int H1, int H2) {
  if ((topology != RLC_TIN && topology != RLC_TMID) || S < 1 || A < 1 || H1 < 1 || H2 < 1) return -1;
  return theta_view(topology, S, A, H1, H2).numel;
}

extern "C" int rlc_theta_offsets(int topology, int S, int A, int H1, int H2, int64_t off[6]) {
  RLC_REQUIRE(off && (topology == RLC_TIN || topology == RLC_TMID) && S >= 1 && A >= 1 && H1 >= 1 && H2 >= 1);
  const ThetaView t = theta_view(topology, S, A, H1, H2);
  off[0] = t.oW1; off[1] = t.ob1; off[2] = t.oW2; off[3] = t.ob2; off[4] = t.ow3; off[5] = t.ob3;
  return RLC_OK;
}

extern "C" int rlc_invalidate_pack(rlc_handle* h, const float* theta) {
  RLC_REQUIRE(h);
  for (int i = 0; i < RLC_MAX_PACKS; ++i)
    if (h->packs[i].theta == theta) h->packs[i].valid = false;
  return RLC_OK;
}

extern "C" int rlc_critic_eval(rlc_handle* h, const rlc_critic* c, const float* s, int B,
                               const float* a, int N, int act_mode, int precision, float* q_out,
                               void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && a && q_out && B >= 0 && N >= 0);
  RLC_REQUIRE(act_mode == RLC_ACT_SHARED || act_mode == RLC_ACT_PER_STATE);
  RLC_REQUIRE(precision >= RLC_PREC_FP32 && precision <= RLC_PREC_FP16C8);
  if ((long long)B * N == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RLC_PREC_AUTO) {
    // AUTO never trades parity for speed: the split tensor mode (fp32-class, 1e-5 of the exact Q) where it applies
    // (large shared-grid T-in evaluations), the fp32 CUDA-core path everywhere else.  The single-rounding fp16/bf16
    // modes (5e-3 max error) are only used when asked for by name.
    precision = (c->topology == RLC_TIN && act_mode == RLC_ACT_SHARED && (long long)B * N >= 16384 &&
                 rlc_umma3_supported(h, c, RLC_PREC_FP16X3))
                    ? RLC_PREC_FP16X3
                    : RLC_PREC_FP32;
  }
  if (precision == RLC_PREC_FP32) return rlc_eval_fp32(h, c, s, B, a, N, act_mode, q_out, st);
  // tensor-core path: T-in only (T-mid is not a dense contraction once hoisted)
  if (c->topology != RLC_TIN) return RLC_ERR_UNSUPPORTED;
  if (h->sm_major != 10) return RLC_ERR_ARCH;
  if ((precision == RLC_PREC_FP16X3 || precision == RLC_PREC_FP16C8) ? !rlc_umma3_supported(h, c, precision)
                                                                     : !rlc_umma_supported(h, c, B, N))
    return RLC_ERR_UNSUPPORTED;
  return rlc_eval_umma(h, c, s, B, a, N, act_mode, precision, q_out, st);
}

// Sampled-action step in one call: Q on the B x N grid and its per-state policy reduction.  fuse != 0: on the split tensor
// kernels the reduction runs inside K1's epilogue (state-major tiles, online softmax) and q[B,N] is only written when q_out
// is given; fuse == 0 (the faster choice at cfg4): evaluation kernel + policy-fused reduction kernel.
extern "C" int rlc_critic_eval_reduce_policy(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* grid,
                                             int N, const float* w, float action_scale, const float* mean,
                                             const float* log_std, const float* v, float entropy_scale, int mode,
                                             int hard, int B_total, int precision, int fuse, float* q_out,
                                             float* loss_b_out, float* dmean_out, float* dlog_std_out, void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && grid && w && mean && log_std && loss_b_out && B >= 0 && N >= 1);
  RLC_REQUIRE((mode == 0 || mode == 1) && (mode == 0 || v) && B_total >= B && B_total >= 1 && action_scale > 0.f);
  RLC_REQUIRE(precision >= RLC_PREC_FP32 && precision <= RLC_PREC_FP16C8 && (mode == 1 || entropy_scale > 0.f));
  if (B == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float alpha = (mode == 1 && hard) ? 0.f : entropy_scale;
  int prec = precision;
  if (prec == RLC_PREC_AUTO)
    prec = (c->topology == RLC_TIN && (long long)B * N >= 16384 && rlc_umma3_supported(h, c, RLC_PREC_FP16X3)) ? RLC_PREC_FP16X3
                                                                                                                : RLC_PREC_FP32;
  if (fuse && (prec == RLC_PREC_FP16X3 || prec == RLC_PREC_FP16C8)) {
    rlc_fuse_args f;
    f.mode = mode + 1; f.A = c->A; f.w = w; f.grid = grid; f.mean = mean; f.log_std = log_std; f.v = v;
    f.action_scale = action_scale; f.alpha = alpha; f.B_total = B_total;
    f.loss_b = loss_b_out; f.dmean = dmean_out; f.dlog_std = dlog_std_out;
    const int rc = rlc_eval_umma_grid3_fused(h, c, s, B, grid, N, prec, q_out, &f, st);
    if (rc != RLC_ERR_UNSUPPORTED) return rc;
  }
  // composition: evaluation into q (scratch when the caller does not want it), then the policy-fused reduction kernel
  float* q = q_out;
  if (!q) {
    // the evaluation kernels use the front of the handle's workspace themselves: a separate block, retired with the handle
    static thread_local float* q_scratch = nullptr;
    static thread_local size_t q_scratch_n = 0;
    const size_t need = (size_t)B * N;
    if (need > q_scratch_n) {
      if (q_scratch) { const int rcr = rlc_retire_block(h, q_scratch); if (rcr) return rcr; }
      RLC_CUDA(cudaMalloc(&q_scratch, need * sizeof(float)));
      q_scratch_n = need;
    }
    q = q_scratch;
  }
  int rc = rlc_critic_eval(h, c, s, B, grid, N, RLC_ACT_SHARED, prec, q, stream);
  if (rc) return rc;
  if (mode == 0)
    return rlc_reduce_fkl_policy(h, q, w, grid, c->A, action_scale, mean, log_std, B, N, entropy_scale, B_total, loss_b_out,
                                 dmean_out, dlog_std_out, nullptr, stream);
  return rlc_reduce_rkl_policy(h, q, v, w, grid, c->A, action_scale, mean, log_std, B, N, entropy_scale, hard, B_total,
                               loss_b_out, dmean_out, dlog_std_out, nullptr, stream);
}
