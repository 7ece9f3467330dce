// Small-minibatch fast path of the ForwardKL / ReverseKL update (cfg1 / cfg5: B = 32 rows, 200-200 networks).
//
// At this size one update_network (forwardkl_network.py:123-209, reversekl_network.py:130-217) is ~3 MFLOP per
// network and the generic path's ~60 dependent 3-5 us kernels are pure latency -- and, with 8 sweep runs sharing a
// GPU, they saturate the kernel front end (measured: 1.6 us per kernel node whatever the node does).  Here
//   rlc_sb_forward  = every B-row forward pass of the update in ONE launch (V(s), V_targ(s'), policy head(s) with
//                     PolicyNetwork.evaluate fused as the epilogue, Q(s,a)); CTAs split the rows, 8 per CTA;
//   rlc_sb_update   = every backward pass + Adam (+ Polyak) in ONE launch: the regression targets / head gradients
//                     are formed in the prologue of each CTA (B values), the gradient of a parameter is consumed by
//                     its Adam step in registers and never written to HBM.  A CTA owns 16 hidden-1 units of one
//                     network (rows of W2: dW2[I,:], dh1[:,I], dW1[:,I], db1[I]) plus a 1/C share of W3/b2; nothing
//                     crosses CTAs, so there is no grid synchronisation: every CTA reads only pre-update values (its
//                     own W2 rows staged in shared memory, W3 from a snapshot taken by the forward launch).
// Latency-bound by construction (a few CTAs, ~1 MB of traffic): no roofline claim; the measure is us per update.
#include "common.cuh"
#include "policy_math.cuh"

#define SB_ROWS 8        // minibatch rows per forward CTA
#define SB_THREADS 256
#define SB_NI 16         // hidden-1 units per backward CTA

struct SbFwdArgs {
  rlc_sb_net net[RLC_SB_MAX_NETS];
  int n_nets, B;
};
struct SbUpdArgs {
  rlc_sb_train net[RLC_SB_MAX_NETS];
  int cta_base[RLC_SB_MAX_NETS + 1];
  int n_nets, B;
  float inv_btotal;
};

__device__ __forceinline__ float sb_x(const rlc_sb_net& n, int b, int k) {
  return k < n.n0 ? n.x0[(long long)b * n.n0 + k] : n.x1[(long long)b * n.n1 + (k - n.n0)];
}
__device__ __forceinline__ float sb_xt(const rlc_sb_train& n, int b, int k) {
  return k < n.n0 ? n.x0[(long long)b * n.n0 + k] : n.x1[(long long)b * n.n1 + (k - n.n0)];
}

// one layer for SB_ROWS rows: out[r][j] = bias[j] + sum_k in[k][r] * W[k][j]; `in` is k-major in shared memory so a
// thread reads its 8 row operands as two broadcast float4.
__device__ __forceinline__ void sb_layer(const float* __restrict__ W, const float* __restrict__ bias, int K, int J,
                                         const float* __restrict__ in_s, int j, float acc[SB_ROWS]) {
  const float bj = bias[j];
#pragma unroll
  for (int r = 0; r < SB_ROWS; ++r) acc[r] = bj;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(W + (long long)k * J + j);
    const float4 a = *reinterpret_cast<const float4*>(in_s + k * SB_ROWS);
    const float4 c = *reinterpret_cast<const float4*>(in_s + k * SB_ROWS + 4);
    acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]);
    acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
    acc[4] = fmaf(c.x, w, acc[4]); acc[5] = fmaf(c.y, w, acc[5]);
    acc[6] = fmaf(c.z, w, acc[6]); acc[7] = fmaf(c.w, w, acc[7]);
  }
}

__global__ void __launch_bounds__(SB_THREADS) k_sb_forward(const __grid_constant__ SbFwdArgs args) {
  extern __shared__ __align__(16) float sm[];
  const int row_ctas = (args.B + SB_ROWS - 1) / SB_ROWS;
  const int ni = blockIdx.x / row_ctas, rc = blockIdx.x % row_ctas;
  const rlc_sb_net& n = args.net[ni];
  const int tid = threadIdx.x, b0 = rc * SB_ROWS;
  const int inp = n.inp, H1 = n.H1, H2 = n.H2, O = n.O;
  float* x_s = sm;                          // [inp][8]
  float* h1_s = x_s + inp * SB_ROWS;        // [H1][8]
  float* h2_s = h1_s + H1 * SB_ROWS;        // [H2][8]
  float* o_s = h2_s + H2 * SB_ROWS;         // [8][O]
  const float* W1 = n.theta;
  const float* b1 = W1 + (long long)inp * H1;
  const float* W2 = b1 + H1;
  const float* b2 = W2 + (long long)H1 * H2;
  const float* W3 = b2 + H2;
  const float* b3 = W3 + (long long)H2 * O;
  if (rc == 0) {
    if (n.w3_snapshot)
      for (int i = tid; i < H2 * O; i += SB_THREADS) n.w3_snapshot[i] = W3[i];
    if (n.adam_state && tid == 0) {         // as k_adam_prep (critic_fp32.cu): the update launch reads the factors
      const int t = ++n.adam_state[0];
      const double bc1 = 1.0 - pow((double)n.beta1, (double)t), bc2 = 1.0 - pow((double)n.beta2, (double)t);
      float lr_eff, isb2;
      if (n.adam_variant == RLC_ADAM_TORCH) {
        lr_eff = (float)((double)n.lr / bc1);
        isb2 = (float)(1.0 / sqrt(bc2));
      } else {
        lr_eff = (float)((double)n.lr * sqrt(bc2) / bc1);
        isb2 = 1.f;
      }
      reinterpret_cast<float*>(n.adam_state)[1] = lr_eff;
      reinterpret_cast<float*>(n.adam_state)[2] = isb2;
    }
  }
  for (int i = tid; i < inp * SB_ROWS; i += SB_THREADS) {
    const int k = i / SB_ROWS, r = i % SB_ROWS;
    x_s[i] = (b0 + r < args.B) ? sb_x(n, b0 + r, k) : 0.f;
  }
  __syncthreads();
  float acc[SB_ROWS];
  for (int j = tid; j < H1; j += SB_THREADS) {
    sb_layer(W1, b1, inp, H1, x_s, j, acc);
#pragma unroll
    for (int r = 0; r < SB_ROWS; ++r) {
      const float a = fmaxf(acc[r], 0.f);
      h1_s[j * SB_ROWS + r] = a;
      if (n.h1 && b0 + r < args.B) n.h1[(long long)(b0 + r) * H1 + j] = a;
    }
  }
  __syncthreads();
  for (int j = tid; j < H2; j += SB_THREADS) {
    sb_layer(W2, b2, H1, H2, h1_s, j, acc);
#pragma unroll
    for (int r = 0; r < SB_ROWS; ++r) {
      const float a = fmaxf(acc[r], 0.f);
      h2_s[j * SB_ROWS + r] = a;
      if (n.h2 && b0 + r < args.B) n.h2[(long long)(b0 + r) * H2 + j] = a;
    }
  }
  __syncthreads();
  // output layer: O is tiny (1 or 2A); one warp per (row, output) dot product
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = warp; p < SB_ROWS * O; p += SB_THREADS / 32) {
    const int r = p / O, o = p % O;
    float s = 0.f;
    for (int j = lane; j < H2; j += 32) s = fmaf(h2_s[j * SB_ROWS + r], __ldg(W3 + (long long)j * O + o), s);
    s = warp_sum(s);
    if (lane == 0) {
      s += b3[o];
      o_s[r * O + o] = s;
      if (b0 + r < args.B) n.out[(long long)(b0 + r) * O + o] = s;
    }
  }
  if (n.policy) {
    __syncthreads();
    if (tid < SB_ROWS && b0 + tid < args.B) {
      const int b = b0 + tid, A = O / 2;
      policy_evaluate_row(o_s + tid * O, n.eps ? n.eps + (long long)b * A : nullptr, A, n.action_scale,
                          n.log_std_min, n.log_std_max, n.action ? n.action + (long long)b * A : nullptr,
                          n.logp ? n.logp + b : nullptr, n.mean ? n.mean + (long long)b * A : nullptr,
                          n.mu_raw ? n.mu_raw + (long long)b * A : nullptr,
                          n.log_std ? n.log_std + (long long)b * A : nullptr, n.z ? n.z + (long long)b * A : nullptr);
    }
  }
}

__device__ __forceinline__ void sb_adam(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                                        float* __restrict__ target, long long i, float g, float lr_eff, float isb2,
                                        float b1, float b2, float eps, float tau) {
  const float mi = b1 * m[i] + (1.f - b1) * g;
  const float vi = b2 * v[i] + (1.f - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  const float pn = theta[i] - lr_eff * mi / (sqrtf(vi) * isb2 + eps);
  theta[i] = pn;
  if (target) target[i] += tau * (pn - target[i]);
}

__global__ void __launch_bounds__(SB_THREADS) k_sb_update(const __grid_constant__ SbUpdArgs args) {
  extern __shared__ __align__(16) float sm[];
  int ni = 0;
  while (ni + 1 < args.n_nets && (int)blockIdx.x >= args.cta_base[ni + 1]) ++ni;
  const rlc_sb_train& n = args.net[ni];
  const int c = blockIdx.x - args.cta_base[ni], C = args.cta_base[ni + 1] - args.cta_base[ni];
  const int tid = threadIdx.x, B = args.B;
  const int inp = n.inp, H1 = n.H1, H2 = n.H2, O = n.O;
  const int H2p = H2 + 1;                       // padded rows: (b, j) and (i, j) walks hit distinct banks
  float* h1_s = sm;                             // [B][SB_NI]     (16-byte aligned rows: read as float4)
  float* dz1_s = h1_s + B * SB_NI;              // [B][SB_NI]
  float* red_s = dz1_s + B * SB_NI;             // [SB_THREADS / 32]
  float* dout_s = red_s + SB_THREADS / 32;      // [B][O]
  float* dz2_s = dout_s + B * O;                // [B][H2p]
  float* w2_s = dz2_s + B * H2p;                // [SB_NI][H2p]   pre-update rows of W2 owned by this CTA
  const long long oW1 = 0, ob1 = (long long)inp * H1, oW2 = ob1 + H1, ob2 = oW2 + (long long)H1 * H2, oW3 = ob2 + H2,
                  ob3 = oW3 + (long long)H2 * O;
  const float lr_eff = reinterpret_cast<const float*>(n.adam_state)[1];
  const float isb2 = reinterpret_cast<const float*>(n.adam_state)[2];
  const int i0 = c * SB_NI, nI = min(SB_NI, H1 - i0);
  // ---- prologue: dLoss/dout for this network's role (B x O values), loss on CTA 0
  float loss_part = 0.f;
  for (int p = tid; p < B * O; p += SB_THREADS) {
    const int b = p / O, o = p % O;
    float d;
    if (n.role == RLC_SB_ROLE_V) {             // value regression (forwardkl_network.py:141-150)
      const float gv = n.gamma[b] * n.v_next[b];
      const float tv = n.sac ? (n.q_new[b] - n.entropy_scale * n.logp[b]) : ((n.r[b] - n.entropy_scale * n.logp[b]) + gv);
      const float e = n.out[b] - tv;
      d = 2.f * args.inv_btotal * e;
      loss_part += e * e * args.inv_btotal;
    } else if (n.role == RLC_SB_ROLE_Q) {      // Q regression on y = r + gamma V_targ(s') (:133-140)
      const float e = n.out[b] - (n.r[b] + n.gamma[b] * n.v_next[b]);
      d = 2.f * args.inv_btotal * e;
      loss_part += e * e * args.inv_btotal;
    } else if (n.role == RLC_SB_ROLE_PI) {     // (dmean, dlog_std) through the log_std clamp (torch.clamp backward)
      const int A = O / 2;
      if (o < A) {
        d = n.dmean[b * A + o];
        if (o == 0) loss_part += n.loss_b[b] / (float)B;
      } else {
        const float raw = n.out[b * O + o];
        d = (raw >= n.log_std_min && raw <= n.log_std_max) ? n.dlog_std[b * A + (o - A)] : 0.f;
      }
    } else {
      d = n.dout[p];
    }
    dout_s[p] = d;
  }
  if (c == 0 && n.loss_out) {                  // deterministic: warp sums, then a serial sum over the 8 warps
    loss_part = warp_sum(loss_part);
    if ((tid & 31) == 0) red_s[tid >> 5] = loss_part;
  }
  // stage this CTA's rows of W2 and its hidden-1 activations
  for (int p = tid; p < nI * H2; p += SB_THREADS) {
    const int i = p / H2, j = p % H2;
    w2_s[i * H2p + j] = n.theta[oW2 + (long long)(i0 + i) * H2 + j];
  }
  for (int p = tid; p < B * SB_NI; p += SB_THREADS) {
    const int b = p / SB_NI, i = p % SB_NI;
    h1_s[p] = i < nI ? n.h1[(long long)b * H1 + i0 + i] : 0.f;
  }
  __syncthreads();
  if (c == 0 && n.loss_out && tid == 0) {
    float s = 0.f;
    for (int w = 0; w < SB_THREADS / 32; ++w) s += red_s[w];
    n.loss_out[0] = s;
  }
  // ---- dz2[b][j] = (dout[b,:] . W3[j,:]) * relu'(h2[b,j])
  for (int p = tid; p < B * H2; p += SB_THREADS) {
    const int b = p / H2, j = p % H2;
    float s = 0.f;
    for (int o = 0; o < O; ++o) s = fmaf(dout_s[b * O + o], n.w3_snapshot[j * O + o], s);
    dz2_s[b * H2p + j] = n.h2[(long long)b * H2 + j] > 0.f ? s : 0.f;
  }
  __syncthreads();
  // ---- dh1[b][i] = dz2[b,:] . W2[i,:] ; dz1 = dh1 * relu'(h1)      (B x nI dot products of length H2)
  for (int p = tid; p < B * SB_NI; p += SB_THREADS) {
    const int b = p / SB_NI, i = p % SB_NI;
    float s = 0.f;
    if (i < nI) {
      const float* dz = dz2_s + b * H2p;
      const float* w = w2_s + i * H2p;
      for (int j = 0; j < H2; ++j) s = fmaf(dz[j], w[j], s);
      if (!(h1_s[p] > 0.f)) s = 0.f;
    }
    dz1_s[p] = s;
  }
  // ---- dW2[i][j] = sum_b h1[b,i] dz2[b,j], straight into Adam (thread = column j, 16 accumulators)
  for (int j = tid; j < H2; j += SB_THREADS) {
    float acc[SB_NI];
#pragma unroll
    for (int i = 0; i < SB_NI; ++i) acc[i] = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = dz2_s[b * H2p + j];
      const float4* h4 = reinterpret_cast<const float4*>(h1_s + b * SB_NI);
#pragma unroll
      for (int q = 0; q < SB_NI / 4; ++q) {
        const float4 h = h4[q];
        acc[4 * q + 0] = fmaf(h.x, d, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(h.y, d, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(h.z, d, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(h.w, d, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < SB_NI; ++i)
      if (i < nI)
        sb_adam(n.theta, n.m, n.v, n.target, oW2 + (long long)(i0 + i) * H2 + j, acc[i], lr_eff, isb2, n.beta1,
                n.beta2, n.eps, n.tau);
  }
  __syncthreads();
  // ---- dW1[k][i] = sum_b x[b,k] dz1[b,i] ; db1[i] = sum_b dz1[b,i]   ((inp + 1) x nI outputs)
  for (int p = tid; p < (inp + 1) * SB_NI; p += SB_THREADS) {
    const int k = p / SB_NI, i = p % SB_NI;
    if (i >= nI) continue;
    float s = 0.f;
    if (k < inp) {
      for (int b = 0; b < B; ++b) s = fmaf(sb_xt(n, b, k), dz1_s[b * SB_NI + i], s);
      sb_adam(n.theta, n.m, n.v, n.target, oW1 + (long long)k * H1 + i0 + i, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
    } else {
      for (int b = 0; b < B; ++b) s += dz1_s[b * SB_NI + i];
      sb_adam(n.theta, n.m, n.v, n.target, ob1 + i0 + i, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
    }
  }
  // ---- this CTA's share of the output layer: dW3[j][o] = sum_b h2[b,j] dout[b,o] ; db2[j] = sum_b dz2[b,j]
  const int Jc = (H2 + C - 1) / C, j0 = c * Jc, nJ = max(0, min(Jc, H2 - j0));
  for (int p = tid; p < nJ * (O + 1); p += SB_THREADS) {
    const int j = j0 + p / (O + 1), o = p % (O + 1);
    float s = 0.f;
    if (o < O) {
      for (int b = 0; b < B; ++b) s = fmaf(n.h2[(long long)b * H2 + j], dout_s[b * O + o], s);
      sb_adam(n.theta, n.m, n.v, n.target, oW3 + (long long)j * O + o, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
    } else {
      for (int b = 0; b < B; ++b) s += dz2_s[b * H2p + j];
      sb_adam(n.theta, n.m, n.v, n.target, ob2 + j, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
    }
  }
  if (c == C - 1 && tid < O) {                 // db3[o] = sum_b dout[b,o]
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dout_s[b * O + tid];
    sb_adam(n.theta, n.m, n.v, n.target, ob3 + tid, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
  }
}

static bool sb_dims_ok(int inp, int H1, int H2, int O, int n0, int n1) {
  return inp >= 1 && inp <= 256 && H1 >= 1 && H1 <= 512 && H2 >= 1 && H2 <= 512 && O >= 1 && O <= 32 &&
         n0 >= 0 && n1 >= 0 && n0 + n1 == inp;
}

extern "C" int rlc_sb_forward(rlc_handle* h, const rlc_sb_net* nets, int n_nets, int B, void* stream) {
  RLC_REQUIRE(h && nets && n_nets >= 1 && n_nets <= RLC_SB_MAX_NETS && B >= 1 && B <= RLC_SB_MAX_B);
  SbFwdArgs args;
  size_t smem = 0;
  for (int i = 0; i < n_nets; ++i) {
    const rlc_sb_net& n = nets[i];
    RLC_REQUIRE(n.theta && n.out && sb_dims_ok(n.inp, n.H1, n.H2, n.O, n.n0, n.n1));
    RLC_REQUIRE((n.n0 == 0 || n.x0) && (n.n1 == 0 || n.x1));
    RLC_REQUIRE(!n.policy || (n.O % 2 == 0 && n.O / 2 <= 16 && n.log_std_min <= n.log_std_max));
    RLC_REQUIRE(!n.adam_state || n.adam_variant == RLC_ADAM_TORCH || n.adam_variant == RLC_ADAM_TF);
    args.net[i] = n;
    const size_t s = sizeof(float) * ((size_t)(n.inp + n.H1 + n.H2) * SB_ROWS + (size_t)SB_ROWS * n.O);
    smem = s > smem ? s : smem;
  }
  args.n_nets = n_nets;
  args.B = B;
  if (smem > 48 * 1024) {
    RLC_REQUIRE(smem <= h->smem_optin);
    RLC_CUDA(cudaFuncSetAttribute(k_sb_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int row_ctas = (B + SB_ROWS - 1) / SB_ROWS;
  k_sb_forward<<<n_nets * row_ctas, SB_THREADS, smem, (cudaStream_t)stream>>>(args);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_sb_update(rlc_handle* h, const rlc_sb_train* nets, int n_nets, int B, int B_total, void* stream) {
  RLC_REQUIRE(h && nets && n_nets >= 1 && n_nets <= RLC_SB_MAX_NETS && B >= 1 && B <= RLC_SB_MAX_B && B_total >= B);
  static int smem_set = 0;
  SbUpdArgs args;
  size_t smem = 0;
  int base = 0;
  for (int i = 0; i < n_nets; ++i) {
    const rlc_sb_train& n = nets[i];
    RLC_REQUIRE(n.theta && n.m && n.v && n.adam_state && sb_dims_ok(n.inp, n.H1, n.H2, n.O, n.n0, n.n1));
    RLC_REQUIRE((n.n0 == 0 || n.x0) && (n.n1 == 0 || n.x1) && n.h1 && n.h2 && n.out && n.w3_snapshot);
    switch (n.role) {
      case RLC_SB_ROLE_DOUT: RLC_REQUIRE(n.dout); break;
      case RLC_SB_ROLE_V:
        RLC_REQUIRE(n.O == 1 && n.r && n.gamma && n.v_next && n.logp && (!n.sac || n.q_new));
        break;
      case RLC_SB_ROLE_Q: RLC_REQUIRE(n.O == 1 && n.r && n.gamma && n.v_next); break;
      case RLC_SB_ROLE_PI:
        RLC_REQUIRE(n.O % 2 == 0 && n.dmean && n.dlog_std && (!n.loss_out || n.loss_b));
        break;
      default: return RLC_ERR_INVALID;
    }
    args.net[i] = n;
    args.cta_base[i] = base;
    base += (n.H1 + SB_NI - 1) / SB_NI;
    const size_t s = sizeof(float) * ((size_t)B * n.O + (size_t)(B + SB_NI) * (n.H2 + 1) + 2 * (size_t)B * SB_NI +
                                      SB_THREADS / 32);
    smem = s > smem ? s : smem;
    rlc_invalidate_pack(h, n.theta);
    if (n.target) rlc_invalidate_pack(h, n.target);
  }
  args.cta_base[n_nets] = base;
  args.n_nets = n_nets;
  args.B = B;
  args.inv_btotal = 1.f / (float)B_total;
  if (smem > 48 * 1024 && (int)smem > smem_set) {
    RLC_REQUIRE(smem <= h->smem_optin);
    RLC_CUDA(cudaFuncSetAttribute(k_sb_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  k_sb_update<<<base, SB_THREADS, smem, (cudaStream_t)stream>>>(args);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
