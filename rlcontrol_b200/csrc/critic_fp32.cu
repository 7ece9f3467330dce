// CUDA-core fp32 critic kernels: fused B x N evaluation for both topologies, the hoisted T-mid
// row kernel (with dQ/da), and the B-row training path (forward, backward, weight grads) built on
// one tiled SGEMM.  These are the exact-parity path and the path for small, launch-bound
// configs; the large T-in evaluation runs on tcgen05 (critic_umma.cu).
#include "common.cuh"
#include "rows_gemm.cuh"

// =============================================================================================
// Fused two-layer kernel: TM = 32 (or 8, for few rows) rows per CTA, layer-1
// activations kept in shared memory, W2 streamed through a double-buffered shared-memory ring in
// 16-row chunks (register-staged: the loads of chunk c+1 are in flight while chunk c is consumed),
// layer 2 register-tiled: a warp owns TM/8 rows (activation reads are warp-wide broadcasts, float4
// along k) and a lane owns the columns lane, lane+32, ... (conflict-free W2 reads).
// Epilogue = w3 dot (T-in q) or store p (T-mid state term).
// =============================================================================================
enum { MODE_TIN_Q = 0, MODE_TMID_P = 1 };
#define MLP2_KC 16

template <int NJ, int MODE, int TM>
__global__ void __launch_bounds__(256)
k_mlp2_rows(const float* __restrict__ s, const float* __restrict__ a, int act_per_state,
            long long R, int N, int S, int A, int H1, int H2, const float* __restrict__ W1,
            const float* __restrict__ b1, const float* __restrict__ W2,
            const float* __restrict__ b2, const float* __restrict__ w3,
            const float* __restrict__ b3, const float* __restrict__ smin,
            const float* __restrict__ smax, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  constexpr int H2S = NJ * 32;
  const int K1 = (MODE == MODE_TIN_Q) ? S + A : S;
  const int K1P = (K1 + 3) & ~3;
  const int nchunks = (H1 + MLP2_KC - 1) / MLP2_KC;
  const int H1P = nchunks * MLP2_KC;
  float* xs = sm;                        // [TM][K1P]
  float* h1s = xs + TM * K1P;       // [TM][H1P]   (columns >= H1 are zero)
  float* w2s = h1s + TM * H1P;      // [2][KC][H2S]
  const long long r0 = (long long)blockIdx.x * TM;
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  constexpr int RW = TM / 8;             // rows per warp

  // first W2 chunk: issue the loads before the layer-1 work so their latency is hidden behind it
  constexpr int NST = NJ * MLP2_KC / 8;  // chunk elements per thread
  float stg[NST];
#pragma unroll
  for (int t = 0; t < NST; ++t) {
    const int e = tid + 256 * t, kr = e / H2S, col = e - kr * H2S;
    stg[t] = (kr < H1 && col < H2) ? __ldg(W2 + (long long)kr * H2 + col) : 0.f;
  }

  for (int i = tid; i < TM * K1; i += 256) {
    const int r = i / K1, k = i - r * K1;
    const long long row = r0 + r;
    float v = 0.f;
    if (row < R) {
      if (MODE == MODE_TIN_Q) {
        const long long b = row / N;
        if (k < S) {
          v = s[b * S + k];
          if (smin) v = fminf(fmaxf(v, smin[k]), smax[k]);
        } else {
          const long long arow = act_per_state ? row : (row - b * N);
          v = a[arow * A + (k - S)];
        }
      } else {
        v = s[row * S + k];
        if (smin) v = fminf(fmaxf(v, smin[k]), smax[k]);
      }
    }
    xs[r * K1P + k] = v;
  }
  __syncthreads();

  for (int i = tid; i < TM * H1P; i += 256) {
    const int r = i / H1P, j = i - r * H1P;
    float v = 0.f;
    if (j < H1) {
      float acc = b1[j];
      const float* xr = xs + r * K1P;
      for (int k = 0; k < K1; ++k) acc = fmaf(xr[k], __ldg(W1 + (long long)k * H1 + j), acc);
      v = fmaxf(acc, 0.f);
    }
    h1s[i] = v;
  }
#pragma unroll
  for (int t = 0; t < NST; ++t) w2s[tid + 256 * t] = stg[t];
  __syncthreads();

  float acc[RW][NJ];
#pragma unroll
  for (int i = 0; i < RW; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  const float* hrow = h1s + (rg * RW) * H1P;
  for (int c = 0; c < nchunks; ++c) {
    const bool more = c + 1 < nchunks;
    if (more) {
      const int k0 = (c + 1) * MLP2_KC;
#pragma unroll
      for (int t = 0; t < NST; ++t) {
        const int e = tid + 256 * t, kr = k0 + e / H2S, col = e % H2S;
        stg[t] = (kr < H1 && col < H2) ? __ldg(W2 + (long long)kr * H2 + col) : 0.f;
      }
    }
    const float* wb = w2s + (c & 1) * (MLP2_KC * H2S) + lane;
#pragma unroll
    for (int kk = 0; kk < MLP2_KC; kk += 4) {
      float4 hv[RW];
#pragma unroll
      for (int i = 0; i < RW; ++i)
        hv[i] = *reinterpret_cast<const float4*>(hrow + i * H1P + c * MLP2_KC + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float w = wb[(kk + k4) * H2S + 32 * j];
#pragma unroll
          for (int i = 0; i < RW; ++i) {
            const float hval = k4 == 0 ? hv[i].x : k4 == 1 ? hv[i].y : k4 == 2 ? hv[i].z : hv[i].w;
            acc[i][j] = fmaf(hval, w, acc[i][j]);
          }
        }
      }
    }
    if (more) {
      float* wn = w2s + ((c + 1) & 1) * (MLP2_KC * H2S);
#pragma unroll
      for (int t = 0; t < NST; ++t) wn[tid + 256 * t] = stg[t];
    }
    __syncthreads();
  }

  if (MODE == MODE_TIN_Q) {
    float qs[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) qs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = lane + 32 * j;
      if (col < H2) {
        const float bb = __ldg(b2 + col), ww = __ldg(w3 + col);
#pragma unroll
        for (int i = 0; i < RW; ++i) qs[i] = fmaf(ww, fmaxf(acc[i][j] + bb, 0.f), qs[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < RW; ++i) qs[i] = warp_sum(qs[i]);
    if (lane == 0) {
      const float bb3 = __ldg(b3);
#pragma unroll
      for (int i = 0; i < RW; ++i) {
        const long long row = r0 + rg * RW + i;
        if (row < R) out[row] = qs[i] + bb3;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = lane + 32 * j;
      if (col < H2) {
        const float bb = __ldg(b2 + col);
#pragma unroll
        for (int i = 0; i < RW; ++i) {
          const long long row = r0 + rg * RW + i;
          if (row < R) out[row * H2 + col] = acc[i][j] + bb;
        }
      }
    }
  }
}

template <int MODE>
static int launch_mlp2(rlc_handle* h, const float* s, const float* a, int act_per_state,
                       long long R, int N, int S, int A, int H1, int H2, const float* W1,
                       const float* b1, const float* W2, const float* b2, const float* w3,
                       const float* b3, const float* smin, const float* smax, float* out,
                       cudaStream_t st) {
  if (R == 0) return RLC_OK;
  const int K1 = (MODE == MODE_TIN_Q) ? S + A : S;
  const int K1P = (K1 + 3) & ~3;
  const int H1P = (H1 + MLP2_KC - 1) / MLP2_KC * MLP2_KC;
  const int NJ = H2 <= 64 ? 2 : H2 <= 128 ? 4 : H2 <= 224 ? 7 : H2 <= 320 ? 10 : 16;
  if (H2 > 512) return RLC_ERR_UNSUPPORTED;
  // few rows (a CEM call's B = 256 state terms, a small grid evaluation): 8 rows per CTA (one per warp) instead of 32, so
  // the launch covers 4x as many SMs and a CTA's serial layer-1 + W2-streaming chain is 4x shorter; same k order, so the
  // results are bit-identical
  const int TMv = R <= 16LL * h->num_sms ? 8 : 32;
  const size_t smem =
      (size_t)(TMv * K1P + TMv * H1P + 2 * MLP2_KC * NJ * 32) * sizeof(float);
  if (smem > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  const long long blocks = (R + TMv - 1) / TMv;
  if (blocks > 0x7fffffffLL) return RLC_ERR_INVALID;
#define RLC_MLP2_LAUNCH(NJV, TMV)                                                                \
  {                                                                                              \
    auto kern = k_mlp2_rows<NJV, MODE, TMV>;                                                     \
    RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<(unsigned)blocks, 256, smem, st>>>(s, a, act_per_state, R, N, S, A, H1, H2, W1, b1,   \
                                              W2, b2, w3, b3, smin, smax, out);                  \
  }
#define RLC_MLP2_CASE(NJV)                                                                       \
  {                                                                                              \
    if (TMv == 8) RLC_MLP2_LAUNCH(NJV, 8) else RLC_MLP2_LAUNCH(NJV, 32)                          \
  }
  if (NJ == 2) RLC_MLP2_CASE(2)
  else if (NJ == 4) RLC_MLP2_CASE(4)
  else if (NJ == 7) RLC_MLP2_CASE(7)
  else if (NJ == 10) RLC_MLP2_CASE(10)
  else RLC_MLP2_CASE(16)
#undef RLC_MLP2_CASE
#undef RLC_MLP2_LAUNCH
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// T-mid state term for FEW rows (a CEM call's B = 256, an Actor-Expert minibatch): the launch is a latency chain, not
// a throughput problem -- k_mlp2_rows walks W2 in 25 dependent 16-row chunks per CTA.  Here a CTA owns 8 rows x 32
// output columns, so its whole W2 slice (H1 x 32 floats) fits in shared memory and is fetched by ONE wave of cp.async
// that flies under the layer-1 work; a warp is a row, a lane a column.  grid = ceil(R/8) x ceil(H2/32) CTAs (320 at
// B=256, H2=300 instead of 32).  Same operation order as k_mlp2_rows (k ascending, bias last): bit-identical results.
#define STC_ROWS 8
__global__ void __launch_bounds__(256)
k_state_term_cols(const float* __restrict__ s, int R, int S, int H1, int H2, const float* __restrict__ W1,
                  const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                  const float* __restrict__ smin, const float* __restrict__ smax, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int K1P = (S + 3) & ~3, H1P = (H1 + 3) & ~3;
  float* xs = sm;                          // [8][K1P]
  float* h1s = xs + STC_ROWS * K1P;        // [8][H1P]  (columns >= H1 are zero)
  float* w2s = h1s + STC_ROWS * H1P;       // [H1][32]
  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
  const int r0 = blockIdx.x * STC_ROWS, col0 = blockIdx.y * 32;
  // the CTA's W2 slice, straight into shared memory (4-byte cp.async: the slice rows are not 16-byte aligned in general)
  for (int e = tid; e < H1 * 32; e += 256) {
    const int k = e >> 5, col = col0 + (e & 31);
    if (col < H2) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(w2s + e);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(W2 + (long long)k * H2 + col) : "memory");
    } else {
      w2s[e] = 0.f;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < STC_ROWS * S; i += 256) {
    const int r = i / S, k = i - r * S;
    float v = 0.f;
    if (r0 + r < R) {
      v = s[(long long)(r0 + r) * S + k];
      if (smin) v = fminf(fmaxf(v, smin[k]), smax[k]);
    }
    xs[r * K1P + k] = v;
  }
  __syncthreads();
  for (int i = tid; i < STC_ROWS * H1P; i += 256) {
    const int r = i / H1P, j = i - r * H1P;
    float v = 0.f;
    if (j < H1) {
      float acc = b1[j];
      const float* xr = xs + r * K1P;
      for (int k = 0; k < S; ++k) acc = fmaf(xr[k], __ldg(W1 + (long long)k * H1 + j), acc);
      v = fmaxf(acc, 0.f);
    }
    h1s[i] = v;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float* hrow = h1s + wr * H1P;
  const float* wb = w2s + lane;
  float acc = 0.f;
  for (int k = 0; k < H1P; k += 4) {
    const float4 hv = *reinterpret_cast<const float4*>(hrow + k);
    acc = fmaf(hv.x, wb[(k + 0) * 32], acc);
    if (k + 1 < H1) acc = fmaf(hv.y, wb[(k + 1) * 32], acc);
    if (k + 2 < H1) acc = fmaf(hv.z, wb[(k + 2) * 32], acc);
    if (k + 3 < H1) acc = fmaf(hv.w, wb[(k + 3) * 32], acc);
  }
  const int row = r0 + wr, col = col0 + lane;
  if (row < R && col < H2) out[(long long)row * H2 + col] = acc + __ldg(b2 + col);
}

__global__ void k_build_x(const float* __restrict__ s, const float* __restrict__ a, long long R, int S, int A, int tin,
                          const float* __restrict__ smin, const float* __restrict__ smax, float* __restrict__ X, int rep);

size_t rlc_tmid_state_scratch_floats(const rlc_critic* c, int B) {
  return (((size_t)B * c->S + 3) & ~(size_t)3) + (size_t)B * c->H1 + 64;
}

int rlc_tmid_state_term(rlc_handle* h, const rlc_critic* c, const float* s, int B, float* p_out,
                        cudaStream_t st, float* scratch) {
  const ThetaView t = theta_view(RLC_TMID, c->S, c->A, c->H1, c->H2);
  const float* th = c->theta;
  if (scratch && B > 0 && rlc_gemm_tc_ok(h, B, c->H2, c->H1)) {
    // dense batch of states: two GEMMs on the tensor cores (3xTF32, fp32-class) instead of the fused fp32 row kernel
    float* X = scratch;
    float* Z1 = scratch + (((size_t)B * c->S + 3) & ~(size_t)3);
    const long long n = (long long)B * c->S;
    k_build_x<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, nullptr, B, c->S, c->A, 0, c->smin, c->smax, X, 1);
    RLC_LAUNCH_CHECK(h);
    GemmEpi e1{th + t.ob1, nullptr, 0, 0, 1.f};
    int rc = gemm(h, false, false, B, c->H1, c->S, X, c->S, th + t.oW1, c->H1, Z1, c->H1, e1, st);
    if (rc) return rc;
    GemmEpi e2{th + t.ob2, nullptr, 0, 1, 1.f};   // A := relu(Z1); W2[:H1] are the first H1 rows of W2
    return gemm(h, false, false, B, c->H2, c->H1, Z1, c->H1, th + t.oW2, c->H2, p_out, c->H2, e2, st);
  }
  {
    const int K1P = (c->S + 3) & ~3, H1P = (c->H1 + 3) & ~3;
    const size_t smem = (size_t)(STC_ROWS * K1P + STC_ROWS * H1P + c->H1 * 32) * sizeof(float);
    if (B > 0 && B <= 16 * h->num_sms && smem <= h->smem_optin) {
      RLC_CUDA(cudaFuncSetAttribute(k_state_term_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const dim3 grid((unsigned)((B + STC_ROWS - 1) / STC_ROWS), (unsigned)((c->H2 + 31) / 32));
      k_state_term_cols<<<grid, 256, smem, st>>>(s, B, c->S, c->H1, c->H2, th + t.oW1, th + t.ob1, th + t.oW2,
                                                 th + t.ob2, c->smin, c->smax, p_out);
      RLC_LAUNCH_CHECK(h);
      return RLC_OK;
    }
  }
  return launch_mlp2<MODE_TMID_P>(h, s, nullptr, 0, B, 1, c->S, c->A, c->H1, c->H2, th + t.oW1,
                                  th + t.ob1, th + t.oW2, th + t.ob2, th + t.ow3, th + t.ob3,
                                  c->smin, c->smax, p_out, st);
}

// =============================================================================================
// T-mid row kernel (hoisted): q = sum_j w3_j relu(p[b][j] + sum_i a_i W2a[i][j]) + b3, and
// optionally dq/da_i = sum_j [z_j>0] w3_j W2a[i][j].  One thread per (b,n) row; W2a / w3 in smem.
// =============================================================================================
template <int AT, bool GRAD>
__global__ void __launch_bounds__(256)
k_tmid_rows(const float* __restrict__ p, const float* __restrict__ a, int act_per_state,
            long long R, int N, int A, int H2, const float* __restrict__ W2a,
            const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ q_out,
            float* __restrict__ dqda_out) {
  extern __shared__ float sm[];
  const int H2P = (H2 + 3) & ~3;
  float* w3s = sm;         // [H2P]
  float* was = sm + H2P;   // [AT][H2P]
  const int tid = threadIdx.x;
  for (int i = tid; i < H2P; i += 256) w3s[i] = (i < H2) ? w3[i] : 0.f;
  for (int i = tid; i < AT * H2P; i += 256) {
    const int ai = i / H2P, j = i - ai * H2P;
    was[i] = (ai < A && j < H2) ? W2a[(long long)ai * H2 + j] : 0.f;
  }
  __syncthreads();
  const long long row = (long long)blockIdx.x * 256 + tid;
  if (row >= R) return;
  const long long b = row / N;
  const long long arow = act_per_state ? row : (row - b * N);
  float ar[AT];
#pragma unroll
  for (int i = 0; i < AT; ++i) ar[i] = (i < A) ? a[arow * A + i] : 0.f;
  const float* pb = p + b * H2;
  float q = 0.f;
  float g[AT];
#pragma unroll
  for (int i = 0; i < AT; ++i) g[i] = 0.f;
  for (int j = 0; j < H2P; j += 4) {
    float z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) z[u] = (j + u < H2) ? __ldg(pb + j + u) : 0.f;
#pragma unroll
    for (int i = 0; i < AT; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(was + i * H2P + j);
      z[0] = fmaf(ar[i], w.x, z[0]);
      z[1] = fmaf(ar[i], w.y, z[1]);
      z[2] = fmaf(ar[i], w.z, z[2]);
      z[3] = fmaf(ar[i], w.w, z[3]);
    }
    const float4 w3v = *reinterpret_cast<const float4*>(w3s + j);
    const float w3a[4] = {w3v.x, w3v.y, w3v.z, w3v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      q = fmaf(w3a[u], fmaxf(z[u], 0.f), q);
      if (GRAD) {
        const float gw = (z[u] > 0.f) ? w3a[u] : 0.f;
#pragma unroll
        for (int i = 0; i < AT; ++i) g[i] = fmaf(gw, was[i * H2P + j + u], g[i]);
      }
    }
  }
  if (q_out) q_out[row] = q + __ldg(b3);
  if (GRAD) {
#pragma unroll
    for (int i = 0; i < AT; ++i)
      if (i < A) dqda_out[row * A + i] = g[i];
  }
}

// Forward-only variant with a 4-row register tile per thread (rows base + tid + 256 u): every broadcast read of
// W2a / w3 from shared memory feeds four rows, and the action dimension is an exact template parameter.  With one row per
// thread (k_tmid_rows) the kernel was bound by the load/store unit -- 9 broadcast LDS.128 per 4 columns per row, and
// A = 6 padded to 8 -- at 18 TFLOP/s fp32 on 2(A+1)H2 flop/row.
#define TMID_RPT 4
template <int AT, bool GRAD>
__global__ void __launch_bounds__(256)
k_tmid_rows4(const float* __restrict__ p, const float* __restrict__ a, int act_per_state, long long R, int N,
             int H2, const float* __restrict__ W2a, const float* __restrict__ w3, const float* __restrict__ b3,
             float* __restrict__ q_out, float* __restrict__ dqda_out) {
  extern __shared__ float sm[];
  const int H2P = (H2 + 3) & ~3;
  float* w3s = sm;         // [H2P]
  float* was = sm + H2P;   // [AT][H2P]
  const int tid = threadIdx.x;
  for (int i = tid; i < H2P; i += 256) w3s[i] = (i < H2) ? w3[i] : 0.f;
  for (int i = tid; i < AT * H2P; i += 256) {
    const int ai = i / H2P, j = i - ai * H2P;
    was[i] = (j < H2) ? W2a[(long long)ai * H2 + j] : 0.f;
  }
  __syncthreads();
  const long long row0 = (long long)blockIdx.x * (256 * TMID_RPT) + tid;
  float ar[TMID_RPT][AT], q[TMID_RPT];
  float g[GRAD ? TMID_RPT : 1][AT];
  const float* pb[TMID_RPT];
#pragma unroll
  for (int u = 0; u < TMID_RPT; ++u) {
    const long long row = min(row0 + 256 * u, R - 1);       // clamped: tail threads recompute the last row, never store it
    const long long b = row / N;
    const long long arow = act_per_state ? row : (row - b * N);
#pragma unroll
    for (int i = 0; i < AT; ++i) ar[u][i] = a[arow * AT + i];
    pb[u] = p + b * H2;
    q[u] = 0.f;
    if (GRAD) {
#pragma unroll
      for (int i = 0; i < AT; ++i) g[u][i] = 0.f;
    }
  }
  for (int j = 0; j < H2P; j += 4) {
    float z[TMID_RPT][4];
#pragma unroll
    for (int u = 0; u < TMID_RPT; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) z[u][c] = (j + c < H2) ? __ldg(pb[u] + j + c) : 0.f;
    float4 w[AT];
#pragma unroll
    for (int i = 0; i < AT; ++i) {
      w[i] = *reinterpret_cast<const float4*>(was + i * H2P + j);
#pragma unroll
      for (int u = 0; u < TMID_RPT; ++u) {
        z[u][0] = fmaf(ar[u][i], w[i].x, z[u][0]);
        z[u][1] = fmaf(ar[u][i], w[i].y, z[u][1]);
        z[u][2] = fmaf(ar[u][i], w[i].z, z[u][2]);
        z[u][3] = fmaf(ar[u][i], w[i].w, z[u][3]);
      }
    }
    const float4 w3v = *reinterpret_cast<const float4*>(w3s + j);
#pragma unroll
    for (int u = 0; u < TMID_RPT; ++u) {
      q[u] = fmaf(w3v.x, fmaxf(z[u][0], 0.f), q[u]);
      q[u] = fmaf(w3v.y, fmaxf(z[u][1], 0.f), q[u]);
      q[u] = fmaf(w3v.z, fmaxf(z[u][2], 0.f), q[u]);
      q[u] = fmaf(w3v.w, fmaxf(z[u][3], 0.f), q[u]);
      if (GRAD) {   // dq/da_i += [z_j > 0] w3_j W2a[i][j]: same order over j as the one-row kernel (bit-identical)
        const float g0 = (z[u][0] > 0.f) ? w3v.x : 0.f, g1 = (z[u][1] > 0.f) ? w3v.y : 0.f;
        const float g2 = (z[u][2] > 0.f) ? w3v.z : 0.f, g3 = (z[u][3] > 0.f) ? w3v.w : 0.f;
#pragma unroll
        for (int i = 0; i < AT; ++i) {
          g[u][i] = fmaf(g0, w[i].x, g[u][i]);
          g[u][i] = fmaf(g1, w[i].y, g[u][i]);
          g[u][i] = fmaf(g2, w[i].z, g[u][i]);
          g[u][i] = fmaf(g3, w[i].w, g[u][i]);
        }
      }
    }
  }
  const float bias3 = __ldg(b3);
#pragma unroll
  for (int u = 0; u < TMID_RPT; ++u) {
    if (row0 + 256 * u < R) {
      if (q_out) q_out[row0 + 256 * u] = q[u] + bias3;
      if (GRAD) {
#pragma unroll
        for (int i = 0; i < AT; ++i) dqda_out[(row0 + 256 * u) * AT + i] = g[u][i];
      }
    }
  }
}

template <int AT, bool GRAD>
static int launch_tmid_rows4(rlc_handle* h, const rlc_critic* c, const float* p, const float* a, int act_per_state,
                             long long R, int N, float* q_out, float* dqda_out, cudaStream_t st) {
  const ThetaView t = theta_view(RLC_TMID, c->S, c->A, c->H1, c->H2);
  const int H2P = (c->H2 + 3) & ~3;
  const size_t smem = (size_t)(H2P * (1 + AT)) * sizeof(float);
  if (smem > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  const long long blocks = (R + 256 * TMID_RPT - 1) / (256 * TMID_RPT);
  if (blocks > 0x7fffffffLL) return RLC_ERR_INVALID;
  auto kern = k_tmid_rows4<AT, GRAD>;
  RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, 256, smem, st>>>(p, a, act_per_state, R, N, c->H2, c->theta + t.oW2 + (int64_t)c->H1 * c->H2,
                                            c->theta + t.ow3, c->theta + t.ob3, q_out, dqda_out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

template <bool GRAD>
static int launch_tmid_rows(rlc_handle* h, const rlc_critic* c, const float* p, const float* a,
                            int act_per_state, long long R, int N, float* q_out, float* dqda_out,
                            cudaStream_t st) {
  if (R == 0) return RLC_OK;
  if (!GRAD && q_out && N > 0 && R % N == 0 && rlc_tmid_tc_ok(h, c, R, N))   // large stacks: 128-row tiles on tcgen05
    return rlc_tmid_rows_tc(h, c, p, a, act_per_state, (int)(R / N), N, q_out, st);
  if ((GRAD || q_out) && R >= (long long)h->num_sms * 256 * TMID_RPT && c->A <= 8) {   // enough rows to fill the machine
    switch (c->A) {
      case 1: return launch_tmid_rows4<1, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 2: return launch_tmid_rows4<2, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 3: return launch_tmid_rows4<3, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 4: return launch_tmid_rows4<4, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 5: return launch_tmid_rows4<5, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 6: return launch_tmid_rows4<6, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      case 7: return launch_tmid_rows4<7, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
      default: return launch_tmid_rows4<8, GRAD>(h, c, p, a, act_per_state, R, N, q_out, dqda_out, st);
    }
  }
  const ThetaView t = theta_view(RLC_TMID, c->S, c->A, c->H1, c->H2);
  const float* W2a = c->theta + t.oW2 + (int64_t)c->H1 * c->H2;
  const float* w3 = c->theta + t.ow3;
  const float* b3 = c->theta + t.ob3;
  const int H2P = (c->H2 + 3) & ~3;
  const long long blocks = (R + 255) / 256;
  if (blocks > 0x7fffffffLL) return RLC_ERR_INVALID;
#define RLC_TMID_CASE(AT)                                                                        \
  {                                                                                              \
    const size_t smem = (size_t)(H2P * (1 + AT)) * sizeof(float);                                \
    if (smem > h->smem_optin) return RLC_ERR_UNSUPPORTED;                                        \
    auto kern = k_tmid_rows<AT, GRAD>;                                                           \
    RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<(unsigned)blocks, 256, smem, st>>>(p, a, act_per_state, R, N, c->A, c->H2, W2a, w3,   \
                                              b3, q_out, dqda_out);                              \
  }
  if (c->A <= 1) RLC_TMID_CASE(1)
  else if (c->A <= 2) RLC_TMID_CASE(2)
  else if (c->A <= 4) RLC_TMID_CASE(4)
  else if (c->A <= 8) RLC_TMID_CASE(8)
  else if (c->A <= 16) RLC_TMID_CASE(16)
  else if (c->A <= 32) RLC_TMID_CASE(32)
  else return RLC_ERR_UNSUPPORTED;
#undef RLC_TMID_CASE
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

int rlc_eval_fp32(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                  int act_mode, float* q_out, cudaStream_t st) {
  const long long R = (long long)B * N;
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  const float* th = c->theta;
  if (c->topology == RLC_TIN) {
    return launch_mlp2<MODE_TIN_Q>(h, s, a, act_mode == RLC_ACT_PER_STATE, R, N, c->S, c->A, c->H1,
                                   c->H2, th + t.oW1, th + t.ob1, th + t.oW2, th + t.ob2,
                                   th + t.ow3, th + t.ob3, c->smin, c->smax, q_out, st);
  }
  void* ws = nullptr;
  const size_t np = ((size_t)B * c->H2 + 3) & ~(size_t)3;
  int rc = rlc_workspace(h, (np + rlc_tmid_state_scratch_floats(c, B)) * sizeof(float), &ws);
  if (rc) return rc;
  float* p = (float*)ws;
  rc = rlc_tmid_state_term(h, c, s, B, p, st, p + np);
  if (rc) return rc;
  return launch_tmid_rows<false>(h, c, p, a, act_mode == RLC_ACT_PER_STATE, R, N, q_out, nullptr,
                                 st);
}

// X[R, in1] = T-in: [clip(s), a] ; T-mid: clip(s)
__global__ void k_build_x(const float* __restrict__ s, const float* __restrict__ a, long long R,
                          int S, int A, int tin, const float* __restrict__ smin,
                          const float* __restrict__ smax, float* __restrict__ X, int rep) {
  const int K1 = tin ? S + A : S;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * K1) return;
  const long long r = i / K1;
  const int k = (int)(i - r * K1);
  float v;
  if (k < S) {
    v = s[(r / rep) * S + k];  // rep > 1: every state row serves `rep` consecutive stacked rows
    if (smin) v = fminf(fmaxf(v, smin[k]), smax[k]);
  } else {
    v = a[r * A + (k - S)];
  }
  X[i] = v;
}

// T-mid: ZC[R, H1+A] = [relu(Z1), a]
__global__ void k_build_zc(const float* __restrict__ Z1, const float* __restrict__ a, long long R,
                           int H1, int A, float* __restrict__ ZC) {
  const int W = H1 + A;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * W) return;
  const long long r = i / W;
  const int k = (int)(i - r * W);
  ZC[i] = (k < H1) ? fmaxf(Z1[r * H1 + k], 0.f) : a[r * A + (k - H1)];
}

// Per row: q = relu(Z2) . w3 + b3 ; dq = scale*(q - y) (if y) else 1 ; G2 = dq * w3 * [Z2>0].
// One warp per row. loss_acc += sum (q-y)^2 / B_total.
__global__ void k_head(const float* __restrict__ Z2, long long R, int H2,
                       const float* __restrict__ w3, const float* __restrict__ b3,
                       const float* __restrict__ y, float inv_btotal, float* __restrict__ q_out,
                       float* __restrict__ dq_out, float* __restrict__ G2,
                       float* __restrict__ loss_acc) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* z = Z2 + row * H2;
  float q = 0.f;
  for (int j = lane; j < H2; j += 32) q = fmaf(w3[j], fmaxf(z[j], 0.f), q);
  q = warp_sum(q) + b3[0];
  float dq = 1.f;
  if (y) {
    const float d = q - y[row];
    dq = 2.f * inv_btotal * d;
    if (lane == 0 && loss_acc) atomicAdd(loss_acc, d * d * inv_btotal);
  }
  if (lane == 0) {
    if (q_out) q_out[row] = q;
    if (dq_out) dq_out[row] = dq;
  }
  if (G2)
    for (int j = lane; j < H2; j += 32) G2[row * H2 + j] = (z[j] > 0.f) ? dq * w3[j] : 0.f;
}

struct TrainWs {
  float *X, *Z1, *ZC, *Z2, *G2, *G1, *dq, *slabs, *part;
};

static size_t train_ws_bytes(const rlc_critic* c, long long R) {
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  size_t n = (size_t)R * (t.in1 + 2 * (size_t)c->H1 + 2 * (size_t)c->H2 + 1);
  if (c->topology == RLC_TMID) n += (size_t)R * (c->H1 + c->A);
  const size_t maxw = (size_t)(c->H1 > c->H2 ? c->H1 : c->H2) + 1;
  n += (size_t)SPLITK_MAX * ((size_t)(t.in2 > t.in1 ? t.in2 : t.in1) + 1) * maxw;   // split-K slabs of the largest weight
  n += (size_t)COLRED_MAX_CHUNKS * maxw;                                        // column-reduction partials
  return n * sizeof(float) + 64 * 16;
}

static TrainWs carve(const rlc_critic* c, long long R, float* base) {
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  auto take = [&](size_t n) {
    float* p = base;
    base += (n + 3) & ~(size_t)3;
    return p;
  };
  TrainWs w;
  w.X = take((size_t)R * t.in1);
  w.Z1 = take((size_t)R * c->H1);
  w.ZC = (c->topology == RLC_TMID) ? take((size_t)R * (c->H1 + c->A)) : nullptr;
  w.Z2 = take((size_t)R * c->H2);
  w.G2 = take((size_t)R * c->H2);
  w.G1 = take((size_t)R * c->H1);
  w.dq = take((size_t)R);
  const size_t maxw = (size_t)(c->H1 > c->H2 ? c->H1 : c->H2) + 1;
  w.slabs = take((size_t)SPLITK_MAX * ((size_t)(t.in2 > t.in1 ? t.in2 : t.in1) + 1) * maxw);
  w.part = take((size_t)COLRED_MAX_CHUNKS * maxw);
  return w;
}

// Forward over R stacked rows storing pre-activations; then head.  y==nullptr -> dq=1 (dQ/da).
static int forward_rows(rlc_handle* h, const rlc_critic* c, const float* s, const float* a,
                        long long R, const float* y, float inv_btotal, TrainWs& w, float* q_out,
                        float* loss_acc, cudaStream_t st, int rep = 1) {
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  const float* th = c->theta;
  const int tin = c->topology == RLC_TIN;
  {
    const long long n = R * t.in1;
    k_build_x<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, a, R, c->S, c->A, tin, c->smin,
                                                           c->smax, w.X, rep);
    RLC_LAUNCH_CHECK(h);
  }
  GemmEpi e1{th + t.ob1, nullptr, 0, 0, 1.f};
  int rc = gemm(h, false, false, (int)R, c->H1, t.in1, w.X, t.in1, th + t.oW1, c->H1, w.Z1, c->H1,
                e1, st);
  if (rc) return rc;
  GemmEpi e2{th + t.ob2, nullptr, 0, 0, 1.f};
  if (tin) {
    e2.reluA = 1;
    rc = gemm(h, false, false, (int)R, c->H2, c->H1, w.Z1, c->H1, th + t.oW2, c->H2, w.Z2, c->H2,
              e2, st);
  } else {
    const long long n = R * (c->H1 + c->A);
    k_build_zc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.Z1, a, R, c->H1, c->A, w.ZC);
    RLC_LAUNCH_CHECK(h);
    rc = gemm(h, false, false, (int)R, c->H2, c->H1 + c->A, w.ZC, c->H1 + c->A, th + t.oW2, c->H2,
              w.Z2, c->H2, e2, st);
  }
  if (rc) return rc;
  k_head<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(w.Z2, R, c->H2, th + t.ow3, th + t.ob3,
                                                           y, inv_btotal, q_out, w.dq, w.G2,
                                                           loss_acc);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// dQ/da on R stacked rows whose states repeat `rep` times (s has R/rep rows): rep = 1 is the plain
// stacked form, rep = N the un-materialised B x N form (SQL's SVGD particles, sql_network.py:101-107).
int rlc_critic_grad_action_rep(rlc_handle* h, const rlc_critic* c, const float* s, int rep,
                               const float* a, long long R, float* dqda_out, float* q_out,
                               cudaStream_t st) {
  if (R == 0) return RLC_OK;
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  const float* th = c->theta;
  const long long CH = 16384LL / rep * rep > 0 ? 16384LL / rep * rep : rep;  // row chunk (a multiple of rep) bounds the workspace
  void* ws = nullptr;
  int rc = rlc_workspace(h, train_ws_bytes(c, R < CH ? R : CH), &ws);
  if (rc) return rc;
  for (long long r0 = 0; r0 < R; r0 += CH) {
    const long long n = (R - r0 < CH) ? (R - r0) : CH;
    TrainWs w = carve(c, n, (float*)ws);
    rc = forward_rows(h, c, s + (r0 / rep) * c->S, a + r0 * c->A, n, nullptr, 0.f, w,
                      q_out ? q_out + r0 : nullptr, nullptr, st, rep);
    if (rc) return rc;
    GemmEpi e{nullptr, nullptr, 0, 0, 1.f};
    if (c->topology == RLC_TMID) {
      // dA = G2 * W2a^T, W2a = rows H1.. of W2 ([A,H2])
      rc = gemm(h, false, true, (int)n, c->A, c->H2, w.G2, c->H2,
                th + t.oW2 + (int64_t)c->H1 * c->H2, c->H2, dqda_out + r0 * c->A, c->A, e, st);
    } else {
      GemmEpi em{nullptr, w.Z1, c->H1, 0, 1.f};
      rc = gemm(h, false, true, (int)n, c->H1, c->H2, w.G2, c->H2, th + t.oW2, c->H2, w.G1, c->H1,
                em, st);
      if (rc) return rc;
      // dA = G1 * W1a^T, W1a = rows S.. of W1 ([A,H1])
      rc = gemm(h, false, true, (int)n, c->A, c->H1, w.G1, c->H1,
                th + t.oW1 + (int64_t)c->S * c->H1, c->H1, dqda_out + r0 * c->A, c->A, e, st);
    }
    if (rc) return rc;
  }
  return RLC_OK;
}

extern "C" int rlc_critic_grad_action(rlc_handle* h, const rlc_critic* c, const float* s,
                                      const float* a, int R, float* dqda_out, float* q_out,
                                      void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && a && dqda_out && R >= 0);
  return rlc_critic_grad_action_rep(h, c, s, 1, a, R, dqda_out, q_out, (cudaStream_t)stream);
}

extern "C" int rlc_critic_grads(rlc_handle* h, const rlc_critic* c, const float* s, const float* a,
                                const float* y, int B, int B_total, float* grad_out,
                                float* loss_out, float* q_out, void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && s && a && y && grad_out && B >= 1 && B_total >= B);
  cudaStream_t st = (cudaStream_t)stream;
  const ThetaView t = theta_view(c->topology, c->S, c->A, c->H1, c->H2);
  const float* th = c->theta;
  void* ws = nullptr;
  int rc = rlc_workspace(h, train_ws_bytes(c, B), &ws);
  if (rc) return rc;
  TrainWs w = carve(c, B, (float*)ws);
  if (loss_out) RLC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  rc = forward_rows(h, c, s, a, B, y, 1.f / (float)B_total, w, q_out, loss_out, st);
  if (rc) return rc;
  // head grads
  // gw3[j] = sum_r relu(Z2[r,j]) dq[r], gb3 = sum_r dq[r]: theta stores [w3 (H2) | b3 (1)] contiguously
  rc = colred<1>(h, w.Z2, w.dq, B, c->H2, c->H2, w.part, grad_out + t.ow3, st);
  if (rc) return rc;
  // [gW2 ; gb2] = [relu(Z1) | 1]^T G2 (theta stores W2 and b2 contiguously)
  GemmEpi e{nullptr, nullptr, 0, 0, 1.f};
  if (c->topology == RLC_TIN) {
    GemmEpi er = e;
    er.reluA = 1;
    rc = gemm_splitk_bias(h, c->H1, c->H2, B, w.Z1, c->H1, w.G2, c->H2, grad_out + t.oW2, er, w.slabs, w.part, st);
  } else {
    rc = gemm_splitk_bias(h, c->H1 + c->A, c->H2, B, w.ZC, c->H1 + c->A, w.G2, c->H2, grad_out + t.oW2, e, w.slabs,
                          w.part, st);
  }
  if (rc) return rc;
  // G1 = (G2 W2[:H1]^T) * [Z1>0]
  GemmEpi em{nullptr, w.Z1, c->H1, 0, 1.f};
  rc = gemm(h, false, true, B, c->H1, c->H2, w.G2, c->H2, th + t.oW2, c->H2, w.G1, c->H1, em, st);
  if (rc) return rc;
  // [gW1 ; gb1] = [X | 1]^T G1
  return gemm_splitk_bias(h, t.in1, c->H1, B, w.X, t.in1, w.G1, c->H1, grad_out + t.oW1, e, w.slabs, w.part, st);
}

// T-mid dQ/da over a B x N block without materialising the stack (AE+ ascent, ae_plus_network.py:
// 310-343, evaluates B*N rows whose states repeat): exposed through rlc_critic_eval's sibling.
extern "C" int rlc_tmid_eval_grad(rlc_handle* h, const rlc_critic* c, const float* s, int B,
                                  const float* a, int N, int act_mode, float* q_out,
                                  float* dqda_out, void* stream) {
  RLC_REQUIRE(h && critic_ok(c) && c->topology == RLC_TMID && s && a && dqda_out && B >= 0 && N >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  if ((long long)B * N == 0) return RLC_OK;
  void* ws = nullptr;
  const size_t np = ((size_t)B * c->H2 + 3) & ~(size_t)3;
  int rc = rlc_workspace(h, (np + rlc_tmid_state_scratch_floats(c, B)) * sizeof(float), &ws);
  if (rc) return rc;
  rc = rlc_tmid_state_term(h, c, s, B, (float*)ws, st, (float*)ws + np);   // dense state batches: tensor-core GEMMs
  if (rc) return rc;
  return launch_tmid_rows<true>(h, c, (const float*)ws, a, act_mode == RLC_ACT_PER_STATE,
                                (long long)B * N, N, q_out, dqda_out, st);
}

// =============================================================================================
// theta pack / unpack, Adam, soft update
// =============================================================================================
// dst[in,out] <- src (layout OUT_IN: src[out,in] ; IN_OUT: copy). unpack = reverse.
__global__ void k_repack(const float* __restrict__ src, float* __restrict__ dst, int in, int out,
                         int src_is_out_in, int reverse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)in * out) return;
  const int r = (int)(i / out), c = (int)(i - (long long)r * out);  // canonical [in=r][out=c]
  const long long ext = src_is_out_in ? ((long long)c * in + r) : i;
  if (!reverse) dst[i] = src[ext];
  else dst[ext] = src[i];
}

static int repack(const float* src, float* dst, int in, int out, int layout, int reverse,
                  cudaStream_t st) {
  const long long n = (long long)in * out;
  k_repack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, in, out,
                                                        layout == RLC_LAYOUT_OUT_IN, reverse);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return rlc_cuda_fail(e, __FILE__, __LINE__);
  return RLC_OK;
}

extern "C" int rlc_pack_theta(int topology, int S, int A, int H1, int H2, int layout,
                              const float* W1, const float* b1, const float* W2, const float* b2,
                              const float* W3, const float* b3, float* theta, void* stream) {
  RLC_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3 && theta);
  RLC_REQUIRE(topology == RLC_TIN || topology == RLC_TMID);
  RLC_REQUIRE(layout == RLC_LAYOUT_OUT_IN || layout == RLC_LAYOUT_IN_OUT);
  cudaStream_t st = (cudaStream_t)stream;
  const ThetaView t = theta_view(topology, S, A, H1, H2);
  int rc = repack(W1, theta + t.oW1, t.in1, H1, layout, 0, st);
  if (rc) return rc;
  rc = repack(W2, theta + t.oW2, t.in2, H2, layout, 0, st);
  if (rc) return rc;
  RLC_CUDA(cudaMemcpyAsync(theta + t.ob1, b1, H1 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(theta + t.ob2, b2, H2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(theta + t.ow3, W3, H2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(theta + t.ob3, b3, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return RLC_OK;
}

extern "C" int rlc_unpack_theta(int topology, int S, int A, int H1, int H2, int layout,
                                const float* theta, float* W1, float* b1, float* W2, float* b2,
                                float* W3, float* b3, void* stream) {
  RLC_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3 && theta);
  RLC_REQUIRE(topology == RLC_TIN || topology == RLC_TMID);
  RLC_REQUIRE(layout == RLC_LAYOUT_OUT_IN || layout == RLC_LAYOUT_IN_OUT);
  cudaStream_t st = (cudaStream_t)stream;
  const ThetaView t = theta_view(topology, S, A, H1, H2);
  int rc = repack(theta + t.oW1, W1, t.in1, H1, layout, 1, st);
  if (rc) return rc;
  rc = repack(theta + t.oW2, W2, t.in2, H2, layout, 1, st);
  if (rc) return rc;
  RLC_CUDA(cudaMemcpyAsync(b1, theta + t.ob1, H1 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(b2, theta + t.ob2, H2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(W3, theta + t.ow3, H2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RLC_CUDA(cudaMemcpyAsync(b3, theta + t.ob3, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return RLC_OK;
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float lr_eff, float b1, float b2,
                       float eps, float inv_sqrt_bc2, float* __restrict__ target, float tau) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float pn = p[i] - lr_eff * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  p[i] = pn;
  if (target) target[i] += tau * (pn - target[i]);
}

extern "C" int rlc_adam_step(rlc_handle* h, float* theta, const float* grad, float* m, float* v,
                             int64_t n, int step, float lr, float beta1, float beta2, float eps,
                             int variant, float* target, float tau, void* stream) {
  RLC_REQUIRE(h && theta && grad && m && v && n >= 0 && step >= 1);
  RLC_REQUIRE(variant == RLC_ADAM_TORCH || variant == RLC_ADAM_TF);
  if (n == 0) return RLC_OK;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  float lr_eff, isb2;
  if (variant == RLC_ADAM_TORCH) {  // denom = sqrt(v)/sqrt(bc2) + eps ; step = lr/bc1
    lr_eff = (float)(lr / bc1);
    isb2 = (float)(1.0 / sqrt(bc2));
  } else {  // lr_t = lr*sqrt(bc2)/bc1 ; denom = sqrt(v) + eps
    lr_eff = (float)(lr * sqrt(bc2) / bc1);
    isb2 = 1.f;
  }
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      theta, grad, m, v, n, lr_eff, beta1, beta2, eps, isb2, target, tau);
  RLC_LAUNCH_CHECK(h);
  rlc_invalidate_pack(h, theta);
  if (target) rlc_invalidate_pack(h, target);
  return RLC_OK;
}

// CUDA-graph-safe Adam: the step count lives on the device (state_dev[0]) so a captured update
// advances it on every replay; a one-thread kernel bumps it and derives the two scalar factors.
__global__ void k_adam_prep(int* state_dev, float lr, float b1, float b2, int variant) {
  const int t = ++state_dev[0];
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  float lr_eff, isb2;
  if (variant == RLC_ADAM_TORCH) {
    lr_eff = (float)((double)lr / bc1);
    isb2 = (float)(1.0 / sqrt(bc2));
  } else {
    lr_eff = (float)((double)lr * sqrt(bc2) / bc1);
    isb2 = 1.f;
  }
  reinterpret_cast<float*>(state_dev)[1] = lr_eff;
  reinterpret_cast<float*>(state_dev)[2] = isb2;
}

__global__ void k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, long long n, const int* __restrict__ state_dev, float b1,
                           float b2, float eps, float* __restrict__ target, float tau) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_eff = reinterpret_cast<const float*>(state_dev)[1];
  const float inv_sqrt_bc2 = reinterpret_cast<const float*>(state_dev)[2];
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float pn = p[i] - lr_eff * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  p[i] = pn;
  if (target) target[i] += tau * (pn - target[i]);
}

extern "C" int rlc_adam_step_dev(rlc_handle* h, float* theta, const float* grad, float* m, float* v,
                                 int64_t n, int32_t* state_dev, float lr, float beta1, float beta2,
                                 float eps, int variant, float* target, float tau, void* stream) {
  RLC_REQUIRE(h && theta && grad && m && v && state_dev && n >= 0);
  RLC_REQUIRE(variant == RLC_ADAM_TORCH || variant == RLC_ADAM_TF);
  if (n == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_prep<<<1, 1, 0, st>>>(state_dev, lr, beta1, beta2, variant);
  RLC_LAUNCH_CHECK(h);
  k_adam_dev<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(theta, grad, m, v, n, state_dev, beta1, beta2, eps,
                                                         target, tau);
  RLC_LAUNCH_CHECK(h);
  rlc_invalidate_pack(h, theta);
  if (target) rlc_invalidate_pack(h, target);
  return RLC_OK;
}

__global__ void k_soft(float* __restrict__ t, const float* __restrict__ o, long long n, float tau) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] += tau * (o[i] - t[i]);
}

extern "C" int rlc_soft_update(rlc_handle* h, float* target, const float* online, int64_t n,
                               float tau, void* stream) {
  RLC_REQUIRE(h && target && online && n >= 0);
  if (n == 0) return RLC_OK;
  k_soft<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(target, online, n, tau);
  RLC_LAUNCH_CHECK(h);
  rlc_invalidate_pack(h, target);
  return RLC_OK;
}
