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
// CACHED: the row lives in registers (N <= 32*NPL), so the k rounds re-read nothing -- the streaming
// variant re-fetches the row from L2 every round (6 x 16.8 MB at cfg4: 40 us instead of 7).
template <int NPL>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_topk(const float* __restrict__ q, int B, int N, int k, long long* __restrict__ idx_out,
       float* __restrict__ q_sel_out, const float* __restrict__ actions, int A, int act_per_state,
       float* __restrict__ elites_out) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = q + (long long)b * N;
  float rv[NPL > 0 ? NPL : 1];
  if (NPL > 0) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int n = lane + 32 * i;
      float v = n < N ? row[n] : -CUDART_INF_F;
      if (v != v) v = CUDART_INF_F;  // NaN ordered like +inf (numpy sorts NaN last ascending)
      rv[i] = v;
    }
  }
  float lastv = CUDART_INF_F;
  int lasti = 0x7fffffff;
  for (int t = 0; t < k; ++t) {
    float bv = -CUDART_INF_F;
    int bi = -1;
    if (NPL > 0) {
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int n = lane + 32 * i;
        const float v = rv[i];
        const bool below = (n < N) && ((v < lastv) || (v == lastv && n < lasti));
        if (below && (bi < 0 || key_gt(v, n, bv, bi))) { bv = v; bi = n; }
      }
    } else {
#pragma unroll 8
      for (int n = lane; n < N; n += 32) {
        float v = row[n];
        if (v != v) v = CUDART_INF_F;
        const bool below = (v < lastv) || (v == lastv && n < lasti);
        if (below && (bi < 0 || key_gt(v, n, bv, bi))) { bv = v; bi = n; }
      }
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
#pragma unroll 8
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
#pragma unroll 8
  for (int n = lane; n < N; n += 32) m = fmaxf(m, row[n]);
  m = warp_max(m);
  float s = 0.f;
#pragma unroll 8
  for (int n = lane; n < N; n += 32) s += expf(row[n] - m);
  s = warp_sum(s);
  if (lane == 0) v_out[b] = m + logf(s) + offset;
}

// ForwardKL (forwardkl_network.py:165-194).  NPL > 0: the scaled row q/alpha lives in registers
// (N <= 32*NPL): one IEEE division and one global read per element instead of three of each.
template <int NPL>
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
  float t[NPL > 0 ? NPL : 1];
  if (NPL > 0) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int n = lane + 32 * i;
      t[i] = n < N ? __fdiv_rn(row[n], alpha) : -CUDART_INF_F;
      m = fmaxf(m, t[i]);
    }
  } else {
#pragma unroll 8
    for (int n = lane; n < N; n += 32) m = fmaxf(m, __fdiv_rn(row[n], alpha));
  }
  m = warp_max(m);
  float z = 0.f;
  if (NPL > 0) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int n = lane + 32 * i;
      if (n < N) {
        t[i] = expf(t[i] - m);
        z = fmaf(t[i], w[n], z);
      }
    }
  } else {
#pragma unroll 8
    for (int n = lane; n < N; n += 32) z = fmaf(expf(__fdiv_rn(row[n], alpha) - m), w[n], z);
  }
  z = warp_sum(z);
  float acc = 0.f;
  if (NPL > 0) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int n = lane + 32 * i;
      if (n < N) {
        const float p = __fdiv_rn(t[i], z);
        const float pw = p * w[n];
        acc = fmaf(pw, lp[n], acc);
        if (boltz) boltz[(long long)b * N + n] = p;
        if (dlogp) dlogp[(long long)b * N + n] = -pw * inv_btotal;
      }
    }
  } else {
#pragma unroll 8
    for (int n = lane; n < N; n += 32) {
      const float p = __fdiv_rn(expf(__fdiv_rn(row[n], alpha) - m), z);
      const float pw = p * w[n];
      acc = fmaf(pw, lp[n], acc);
      if (boltz) boltz[(long long)b * N + n] = p;
      if (dlogp) dlogp[(long long)b * N + n] = -pw * inv_btotal;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) loss_b[b] = -acc;
}

// ForwardKL with the row in shared memory (one warp = one state).  ncu on the streaming variant: 15 % issue-active,
// 32 warps stalled on the long scoreboard per issue -- one dependent HBM round trip after another (three passes, eight
// scalar loads in flight per lane).  Here a lane issues ALL its loads of q and logp up front as 128-bit loads (16 in
// flight at N = 1024), divides once, and every later pass reads shared memory; outputs leave as 128-bit stores.
// Same arithmetic per element; the per-lane summation order follows the float4 layout (n = 4 (lane + 32 i) + c).
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_fkl_smem(const float* __restrict__ q, const float* __restrict__ w, const float* __restrict__ logp,
           int B, int N, float alpha, float inv_btotal, float* __restrict__ loss_b,
           float* __restrict__ boltz, float* __restrict__ dlogp) {
  extern __shared__ __align__(16) float fkl_rows[];
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int N4 = N >> 2;                                       // N % 4 == 0 (host-checked)
  float4* t4 = reinterpret_cast<float4*>(fkl_rows + (size_t)(threadIdx.x >> 5) * 2 * N);
  float4* l4 = t4 + N4;
  const float4* row4 = reinterpret_cast<const float4*>(q + (long long)b * N);
  const float4* lp4 = reinterpret_cast<const float4*>(logp + (long long)b * N);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float m = -CUDART_INF_F;
#pragma unroll 8
  for (int i = lane; i < N4; i += 32) {
    float4 v = __ldg(row4 + i);
    const float4 l = __ldg(lp4 + i);
    v.x = __fdiv_rn(v.x, alpha); v.y = __fdiv_rn(v.y, alpha); v.z = __fdiv_rn(v.z, alpha); v.w = __fdiv_rn(v.w, alpha);
    t4[i] = v;
    l4[i] = l;
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = warp_max(m);
  float z = 0.f;
#pragma unroll 4
  for (int i = lane; i < N4; i += 32) {        // a lane re-reads only what it wrote: no barrier needed
    float4 e = t4[i];
    const float4 ww = __ldg(w4 + i);
    e.x = expf(e.x - m); e.y = expf(e.y - m); e.z = expf(e.z - m); e.w = expf(e.w - m);
    t4[i] = e;
    z = fmaf(e.x, ww.x, z); z = fmaf(e.y, ww.y, z); z = fmaf(e.z, ww.z, z); z = fmaf(e.w, ww.w, z);
  }
  z = warp_sum(z);
  float acc = 0.f;
  float4* bz4 = boltz ? reinterpret_cast<float4*>(boltz + (long long)b * N) : nullptr;
  float4* dl4 = dlogp ? reinterpret_cast<float4*>(dlogp + (long long)b * N) : nullptr;
#pragma unroll 4
  for (int i = lane; i < N4; i += 32) {
    const float4 e = t4[i], l = l4[i], ww = __ldg(w4 + i);
    float4 p, pw;
    p.x = __fdiv_rn(e.x, z); p.y = __fdiv_rn(e.y, z); p.z = __fdiv_rn(e.z, z); p.w = __fdiv_rn(e.w, z);
    pw.x = p.x * ww.x; pw.y = p.y * ww.y; pw.z = p.z * ww.z; pw.w = p.w * ww.w;
    acc = fmaf(pw.x, l.x, acc); acc = fmaf(pw.y, l.y, acc); acc = fmaf(pw.z, l.z, acc); acc = fmaf(pw.w, l.w, acc);
    if (bz4) bz4[i] = p;
    if (dl4) dl4[i] = make_float4(-pw.x * inv_btotal, -pw.y * inv_btotal, -pw.z * inv_btotal, -pw.w * inv_btotal);
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
#pragma unroll 8
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

// ---- policy-fused grid reductions (rows a3/a4 + a5) -------------------------------------------
// PolicyNetwork.get_logprob (forwardkl_network.py:324-351) is evaluated in place from the policy
// head outputs mean[B,A], log_std[B,A]: no [B,N] log-density tensor is read, and the gradient comes
// back as dL/dmean, dL/dlog_std [B,A].  Terms that depend on the grid only are tabulated once:
//   U[k][n] = atanh(a_nk/scale) = (log(1+x) - log(1-x))/2     J[n] = sum_k log(1 - x^2 + eps)
// (dimension-major: lane n reads U[k][n], one coalesced wavefront per load instead of A)
__global__ void k_grid_logterms(const float* __restrict__ grid, int N, int A, float inv_scale, float eps,
                                float* __restrict__ U, float* __restrict__ J) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float j = 0.f;
  for (int k = 0; k < A; ++k) {
    const float x = grid[(size_t)n * A + k] * inv_scale;
    U[(size_t)k * N + n] = (logf(1.f + x) - logf(1.f - x)) * 0.5f;
    j += logf(1.f - x * x + eps);
  }
  J[n] = j;
}

#define POL_MAX_A 8
// 64 threads x 16 cached values per state.  (The first version used 256 threads x 4: ncu showed it
// instruction-bound at 224 instructions per (state, action) pair, half of them the per-thread prologue
// -- 12 IEEE divisions + 6 expf for the log-density constants -- and the 13 final warp reductions,
// amortised over only 4 pairs.)
#define POL_THREADS 64
#define POL_NPT 16   // q values cached per thread (N <= 1024 stays in registers)

// block-wide reductions over POL_THREADS threads (8 warps); `red` is 8 floats of shared memory
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();                                   // protect `red` from the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < POL_THREADS / 32; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < POL_THREADS / 32; ++i) r += red[i];     // fixed order: deterministic
  return r;
}

// One CTA per state (256 threads x up to 4 grid actions each in registers): the N-way reductions are
// block-wide, so B x N = 4 M pairs expose 1 M threads of parallelism instead of 4096 serial warps.
// MODE 0: ForwardKL (Boltzmann weights over the grid, detached); MODE 1: ReverseKL.
template <int MODE, int A>
__global__ void __launch_bounds__(POL_THREADS)
k_policy_reduce(const float* __restrict__ q, const float* __restrict__ v, const float* __restrict__ w,
                const float* __restrict__ U, const float* __restrict__ J, const float* __restrict__ mean,
                const float* __restrict__ log_std, int B, int N, float alpha, float inv_btotal,
                float* __restrict__ loss_b, float* __restrict__ dmean, float* __restrict__ dlstd,
                float* __restrict__ logp_out) {
  __shared__ float red[POL_THREADS / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* row = q + (long long)b * N;
  // per-state constants of the log-density (A == 1: Normal(mean, std); A > 1: the reference hands
  // std to MultivariateNormal as the covariance, forwardkl_network.py:350)
  float mu[A], h[A], gs[A];   // h: coefficient of d^2 ; gs: d^2 factor in dlogp/dlog_std
  float c0 = 0.f;
  // the constants are computed once per state (A lanes) and broadcast through shared memory
  __shared__ float cst[3 * POL_MAX_A + 1];
  if (tid < A) {
    const float ls = log_std[(size_t)b * A + tid];
    const float sd = expf(ls);
    cst[tid] = mean[(size_t)b * A + tid];
    cst[POL_MAX_A + tid] = (A == 1) ? 0.5f / (sd * sd) : 0.5f / sd;
    cst[2 * POL_MAX_A + tid] = (A == 1) ? 1.f / (sd * sd) : 0.5f / sd;
  }
  if (tid == 0) {
    float c = 0.f;
    if (A == 1) {
      c = -log_std[(size_t)b * A] - 0.9189385332046727f;     // log sqrt(2 pi)
    } else {
      for (int k = 0; k < A; ++k) c -= 0.5f * log_std[(size_t)b * A + k];
      c -= 0.9189385332046727f * (float)A;
    }
    cst[3 * POL_MAX_A] = c;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < A; ++k) {
    mu[k] = cst[k];
    h[k] = cst[POL_MAX_A + k];
    gs[k] = cst[2 * POL_MAX_A + k];
  }
  c0 = cst[3 * POL_MAX_A];
  const bool cached = N <= POL_THREADS * POL_NPT;
  float qc[POL_NPT];
  if (cached) {
#pragma unroll
    for (int i = 0; i < POL_NPT; ++i) {
      const int n = tid + POL_THREADS * i;
      qc[i] = n < N ? row[n] : 0.f;
    }
  }
  const float inv_alpha = 1.f / alpha;   // q/alpha as a multiply: 1 ulp from the reference's division, far inside the tolerance
  float m = 0.f, z = 1.f;
  if (MODE == 0) {
    m = -CUDART_INF_F;
    if (cached) {
#pragma unroll
      for (int i = 0; i < POL_NPT; ++i)
        if (tid + POL_THREADS * i < N) m = fmaxf(m, qc[i] * inv_alpha);
    } else {
      for (int n = tid; n < N; n += POL_THREADS) m = fmaxf(m, row[n] * inv_alpha);
    }
    m = block_max(m, red);
    z = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < POL_NPT; ++i) {
        const int n = tid + POL_THREADS * i;
        if (n < N) {
          qc[i] = expf(qc[i] * inv_alpha - m);        // reuse the register for e_n
          z = fmaf(qc[i], w[n], z);
        }
      }
    } else {
      for (int n = tid; n < N; n += POL_THREADS) z = fmaf(expf(row[n] * inv_alpha - m), w[n], z);
    }
    z = block_sum(z, red);
  }
  const float inv_z = 1.f / z;
  const float vb = (MODE == 1) ? v[b] : 0.f;
  float acc = 0.f, gm[A], gl[A];
#pragma unroll
  for (int k = 0; k < A; ++k) { gm[k] = 0.f; gl[k] = 0.f; }
  auto body = [&](int n, float qe) {
    // log pi(a_n | s_b)
    float lp = c0 - J[n];
    float d[A];
#pragma unroll
    for (int k = 0; k < A; ++k) {
      d[k] = U[(size_t)k * N + n] - mu[k];
      lp = fmaf(-h[k] * d[k], d[k], lp);
    }
    float g;   // dL/dlogp_bn
    if (MODE == 0) {
      const float pw = qe * inv_z * w[n];                   // Boltzmann weight x quadrature weight
      acc = fmaf(pw, lp, acc);
      g = -pw * inv_btotal;
    } else {
      const float pe = expf(lp);
      const float inner = (qe - vb) - alpha * lp;
      acc = fmaf(-pe * inner, w[n], acc);
      g = (-pe * (inner - alpha)) * w[n] * inv_btotal;
    }
#pragma unroll
    for (int k = 0; k < A; ++k) {
      gm[k] = fmaf(g, 2.f * h[k] * d[k], gm[k]);
      gl[k] = fmaf(g, (A == 1) ? (gs[k] * d[k] * d[k] - 1.f) : (gs[k] * d[k] * d[k] - 0.5f), gl[k]);
    }
    if (logp_out) logp_out[(long long)b * N + n] = lp;
  };
  if (cached) {
#pragma unroll
    for (int i = 0; i < POL_NPT; ++i) {
      const int n = tid + POL_THREADS * i;
      if (n < N) body(n, qc[i]);
    }
  } else {
    for (int n = tid; n < N; n += POL_THREADS)
      body(n, MODE == 0 ? expf(row[n] * inv_alpha - m) : row[n]);
  }
  // the 2A+1 output sums in one pass: warp shuffles, one shared-memory exchange, fixed-order final sum
  __shared__ float fin[POL_THREADS / 32][2 * POL_MAX_A + 1];
  acc = warp_sum(acc);
#pragma unroll
  for (int k = 0; k < A; ++k) { gm[k] = warp_sum(gm[k]); gl[k] = warp_sum(gl[k]); }
  if ((tid & 31) == 0) {
    fin[tid >> 5][0] = acc;
#pragma unroll
    for (int k = 0; k < A; ++k) { fin[tid >> 5][1 + k] = gm[k]; fin[tid >> 5][1 + A + k] = gl[k]; }
  }
  __syncthreads();
  if (tid < 2 * A + 1) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < POL_THREADS / 32; ++i) r += fin[i][tid];
    if (tid == 0) loss_b[b] = (MODE == 0) ? -r : r;
    else if (tid <= A) { if (dmean) dmean[(size_t)b * A + (tid - 1)] = r; }
    else { if (dlstd) dlstd[(size_t)b * A + (tid - 1 - A)] = r; }
  }
}

static inline unsigned nblocks(int B) { return (unsigned)((B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

extern "C" int rlc_reduce_topk(rlc_handle* h, const float* q, int B, int N, int k,
                               int64_t* idx_out, float* q_sel_out, const float* actions, int A,
                               int act_mode, float* elites_out, void* stream) {
  RLC_REQUIRE(h && q && idx_out && B >= 0 && N >= 1 && k >= 1 && k <= 64 && k <= N);
  RLC_REQUIRE(!elites_out || (actions && A >= 1));
  if (B == 0) return RLC_OK;
  if (N <= 128) k_topk<4><<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, B, N, k, (long long*)idx_out, q_sel_out, actions, A, act_mode == RLC_ACT_PER_STATE,
      elites_out);
  else if (N <= 512) k_topk<16><<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, B, N, k, (long long*)idx_out, q_sel_out, actions, A, act_mode == RLC_ACT_PER_STATE,
      elites_out);
  else if (N <= 1024) k_topk<32><<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      q, B, N, k, (long long*)idx_out, q_sel_out, actions, A, act_mode == RLC_ACT_PER_STATE,
      elites_out);
  else k_topk<0><<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
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
#define RLC_FKL(NPL_)                                                          \
  k_fkl<NPL_><<<nblocks(B), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>( \
      q, w, logp, B, N, entropy_scale, 1.f / (float)B_total, loss_b_out, boltz_out, dlogp_out)
  // measured on B200 at B=4096, N=1024: register-cached (NPL=32, 93 registers) 69 us, streaming 47 us, shared-memory
  // row cache (rows up to 1536 floats at 4 warps per CTA, 16-byte aligned) the fastest; small rows keep the register path
  const size_t row_smem = (size_t)WARPS_PER_BLOCK * 2 * N * sizeof(float);
  const bool vec_ok = (N & 3) == 0 && ((((uintptr_t)q) | ((uintptr_t)logp) | ((uintptr_t)w) | ((uintptr_t)boltz_out) |
                                        ((uintptr_t)dlogp_out)) & 15) == 0;
  if (N <= 128) RLC_FKL(4);
  else if (vec_ok && row_smem <= 48 * 1024)
    k_fkl_smem<<<nblocks(B), 32 * WARPS_PER_BLOCK, row_smem, (cudaStream_t)stream>>>(
        q, w, logp, B, N, entropy_scale, 1.f / (float)B_total, loss_b_out, boltz_out, dlogp_out);
  else RLC_FKL(0);
#undef RLC_FKL
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

// U[k][n], J[n] tables of the grid (used by the fused evaluation + reduction of critic_umma_grid3.cuh as well)
int rlc_launch_grid_logterms(rlc_handle* h, const float* grid, int N, int A, float action_scale, float* U, float* J,
                             cudaStream_t st) {
  k_grid_logterms<<<(N + 127) / 128, 128, 0, st>>>(grid, N, A, 1.f / action_scale, 1e-6f, U, J);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

static int policy_reduce(rlc_handle* h, int mode, const float* q, const float* v, const float* w,
                         const float* grid, int A, float action_scale, const float* mean,
                         const float* log_std, int B, int N, float alpha, int B_total, float* loss_b_out,
                         float* dmean_out, float* dlog_std_out, float* logp_out, cudaStream_t st) {
  RLC_REQUIRE(h && q && w && grid && mean && log_std && loss_b_out && B >= 0 && N >= 1 && A >= 1 &&
              A <= POL_MAX_A && B_total >= B && B_total >= 1 && action_scale > 0.f);
  if (B == 0) return RLC_OK;
  void* ws = nullptr;
  int rc = rlc_workspace(h, ((size_t)N * A + N) * sizeof(float), &ws);
  if (rc) return rc;
  float* U = (float*)ws;
  float* J = U + (size_t)N * A;
  k_grid_logterms<<<(N + 127) / 128, 128, 0, st>>>(grid, N, A, 1.f / action_scale, 1e-6f, U, J);
  RLC_LAUNCH_CHECK(h);
#define POL_LAUNCH(MODE_, A_)                                                                               \
  k_policy_reduce<MODE_, A_><<<(unsigned)B, POL_THREADS, 0, st>>>(                                          \
      q, v, w, U, J, mean, log_std, B, N, alpha, 1.f / (float)B_total, loss_b_out, dmean_out, dlog_std_out, \
      logp_out)
#define POL_SWITCH(MODE_)                    \
  switch (A) {                               \
    case 1: POL_LAUNCH(MODE_, 1); break;     \
    case 2: POL_LAUNCH(MODE_, 2); break;     \
    case 3: POL_LAUNCH(MODE_, 3); break;     \
    case 4: POL_LAUNCH(MODE_, 4); break;     \
    case 5: POL_LAUNCH(MODE_, 5); break;     \
    case 6: POL_LAUNCH(MODE_, 6); break;     \
    case 7: POL_LAUNCH(MODE_, 7); break;     \
    default: POL_LAUNCH(MODE_, 8); break;    \
  }
  if (mode == 0) { POL_SWITCH(0) } else { POL_SWITCH(1) }
#undef POL_SWITCH
#undef POL_LAUNCH
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_reduce_fkl_policy(rlc_handle* h, const float* q, const float* w, const float* grid, int A,
                                     float action_scale, const float* mean, const float* log_std, int B, int N,
                                     float entropy_scale, int B_total, float* loss_b_out, float* dmean_out,
                                     float* dlog_std_out, float* logp_out, void* stream) {
  RLC_REQUIRE(entropy_scale > 0.f);
  return policy_reduce(h, 0, q, nullptr, w, grid, A, action_scale, mean, log_std, B, N, entropy_scale, B_total,
                       loss_b_out, dmean_out, dlog_std_out, logp_out, (cudaStream_t)stream);
}

extern "C" int rlc_reduce_rkl_policy(rlc_handle* h, const float* q, const float* v, const float* w,
                                     const float* grid, int A, float action_scale, const float* mean,
                                     const float* log_std, int B, int N, float entropy_scale, int hard,
                                     int B_total, float* loss_b_out, float* dmean_out, float* dlog_std_out,
                                     float* logp_out, void* stream) {
  RLC_REQUIRE(v);
  return policy_reduce(h, 1, q, v, w, grid, A, action_scale, mean, log_std, B, N, hard ? 0.f : entropy_scale,
                       B_total, loss_b_out, dmean_out, dlog_std_out, logp_out, (cudaStream_t)stream);
}
