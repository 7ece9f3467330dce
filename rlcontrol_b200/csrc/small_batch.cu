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
#include <stdlib.h>

#include "common.cuh"
#include "policy_math.cuh"

#define SB_ROWS 16       // rows per forward CTA
#define SB_LD 20         // floats per k in the activation tiles ([K][SB_LD], rows 0..15 used): 5 quads per row, so float4
                         // stores of consecutive columns by consecutive lanes fall in distinct bank groups
#define SB_THREADS 256
#define SB_NI 8          // hidden-1 units per backward CTA (more, smaller CTAs: every phase that scales with it is latency)
#define SB_KC 16         // weight rows per ring stage
#define SB_NST 4         // ring stages (bulk copies in flight: SB_NST * SB_KC * J * 4 bytes, 51 KB at J = 200)
#define SB_WAIT_LIMIT (1u << 26)

struct SbFwdArgs {
  rlc_sb_net net[RLC_SB_MAX_NETS];
  int cta_base[RLC_SB_MAX_NETS + 1];
  int resident[RLC_SB_MAX_NETS];   // W1 and W2 fit in shared memory whole (and are 16-byte aligned): no ring, no per-chunk barriers
  int n_nets, B;
  unsigned long long* dbg;   // RLC_SB_DEBUG: %globaltimer trace of CTA 0 / thread 0 (debug builds of the timing scripts)
};
__device__ __forceinline__ unsigned long long sb_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SB_TRACE(slot) do { if (dbg && threadIdx.x == 0 && blockIdx.x == 0) dbg[slot] = sb_now(); } while (0)
struct SbUpdArgs {
  rlc_sb_train net[RLC_SB_MAX_NETS];
  int cta_base[RLC_SB_MAX_NETS + 1];
  int n_nets, B;
  float inv_btotal;
  unsigned long long* dbg;
};

__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool sb_mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: never hangs the GPU
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok && ++spins < SB_WAIT_LIMIT);
  return ok != 0;
}

__device__ __forceinline__ float sb_x(const rlc_sb_net& n, int r, int k) {
  if (k < n.n0) return n.x0[(long long)(n.x0_div ? r / n.x0_div : r) * n.n0 + k];
  return n.x1[(long long)(n.x1_mod ? r % n.x1_mod : r) * n.n1 + (k - n.n0)];
}
__device__ __forceinline__ float sb_xt(const rlc_sb_train& n, int b, int k) {
  return k < n.n0 ? n.x0[(long long)b * n.n0 + k] : n.x1[(long long)b * n.n1 + (k - n.n0)];
}

// The weights of one layer, W[K][J] row-major in HBM/L2, streamed through a ring of SB_NST shared-memory stages of
// SB_KC rows each.  When every row segment is 16-byte aligned the stages are filled by cp.async.bulk (one elected
// thread, mbarrier complete_tx): ~50 KB in flight per CTA without a register of staging -- the first version read W
// with __ldg inside the k loop and was bound by L2 latency (17.8 us for four 32-row passes).  Odd shapes fall back to
// cooperative loads of the stage at the point of use.
struct SbRing {
  float* buf;        // [SB_NST][SB_KC][jc]
  uint32_t bar0;     // SB_NST mbarriers
  uint32_t uses;     // stage fills so far (running over all layers: parity = (uses / SB_NST) & 1)
};

__device__ __forceinline__ void sb_fma16(float (&acc)[4][4], const float4 a, const float4 w) {
  acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.y, w.x, acc[0][1]);
  acc[0][2] = fmaf(a.z, w.x, acc[0][2]); acc[0][3] = fmaf(a.w, w.x, acc[0][3]);
  acc[1][0] = fmaf(a.x, w.y, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
  acc[1][2] = fmaf(a.z, w.y, acc[1][2]); acc[1][3] = fmaf(a.w, w.y, acc[1][3]);
  acc[2][0] = fmaf(a.x, w.z, acc[2][0]); acc[2][1] = fmaf(a.y, w.z, acc[2][1]);
  acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.w, w.z, acc[2][3]);
  acc[3][0] = fmaf(a.x, w.w, acc[3][0]); acc[3][1] = fmaf(a.y, w.w, acc[3][1]);
  acc[3][2] = fmaf(a.z, w.w, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
}


// (A 3xTF32 mma.sync.m16n8k8 version of this tile was built and measured: legacy mma.sync issues at ~40 cycles per
// instruction here, 512 ns per 16-row chunk against 290-500 ns for the CUDA-core tile -- no gain, removed.)

// out(r, j) = bias[j] + sum_k in_s[k][r] * W[k][j] for the CTA's 16 rows; `in_s` is k-major ([K][SB_LD]).  A thread owns
// a 4-column x 4-row register tile: columns cq, cq+64, cq+128, cq+192 of the 256-column block (consecutive lanes =
// consecutive columns: conflict-free weight reads and coalesced stores) and rows 4rg..4rg+3 (one broadcast float4 per
// k).  History, all measured: one column x 16 rows per thread was bound by the load/store unit's return path (four
// broadcast LDS.128 per k per thread); `if (kk < nk)` predicates kept the compiler from hoisting loads (the full-chunk
// path is branch-free); owning 4 ADJACENT columns made the epilogue's shared-memory stores a 32-way bank conflict
// (2 us per layer) -- with this mapping and SB_LD = 20 a finished column goes out as one conflict-free float4.
// Row groups beyond `nrows` (padding rows of a short pass) do not compute.  f4(j, r0, v) consumes rows r0..r0+3 of
// column j.
template <class F>
__device__ __forceinline__ void sb_layer(const float* __restrict__ W, const float* __restrict__ bias, int K, int J,
                                         const float* __restrict__ in_s, int nrows, SbRing& ring, F f4,
                                         unsigned long long* dbg = nullptr, int dbg0 = 0) {
  const int tid = threadIdx.x, cq = tid & 63, rg = tid >> 6;
  const bool row_active = 4 * rg < nrows;
  for (int jb = 0; jb < J; jb += SB_THREADS) {
    const int jc = min(SB_THREADS, J - jb), js = (jc + 3) & ~3;       // stage row stride (floats)
    const bool bulk = ((((uintptr_t)(W + jb)) & 15) == 0) && ((J & 3) == 0) && ((jc & 3) == 0);
    const int nch = (K + SB_KC - 1) / SB_KC;
    const uint32_t base_use = ring.uses;
    auto issue = [&](int c) {   // thread 0 only
      const int st = (base_use + c) % SB_NST, k0 = c * SB_KC, nk = min(SB_KC, K - k0);
      const uint32_t bar = ring.bar0 + 8 * st;
      float* dst = ring.buf + (size_t)st * SB_KC * SB_THREADS;      // stage rows are dense: stride jc
      sb_mbar_expect_tx(bar, (uint32_t)(nk * jc * 4));
      if (jc == J) {             // the chunk's rows are contiguous in HBM: ONE copy (a single thread issues these)
        sb_bulk_g2s(sb_smem_u32(dst), W + (long long)k0 * J, (uint32_t)(nk * J * 4), bar);
      } else {
        for (int kk = 0; kk < nk; ++kk)
          sb_bulk_g2s(sb_smem_u32(dst + kk * jc), W + (long long)(k0 + kk) * J + jb, (uint32_t)(jc * 4), bar);
      }
    };
    // (no proxy fence before refilling a stage: its previous contents were only READ through the generic proxy, and
    //  the __syncthreads orders those reads before the copy is issued)
    if (bulk && tid == 0)
      for (int c = 0; c < min(SB_NST, nch); ++c) issue(c);
    int col[4];          // this thread's columns inside the block, clamped so that every load is in range
    float acc[4][4];     // [column][row]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      col[q] = min(cq + 64 * q, jc - 1);
      const float bj = bias[jb + col[q]];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[q][r] = bj;
    }
    for (int c = 0; c < nch; ++c) {
      const uint32_t use = base_use + c;
      const int st = use % SB_NST, k0 = c * SB_KC, nk = min(SB_KC, K - k0);
      const float* wst = ring.buf + (size_t)st * SB_KC * SB_THREADS;
      if (bulk) {
        sb_mbar_wait(ring.bar0 + 8 * st, (use / SB_NST) & 1);
        SB_TRACE(dbg0 + 2 * c);
      } else {
        float* wdst = ring.buf + (size_t)st * SB_KC * SB_THREADS;
        for (int i = tid; i < nk * js; i += SB_THREADS) {
          const int kk = i / js, jj = i % js;
          wdst[i] = jj < jc ? __ldg(W + (long long)(k0 + kk) * J + jb + jj) : 0.f;
        }
        __syncthreads();
      }
      if (row_active) {
        const float* ap = in_s + k0 * SB_LD + 4 * rg;
        if (nk == SB_KC) {       // full chunk: branch-free, operands of 4 k steps loaded ahead of their 64 FMAs
#pragma unroll
          for (int kk0 = 0; kk0 < SB_KC; kk0 += 4) {
            float4 w[4], a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float* wr = wst + (kk0 + u) * js;
              w[u] = make_float4(wr[col[0]], wr[col[1]], wr[col[2]], wr[col[3]]);
              a[u] = *reinterpret_cast<const float4*>(ap + (kk0 + u) * SB_LD);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) sb_fma16(acc, a[u], w[u]);
          }
        } else {
          for (int kk = 0; kk < nk; ++kk) {
            const float* wr = wst + kk * js;
            sb_fma16(acc, *reinterpret_cast<const float4*>(ap + kk * SB_LD),
                     make_float4(wr[col[0]], wr[col[1]], wr[col[2]], wr[col[3]]));
          }
        }
      }
      SB_TRACE(dbg0 + 2 * c + 1);
      __syncthreads();   // every thread is done with this stage
      SB_TRACE(dbg0 + 2 * c + 2);
      if (bulk && tid == 0 && c + SB_NST < nch) issue(c + SB_NST);
    }
    ring.uses = bulk ? base_use + nch : base_use;   // the fallback never touches the mbarriers
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (cq + 64 * q < jc)
        f4(jb + cq + 64 * q, 4 * rg,
           row_active ? make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// The same layer with the whole W[K][J] already in shared memory (networks up to ~200-200: 160 KB): no ring, no
// per-chunk barrier (the ring's wait + __syncthreads cost as much as a chunk's arithmetic).  Short passes (<= 4 rows:
// sample_action's single row, the tail of a minibatch) split K over the four thread groups instead of leaving three of
// them idle, and add the partial sums through `red` ([4][256][4] floats).
template <class F>
__device__ __forceinline__ void sb_layer_res(const float* __restrict__ Ws, const float* __restrict__ bias, int K, int J,
                                             const float* __restrict__ in_s, int nrows, float* __restrict__ red, F f4) {
  const int tid = threadIdx.x, cq = tid & 63, rg = tid >> 6;
  const bool ksplit = nrows <= 4;
  for (int jb = 0; jb < J; jb += SB_THREADS) {
    const int jc = min(SB_THREADS, J - jb);
    int col[4];
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      col[q] = min(cq + 64 * q, jc - 1);
      const float bj = (!ksplit || rg == 0) ? bias[jb + col[q]] : 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[q][r] = bj;
    }
    if (ksplit || 4 * rg < nrows) {
      const float* wp = Ws + jb;
      const float* ap = in_s + (ksplit ? 0 : 4 * rg);
      const int kstep = ksplit ? 4 : 1;
      int k = ksplit ? rg : 0;
      for (; k + 3 * kstep < K; k += 4 * kstep) {
        float4 w[4], a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* wr = wp + (long long)(k + u * kstep) * J;
          w[u] = make_float4(wr[col[0]], wr[col[1]], wr[col[2]], wr[col[3]]);
          a[u] = *reinterpret_cast<const float4*>(ap + (k + u * kstep) * SB_LD);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) sb_fma16(acc, a[u], w[u]);
      }
      for (; k < K; k += kstep) {
        const float* wr = wp + (long long)k * J;
        sb_fma16(acc, *reinterpret_cast<const float4*>(ap + k * SB_LD), make_float4(wr[col[0]], wr[col[1]], wr[col[2]], wr[col[3]]));
      }
    }
    if (ksplit) {
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (cq + 64 * q < jc)
          *reinterpret_cast<float4*>(red + ((rg * SB_THREADS + cq + 64 * q) << 2)) =
              make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
      __syncthreads();
      // thread (column tid): sum the four partials of its column, rows 0..3
      if (tid < jc) {
        float4 s = *reinterpret_cast<const float4*>(red + (tid << 2));
#pragma unroll
        for (int g = 1; g < 4; ++g) {
          const float4 t = *reinterpret_cast<const float4*>(red + ((g * SB_THREADS + tid) << 2));
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        f4(jb + tid, 0, s);
#pragma unroll
        for (int r0 = 4; r0 < SB_ROWS; r0 += 4) f4(jb + tid, r0, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    } else {
      const bool row_active = 4 * rg < nrows;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (cq + 64 * q < jc)
          f4(jb + cq + 64 * q, 4 * rg,
             row_active ? make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

__global__ void __launch_bounds__(SB_THREADS) k_sb_forward(const __grid_constant__ SbFwdArgs args) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bars[SB_NST];
  int ni = 0;
  while (ni + 1 < args.n_nets && (int)blockIdx.x >= args.cta_base[ni + 1]) ++ni;
  const rlc_sb_net& n = args.net[ni];
  const int rc = blockIdx.x - args.cta_base[ni];
  const int rows = n.rows ? n.rows : args.B;
  const int tid = threadIdx.x, b0 = rc * SB_ROWS;
  const int inp = n.inp, H1 = n.H1, H2 = n.H2, O = n.O;
  unsigned long long* dbg = args.dbg;
  SB_TRACE(0);
  const bool resident = args.resident[ni] != 0;
  SbRing ring;
  ring.buf = sm;                                                  // ring: [SB_NST][SB_KC][SB_THREADS] | resident: W1, W2, red
  float* x_s = sm + (resident ? inp * H1 + H1 * H2 + 4 * SB_THREADS * 4 : SB_NST * SB_KC * SB_THREADS);   // [inp][16]
  float* h1_s = x_s + inp * SB_LD;                              // [H1][16]
  float* h2_s = h1_s + H1 * SB_LD;                              // [H2][16]
  float* o_s = h2_s + H2 * SB_LD;                               // [16][O]
  float* w3_s = o_s + SB_ROWS * O;                               // [H2][O] + b3[O]: staged early, with both bias vectors --
  float* b2_s = w3_s + H2 * O + O;                               // [H2]      nothing on the critical path waits on HBM
  float* b1_s = b2_s + H2;                                       // [H1]
  ring.bar0 = sb_smem_u32(bars);
  ring.uses = 0;
  if (tid == 0) {
    for (int i = 0; i < SB_NST; ++i) sb_mbar_init(ring.bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (resident) {   // both weight matrices start moving now; layer 1 runs under W2's copy
      const uint32_t bytes1 = (uint32_t)(inp * H1 * 4), bytes2 = (uint32_t)(H1 * H2 * 4);
      sb_mbar_expect_tx(ring.bar0, bytes1);
      sb_bulk_g2s(sb_smem_u32(sm), n.theta, bytes1, ring.bar0);
      sb_mbar_expect_tx(ring.bar0 + 8, bytes2);
      const float* w2g = n.theta + (long long)inp * H1 + H1;
      for (uint32_t off = 0; off < bytes2; off += 32768u)
        sb_bulk_g2s(sb_smem_u32(sm + inp * H1) + off, reinterpret_cast<const char*>(w2g) + off, min(32768u, bytes2 - off),
                    ring.bar0 + 8);
    }
  }
  const float* W1 = n.theta;
  const float* b1 = W1 + (long long)inp * H1;
  const float* W2 = b1 + H1;
  const float* b2 = W2 + (long long)H1 * H2;
  const float* W3 = b2 + H2;
  const float* b3 = W3 + (long long)H2 * O;
  for (int i = tid; i < H2 * O; i += SB_THREADS) {
    const float w = W3[i];
    w3_s[i] = w;
    if (rc == 0 && n.w3_snapshot) n.w3_snapshot[i] = w;
  }
  for (int i = tid; i < O; i += SB_THREADS) w3_s[H2 * O + i] = b3[i];
  for (int i = tid; i < H2; i += SB_THREADS) b2_s[i] = b2[i];
  for (int i = tid; i < H1; i += SB_THREADS) b1_s[i] = b1[i];
  if (rc == 0) {
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
    x_s[k * SB_LD + r] = (b0 + r < rows) ? sb_x(n, b0 + r, k) : 0.f;
  }
  __syncthreads();
  SB_TRACE(1);
  const int nrows = min(SB_ROWS, rows - b0);
  auto put = [&](float* hs, float* hg, int Hn, int j, int r0, float4 v) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    *reinterpret_cast<float4*>(hs + j * SB_LD + r0) = v;
    if (hg) {
      if (r0 < nrows) hg[(long long)(b0 + r0) * Hn + j] = v.x;
      if (r0 + 1 < nrows) hg[(long long)(b0 + r0 + 1) * Hn + j] = v.y;
      if (r0 + 2 < nrows) hg[(long long)(b0 + r0 + 2) * Hn + j] = v.z;
      if (r0 + 3 < nrows) hg[(long long)(b0 + r0 + 3) * Hn + j] = v.w;
    }
  };
  auto put_h1 = [&](int j, int r0, float4 v) { put(h1_s, n.h1, H1, j, r0, v); };
  auto put_h2 = [&](int j, int r0, float4 v) { put(h2_s, n.h2, H2, j, r0, v); };
  if (resident) {
    float* red = sm + inp * H1 + H1 * H2;
    sb_mbar_wait(ring.bar0, 0);
    sb_layer_res(sm, b1_s, inp, H1, x_s, nrows, red, put_h1);
    __syncthreads();
    SB_TRACE(2);
    sb_mbar_wait(ring.bar0 + 8, 0);
    sb_layer_res(sm + inp * H1, b2_s, H1, H2, h1_s, nrows, red, put_h2);
  } else {
    sb_layer(W1, b1_s, inp, H1, x_s, nrows, ring, put_h1);
    __syncthreads();
    SB_TRACE(2);
    sb_layer(W2, b2_s, H1, H2, h1_s, nrows, ring, put_h2, dbg, 8);
  }
  __syncthreads();
  SB_TRACE(3);
  // output layer: O is tiny (1 or 2A); one warp per (row, output) dot product
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = warp; p < nrows * O; p += SB_THREADS / 32) {
    const int r = p / O, o = p % O;
    float s = 0.f;
    for (int j = lane; j < H2; j += 32) s = fmaf(h2_s[j * SB_LD + r], w3_s[j * O + o], s);
    s = warp_sum(s);
    if (lane == 0) {
      s += w3_s[H2 * O + o];
      o_s[r * O + o] = s;
      if (b0 + r < rows) n.out[(long long)(b0 + r) * O + o] = s;
    }
  }
  SB_TRACE(4);
  if (n.policy) {
    __syncthreads();
    if (tid < SB_ROWS && b0 + tid < rows) {
      const int b = b0 + tid, A = O / 2;
      policy_evaluate_row(o_s + tid * O, n.eps ? n.eps + (long long)b * A : nullptr, A, n.action_scale,
                          n.log_std_min, n.log_std_max, n.action ? n.action + (long long)b * A : nullptr,
                          n.logp ? n.logp + b : nullptr, n.mean ? n.mean + (long long)b * A : nullptr,
                          n.mu_raw ? n.mu_raw + (long long)b * A : nullptr,
                          n.log_std ? n.log_std + (long long)b * A : nullptr, n.z ? n.z + (long long)b * A : nullptr);
    }
  }
}

// Adam (+ Polyak) on one parameter given its gradient; p/m/v(/t) are the pre-update values.
__device__ __forceinline__ void sb_adam_vals(float p, float m, float v, float t, float g, float lr_eff, float isb2,
                                             float b1, float b2, float eps, float tau, float& pn, float& mn, float& vn,
                                             float& tn) {
  mn = b1 * m + (1.f - b1) * g;
  vn = b2 * v + (1.f - b2) * g * g;
  pn = p - lr_eff * mn / (sqrtf(vn) * isb2 + eps);
  tn = t + tau * (pn - t);
}
__device__ __forceinline__ void sb_adam(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                                        float* __restrict__ target, long long i, float g, float lr_eff, float isb2,
                                        float b1, float b2, float eps, float tau) {
  float pn, mn, vn, tn;
  sb_adam_vals(theta[i], m[i], v[i], target ? target[i] : 0.f, g, lr_eff, isb2, b1, b2, eps, tau, pn, mn, vn, tn);
  m[i] = mn;
  v[i] = vn;
  theta[i] = pn;
  if (target) target[i] = tn;
}

__global__ void __launch_bounds__(SB_THREADS) k_sb_update(const __grid_constant__ SbUpdArgs args) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bar, bar_h2;
  int ni = 0;
  while (ni + 1 < args.n_nets && (int)blockIdx.x >= args.cta_base[ni + 1]) ++ni;
  const rlc_sb_train& n = args.net[ni];
  const int c = blockIdx.x - args.cta_base[ni], C = args.cta_base[ni + 1] - args.cta_base[ni];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, B = args.B;
  const int inp = n.inp, H1 = n.H1, H2 = n.H2, O = n.O;
  unsigned long long* dbg = args.dbg;
  SB_TRACE(64);
  // row stride of dz2: a multiple of 4 floats (float4 reads along j) with an ODD number of quads, so that float4 reads
  // by lanes walking b (dh1 phase) fall in distinct bank groups; pad columns stay 0
  const int H2p = 4 * (((H2 + 3) >> 2) | 1);
  const int i0 = c * SB_NI, nI = min(SB_NI, H1 - i0);
  const int Jc = (H2 + C - 1) / C, j0 = c * Jc, nJ = max(0, min(Jc, H2 - j0));
  // shared-memory carve-up (float offsets; the first four blocks are multiples of 4 floats)
  float* pw = sm;                               // [4][SB_NI][H2]  theta | m | v | target rows of W2 owned here
  float* h2f_s = pw + 4 * SB_NI * H2;           // [B][H2]        the whole second hidden layer (bulk copy)
  float* h1_s = h2f_s + ((B * H2 + 3) & ~3);    // [B][SB_NI]
  float* dz1_s = h1_s + B * SB_NI;              // [B][SB_NI]
  float* red_s = dz1_s + B * SB_NI;             // [8]
  float* dz2_s = red_s + 8;                     // [B][H2p]     (16-byte aligned rows)
  float* dout_s = dz2_s + B * H2p;              // [B][O]
  float* x_s = dout_s + B * O;                  // [B][inp]
  float* h2_s = x_s + B * inp;                  // [B][Jc]      this CTA's share of the output layer
  const long long oW1 = 0, ob1 = (long long)inp * H1, oW2 = ob1 + H1, ob2 = oW2 + (long long)H1 * H2, oW3 = ob2 + H2,
                  ob3 = oW3 + (long long)H2 * O;
  const float lr_eff = reinterpret_cast<const float*>(n.adam_state)[1];
  const float isb2 = reinterpret_cast<const float*>(n.adam_state)[2];
  // ---- the CTA's W2 rows (parameters + both moments + the Polyak target) start moving first: one bulk copy each,
  //      in flight under the whole gradient computation; Adam later reads them from shared memory
  const long long w2off = oW2 + (long long)i0 * H2;
  const uint32_t slice_bytes = (uint32_t)(nI * H2 * 4);
  const bool bulk = ((((uintptr_t)(n.theta + w2off)) | ((uintptr_t)(n.m + w2off)) | ((uintptr_t)(n.v + w2off)) |
                      (n.target ? (uintptr_t)(n.target + w2off) : 0)) & 15) == 0 && (slice_bytes & 15) == 0 &&
                    ((SB_NI * H2) & 3) == 0;
  const uint32_t bar_a = sb_smem_u32(&bar), bar_b = sb_smem_u32(&bar_h2);
  const uint32_t h2_bytes = (uint32_t)(B * H2 * 4);
  const bool bulk_h2 = (((uintptr_t)n.h2) & 15) == 0 && (h2_bytes & 15) == 0;
  if (tid == 0) {
    sb_mbar_init(bar_a, 1);
    sb_mbar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bulk_h2) {   // needed first (dz2), so it is issued first
      sb_mbar_expect_tx(bar_b, h2_bytes);
      for (uint32_t off = 0; off < h2_bytes; off += 32768u)
        sb_bulk_g2s(sb_smem_u32(h2f_s) + off, reinterpret_cast<const char*>(n.h2) + off, min(32768u, h2_bytes - off), bar_b);
    }
    if (bulk) {
      sb_mbar_expect_tx(bar_a, slice_bytes * (n.target ? 4 : 3));
      sb_bulk_g2s(sb_smem_u32(pw), n.theta + w2off, slice_bytes, bar_a);
      sb_bulk_g2s(sb_smem_u32(pw + SB_NI * H2), n.m + w2off, slice_bytes, bar_a);
      sb_bulk_g2s(sb_smem_u32(pw + 2 * SB_NI * H2), n.v + w2off, slice_bytes, bar_a);
      if (n.target) sb_bulk_g2s(sb_smem_u32(pw + 3 * SB_NI * H2), n.target + w2off, slice_bytes, bar_a);
    }
  }
  if (!bulk) {
    for (int p = tid; p < nI * H2; p += SB_THREADS) {
      pw[p] = n.theta[w2off + p];
      pw[SB_NI * H2 + p] = n.m[w2off + p];
      pw[2 * SB_NI * H2 + p] = n.v[w2off + p];
      if (n.target) pw[3 * SB_NI * H2 + p] = n.target[w2off + p];
    }
  }
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
        if (o == 0 && n.loss_b) loss_part += n.loss_b[b] / (float)B;
      } else {
        const float raw = n.out[b * O + o];
        d = (raw >= n.log_std_min && raw <= n.log_std_max) ? n.dlog_std[b * A + (o - A)] : 0.f;
      }
    } else {
      d = n.dout[p];
    }
    dout_s[p] = d;
  }
  if (H2p != H2)
    for (int p = tid; p < B * (H2p - H2); p += SB_THREADS) dz2_s[(p / (H2p - H2)) * H2p + H2 + p % (H2p - H2)] = 0.f;
  if (c == 0 && n.loss_out) {                  // deterministic: warp sums, then a serial sum over the 8 warps
    loss_part = warp_sum(loss_part);
    if (lane == 0) red_s[warp] = loss_part;
  }
  // everything else this CTA reads from HBM, in one batch of independent loads
  for (int p = tid; p < B * SB_NI; p += SB_THREADS) {
    const int b = p / SB_NI, i = p % SB_NI;
    h1_s[p] = i < nI ? n.h1[(long long)b * H1 + i0 + i] : 0.f;
  }
  for (int p = tid; p < B * inp; p += SB_THREADS) x_s[p] = sb_xt(n, p / inp, p % inp);
  for (int p = tid; p < B * nJ; p += SB_THREADS) h2_s[(p / nJ) * Jc + p % nJ] = n.h2[(long long)(p / nJ) * H2 + j0 + p % nJ];
  // ---- dz2[b][j] = (dout[b,:] . W3[j,:]) * relu'(h2[b,j])
  __syncthreads();
  SB_TRACE(65);
  if (c == 0 && n.loss_out && tid == 0) {
    float s = 0.f;
    for (int w = 0; w < SB_THREADS / 32; ++w) s += red_s[w];
    n.loss_out[0] = s;
  }
  // thread = column j: its W3 row in registers, coalesced h2 loads over b, no integer division
  if (bulk_h2) sb_mbar_wait(bar_b, 0);
  const float* h2src = bulk_h2 ? h2f_s : n.h2;   // (generic pointer: shared or global)
  for (int j = tid; j < H2; j += SB_THREADS) {
    float w3r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) w3r[o] = o < O ? __ldg(n.w3_snapshot + j * O + o) : 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) {
      const float hv = h2src[(long long)b * H2 + j];
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (o < O) s = fmaf(dout_s[b * O + o], w3r[o], s);
      for (int o = 4; o < O; ++o) s = fmaf(dout_s[b * O + o], __ldg(n.w3_snapshot + j * O + o), s);
      dz2_s[b * H2p + j] = hv > 0.f ? s : 0.f;
    }
  }
  SB_TRACE(66);
  if (bulk) sb_mbar_wait(bar_a, 0);
  __syncthreads();
  SB_TRACE(67);
  // ---- dh1[b][i] = dz2[b,:] . W2[i,:] ; dz1 = dh1 * relu'(h1): a warp takes unit i (its W2 row is a broadcast read),
  //      lanes take b, four partial sums break the FMA dependency chain.  (A variant with the W2 row in registers,
  //      lanes over columns and shuffle reductions measured 3x slower.)
  for (int i = warp; i < SB_NI; i += SB_THREADS / 32) {
    for (int b = lane; b < B; b += 32) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (i < nI) {
        const float* dz = dz2_s + b * H2p;
        const float* w = pw + i * H2;
        for (int j = 0; j < H2; j += 4) {         // pad columns of dz2 are 0
          const float4 d = *reinterpret_cast<const float4*>(dz + j);
          float4 ww;
          if (j + 3 < H2 && (H2 & 3) == 0) ww = *reinterpret_cast<const float4*>(w + j);
          else ww = make_float4(w[j], j + 1 < H2 ? w[j + 1] : 0.f, j + 2 < H2 ? w[j + 2] : 0.f, j + 3 < H2 ? w[j + 3] : 0.f);
          s0 = fmaf(d.x, ww.x, s0);
          s1 = fmaf(d.y, ww.y, s1);
          s2 = fmaf(d.z, ww.z, s2);
          s3 = fmaf(d.w, ww.w, s3);
        }
        s0 = (s0 + s1) + (s2 + s3);
        if (!(h1_s[b * SB_NI + i] > 0.f)) s0 = 0.f;
      }
      dz1_s[b * SB_NI + i] = s0;
    }
  }
  SB_TRACE(68);
  // ---- dW2[i][j] = sum_b h1[b,i] dz2[b,j], straight into Adam.  Register tiles of SB_NI/4 units x 4 columns, so that
  //      all four thread groups work: the Adam arithmetic (an IEEE sqrt and division per parameter) is a long
  //      dependent chain per thread -- with 4-unit tiles only two groups were busy and it took 4.7 us of the kernel.
  constexpr int UPT = SB_NI / 4;                // units per thread
  for (int jb = 0; jb < H2; jb += SB_THREADS) {
    const int jq = tid & 63, ig = tid >> 6, j = jb + 4 * jq;
    if (j >= H2) continue;
    float acc[UPT][4];   // [unit][column]
#pragma unroll
    for (int u = 0; u < UPT; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][q] = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      const float4 d = *reinterpret_cast<const float4*>(dz2_s + b * H2p + j);     // H2p % 4 == 0; pad columns are 0
#pragma unroll
      for (int u = 0; u < UPT; ++u) {
        const float hv = h1_s[b * SB_NI + UPT * ig + u];                          // warp-wide broadcast
        acc[u][0] = fmaf(hv, d.x, acc[u][0]);
        acc[u][1] = fmaf(hv, d.y, acc[u][1]);
        acc[u][2] = fmaf(hv, d.z, acc[u][2]);
        acc[u][3] = fmaf(hv, d.w, acc[u][3]);
      }
    }
    SB_TRACE(72);
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
      const int i = UPT * ig + u;
      if (i >= nI) continue;
      const int e = i * H2 + j;
      if (bulk && (H2 & 3) == 0) {   // aligned rows: 128-bit shared loads and coalesced 128-bit stores
        const float4 p4 = *reinterpret_cast<const float4*>(pw + e), m4 = *reinterpret_cast<const float4*>(pw + SB_NI * H2 + e),
                     v4 = *reinterpret_cast<const float4*>(pw + 2 * SB_NI * H2 + e);
        const float4 t4 = n.target ? *reinterpret_cast<const float4*>(pw + 3 * SB_NI * H2 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pn, mn, vn, tn;
        sb_adam_vals(p4.x, m4.x, v4.x, t4.x, acc[u][0], lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau, pn.x, mn.x, vn.x, tn.x);
        sb_adam_vals(p4.y, m4.y, v4.y, t4.y, acc[u][1], lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau, pn.y, mn.y, vn.y, tn.y);
        sb_adam_vals(p4.z, m4.z, v4.z, t4.z, acc[u][2], lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau, pn.z, mn.z, vn.z, tn.z);
        sb_adam_vals(p4.w, m4.w, v4.w, t4.w, acc[u][3], lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau, pn.w, mn.w, vn.w, tn.w);
        *reinterpret_cast<float4*>(n.theta + w2off + e) = pn;
        *reinterpret_cast<float4*>(n.m + w2off + e) = mn;
        *reinterpret_cast<float4*>(n.v + w2off + e) = vn;
        if (n.target) *reinterpret_cast<float4*>(n.target + w2off + e) = tn;
        continue;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (j + q >= H2) continue;
        float pn, mn, vn, tn;
        sb_adam_vals(pw[e + q], pw[SB_NI * H2 + e + q], pw[2 * SB_NI * H2 + e + q], n.target ? pw[3 * SB_NI * H2 + e + q] : 0.f,
                     acc[u][q], lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau, pn, mn, vn, tn);
        n.theta[w2off + e + q] = pn;
        n.m[w2off + e + q] = mn;
        n.v[w2off + e + q] = vn;
        if (n.target) n.target[w2off + e + q] = tn;
      }
    }
  }
  SB_TRACE(69);
  __syncthreads();
  SB_TRACE(70);
  // ---- everything small, one flattened pass (one round trip to HBM for its Adam state):
  //      dW1[k][i] = sum_b x[b,k] dz1[b,i], db1[i] = sum_b dz1[b,i]            ((inp + 1) x nI outputs)
  //      dW3[j][o] = sum_b h2[b,j] dout[b,o], db2[j] = sum_b dz2[b,j]          (nJ x (O + 1) outputs, this CTA's share)
  //      db3[o] = sum_b dout[b,o]                                             (last CTA)
  const int n1 = (inp + 1) * SB_NI, n2 = nJ * (O + 1), n3 = (c == C - 1) ? O : 0;
  for (int p = tid; p < n1 + n2 + n3; p += SB_THREADS) {
    float s = 0.f;
    long long idx;
    if (p < n1) {
      const int k = p / SB_NI, i = p % SB_NI;
      if (i >= nI) continue;
      if (k < inp) {
        for (int b = 0; b < B; ++b) s = fmaf(x_s[b * inp + k], dz1_s[b * SB_NI + i], s);
        idx = oW1 + (long long)k * H1 + i0 + i;
      } else {
        for (int b = 0; b < B; ++b) s += dz1_s[b * SB_NI + i];
        idx = ob1 + i0 + i;
      }
    } else if (p < n1 + n2) {
      const int q = p - n1, jj = q / (O + 1), o = q % (O + 1);
      if (o < O) {
        for (int b = 0; b < B; ++b) s = fmaf(h2_s[b * Jc + jj], dout_s[b * O + o], s);
        idx = oW3 + (long long)(j0 + jj) * O + o;
      } else {
        for (int b = 0; b < B; ++b) s += dz2_s[b * H2p + j0 + jj];
        idx = ob2 + j0 + jj;
      }
    } else {
      const int o = p - n1 - n2;
      for (int b = 0; b < B; ++b) s += dout_s[b * O + o];
      idx = ob3 + o;
    }
    sb_adam(n.theta, n.m, n.v, n.target, idx, s, lr_eff, isb2, n.beta1, n.beta2, n.eps, n.tau);
  }
  SB_TRACE(71);
}

// RLC_SB_DEBUG = hex device address of a >= 128-slot uint64 buffer (timing scripts set it before the launch they trace;
// re-read on every call because they set it after the library is loaded -- one getenv, no allocation)
static unsigned long long* sb_debug_buffer() {
  const char* e = getenv("RLC_SB_DEBUG");
  return e ? (unsigned long long*)strtoull(e, nullptr, 16) : nullptr;
}

static bool sb_dims_ok(int inp, int H1, int H2, int O, int n0, int n1) {
  return inp >= 1 && inp <= 256 && H1 >= 1 && H1 <= 512 && H2 >= 1 && H2 <= 512 && O >= 1 && O <= 32 &&
         n0 >= 0 && n1 >= 0 && n0 + n1 == inp;
}

extern "C" int rlc_sb_forward(rlc_handle* h, const rlc_sb_net* nets, int n_nets, int B, void* stream) {
  RLC_REQUIRE(h && nets && n_nets >= 1 && n_nets <= RLC_SB_MAX_NETS && B >= 1 && B <= RLC_SB_MAX_B);
  SbFwdArgs args;
  size_t smem = 0;
  int base = 0;
  for (int i = 0; i < n_nets; ++i) {
    const rlc_sb_net& n = nets[i];
    RLC_REQUIRE(n.theta && n.out && sb_dims_ok(n.inp, n.H1, n.H2, n.O, n.n0, n.n1));
    RLC_REQUIRE((n.n0 == 0 || n.x0) && (n.n1 == 0 || n.x1));
    RLC_REQUIRE(n.rows >= 0 && n.rows <= 65536 && n.x0_div >= 0 && n.x1_mod >= 0);
    RLC_REQUIRE(!n.policy || (n.O % 2 == 0 && n.O / 2 <= 16 && n.log_std_min <= n.log_std_max));
    RLC_REQUIRE(!n.adam_state || n.adam_variant == RLC_ADAM_TORCH || n.adam_variant == RLC_ADAM_TF);
    args.net[i] = n;
    args.cta_base[i] = base;
    base += ((n.rows ? n.rows : B) + SB_ROWS - 1) / SB_ROWS;
    const size_t act = (size_t)(n.inp + n.H1 + n.H2) * SB_LD + (size_t)SB_ROWS * n.O + (size_t)n.H2 * n.O + n.O + n.H1 + n.H2;
    const size_t s_ring = sizeof(float) * ((size_t)SB_NST * SB_KC * SB_THREADS + act);
    const size_t s_res = sizeof(float) * ((size_t)n.inp * n.H1 + (size_t)n.H1 * n.H2 + 4 * SB_THREADS * 4 + act);
    // resident weights pay off for the tiny latency-critical launches (sample_action, an evaluation step: one or two
    // CTAs); at ~205 KB they leave room for one CTA per SM, so the big launch of a training step keeps the ring (two
    // CTAs per SM: with 8 runs sharing the GPU its 132 CTAs per run are throughput, not latency)
    int total_ctas = 0;
    for (int q = 0; q < n_nets; ++q) total_ctas += ((nets[q].rows ? nets[q].rows : B) + SB_ROWS - 1) / SB_ROWS;
    const bool res = total_ctas <= 8 && (((uintptr_t)n.theta) & 15) == 0 && ((n.inp * n.H1) & 3) == 0 && (n.H1 & 3) == 0 &&
                     (n.H2 & 3) == 0 && s_res + 1024 <= h->smem_optin && getenv("RLC_SB_NO_RESIDENT") == nullptr;
    args.resident[i] = res ? 1 : 0;
    const size_t s = res ? s_res : s_ring;
    smem = s > smem ? s : smem;
  }
  args.cta_base[n_nets] = base;
  args.n_nets = n_nets;
  args.B = B;
  args.dbg = sb_debug_buffer();
  if (smem > 48 * 1024) {   // a per-device function attribute: set on every launch that needs it (host-side, ~1 us)
    RLC_REQUIRE(smem <= h->smem_optin);
    RLC_CUDA(cudaFuncSetAttribute(k_sb_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  k_sb_forward<<<base, SB_THREADS, smem, (cudaStream_t)stream>>>(args);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_sb_update(rlc_handle* h, const rlc_sb_train* nets, int n_nets, int B, int B_total, void* stream) {
  RLC_REQUIRE(h && nets && n_nets >= 1 && n_nets <= RLC_SB_MAX_NETS && B >= 1 && B <= RLC_SB_MAX_B && B_total >= B);
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
    const int Cn = (n.H1 + SB_NI - 1) / SB_NI, Jc = (n.H2 + Cn - 1) / Cn;
    const size_t s = sizeof(float) * ((size_t)4 * SB_NI * n.H2 + (size_t)B * n.H2 + 4 + 2 * (size_t)B * SB_NI + 8 + (size_t)B * n.O +
                                      (size_t)B * (n.H2 + 8) + (size_t)B * n.inp + (size_t)B * Jc);
    smem = s > smem ? s : smem;
    rlc_invalidate_pack(h, n.theta);
    if (n.target) rlc_invalidate_pack(h, n.target);
  }
  args.cta_base[n_nets] = base;
  args.n_nets = n_nets;
  args.B = B;
  args.inv_btotal = 1.f / (float)B_total;
  args.dbg = sb_debug_buffer();
  if (smem > 48 * 1024) {
    RLC_REQUIRE(smem <= h->smem_optin);
    RLC_CUDA(cudaFuncSetAttribute(k_sb_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  k_sb_update<<<base, SB_THREADS, smem, (cudaStream_t)stream>>>(args);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
