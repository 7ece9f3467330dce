// K1-grid: fused T-in critic evaluation for a SHARED action grid a[N,A] (ForwardKL / ReverseKL
// quadrature, forwardkl_network.py:104-105,160-164).  Included by critic_umma.cu.
//
// With a shared grid, layer 1 separates:  W1 [s_b ; a_n] + b1 = PS[b] + PA[n]  with
//   PS[b] = b1 + W1s s_b   (one row per state)      PA[n] = W1a a_n   (one row per grid action)
// so the B*N x (S+A) x H1 contraction collapses to B + N rows, computed once per launch by a small
// fp32 pre-pass (k_grid_parts) and stored as fp16.  The main kernel then needs no layer-1 MMA, no
// X tile and -- the point -- no tcgen05.ld of a layer-1 accumulator: epilogue 1 forms
//   h1 = relu(PS[b] + PA[n])          (one HFMA2.RELU per two activations, fp16)
// from shared memory and writes it straight into tensor memory as layer 2's A operand, running
// AHEAD of the tensor pipe (it depends on nothing the MMA produces).  Per 128-row tile the warps'
// TMEM traffic drops from 463 KB (205 KB ld + 102 KB st + 156 KB ld) to 258 KB and the tensor pipe
// only runs layer 2.
//
// CTA tile = 4 states x 32 grid actions (TMEM lane r <-> state b0 + r/32, action n0 + r%32), so a
// warp (one TMEM lane quarter) owns one state and PA tiles are reused by four states.  A CTA pair
// (cta_group::2, M=256) covers two consecutive CTA tiles.  Layer 2 and the output head are the
// TS variant's (folded head, sign-partitioned columns; k_pack_head).
//
// Layer 2 runs in TWO PASSES per tile: all K chunks into accumulator half A (columns [0,NA)), then
// all K chunks into half B.  The whole tile's activations (H1P fp16 = H1P/2 TMEM columns) stay
// resident next to the accumulator (H2P + H1P/2 <= 512 columns: 304 + 208 at 400-300), so
// epilogue 2 drains half A while the tensor pipe is busy with pass B, and half B while it runs
// pass A of the next tile: the accumulator hand-over costs the tensor pipe nothing.
//
// Warps: 0 MMA issuer (leader) | 1 loader (cp.async.bulk of PA chunk tiles and PS rows into
// shared-memory rings, mbarrier complete_tx) | 4 x GR_E1G epilogue 1 | 4 x GR_E2G epilogue 2 (each
// group of 4 warps covers the four TMEM lane quarters and takes a share of the columns).
#pragma once

#define GR_E1G 1
#define GR_E2G 2
#define GR_E1C 1
#define GR_VR (GR_E2G >= 3 ? 56 : 96)   // accumulator columns an epilogue-2 thread holds in registers per round
#define GR_W2_0 (4 + 4 * GR_E1G * GR_E1C)                    // first epilogue-2 warp
#define GR_THREADS (32 * (4 + 4 * GR_E1G * GR_E1C + 4 * GR_E2G))
#define GR_PA_STAGES 4
#define GR_MAX_NCH 8
#define GR_PITCH_PAD 8  // halfs of padding per PA row: pitch (CH+8)*2 B keeps 16-byte LDS conflict-free

struct GridParams {
  float* q;
  int B, N;
  int H1P, H2P, NA, NB;
  int CH, nch, nbuf;          // layer-1 feature chunking: chunk c = features [c*CH, min(H1P,(c+1)*CH))
  int NT;                     // 32-action blocks per state = ceil(N/32)
  long long num_cta_tiles;    // ceil(B/4) * NT
  int num_pair_tiles;
  const __half* ps;           // [B][H1P]                fp16/bf16 bits
  const __half* pa;           // [nch][NT*32][CH+8]      fp16/bf16 bits, zero rows past N
  const unsigned char* blob[2];
  int off_w2, off_w1, off_c0;
  int sm_w2, sm_pa, sm_ps, sm_qp, sm_bar, pa_stage_bytes, ps_stage_bytes;
  int* err;
  long long* prof;            // RLC_UMMA_PROF=1: 32 x int64 per pair
  int micro;                  // debug: 1 = tensor-pipe microbenchmark (MMA issuer free-runs, epilogues idle, output garbage)
};

enum {
  GB_PA_FULL = 0,     // [4]  count 1 + tx   (local)    loader -> ep1
  GB_PA_EMPTY = 4,    // [4]  count 4        (local)    ep1 warps -> loader
  GB_PS_FULL = 8,     // [2]  count 1 + tx   (local)    loader -> ep1
  GB_PS_EMPTY = 10,   // [2]  count 4        (local)    ep1 warps -> loader
  GB_H1_FULL = 12,    // [8]  count 8        (leader)   ep1 (4 warps x 2 CTAs) -> MMA, one per K chunk
  GB_H1_EMPTY = 20,   // [8]  count 1        (both)     MMA commit (last pass over the chunk) -> ep1
  GB_L2_FULL = 28,    // [2]  count 1        (both)     MMA commit -> ep2 (half A, half B)
  GB_L2_EMPTY = 30,   // [2]  count 16       (leader)   ep2 (8 warps x 2 CTAs) -> MMA
  GB_COUNT = 32
};

namespace um {
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
template <int PREC>
__device__ __forceinline__ uint32_t add_relu2(uint32_t a, uint32_t b) {
  uint32_t r;   // relu(a + b) on two packed halves, one rounding (fma with 1.0)
  if (PREC == RLC_PREC_BF16)
    asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3F803F80u), "r"(b));
  else
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3C003C00u), "r"(b));
  return r;
}
}  // namespace um

// Pre-pass (fp32 FMA, then one rounding to the operand type):
//   PS[b][j] = b1[j] + sum_k clip(s[b][k]) W1[k][j]   (j < H1);  PS[b][H1] = 1 (the bias carrier of
//   layer 2, see k_pack_head);  PA[c][n][jj] = sum_k a[n][k] W1[S+k][c*CH+jj].
template <int PREC>
__global__ void __launch_bounds__(128)
k_grid_parts(const float* __restrict__ theta, const float* __restrict__ s, const float* __restrict__ a,
             const float* __restrict__ smin, const float* __restrict__ smax, int B, int N, int S, int A,
             int H1, int H2, int H1P, int CH, int nch, int NT, unsigned short* __restrict__ PS,
             unsigned short* __restrict__ PA) {
  // Programmatic dependent launch: let the main kernel start its prologue (weights -> smem, barriers,
  // TMEM) while this pre-pass runs; its loader waits (griddepcontrol.wait) before touching PS/PA.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // grid: x = row (states first, then padded grid actions), y = feature block of 128; no 64-bit divisions
  const ThetaView t = theta_view(RLC_TIN, S, A, H1, H2);
  const float* W1 = theta + t.oW1;   // [S+A][H1]
  const float* b1 = theta + t.ob1;
  const int j = blockIdx.y * 128 + threadIdx.x;
  const int row = blockIdx.x;
  if (j >= H1P) return;
  if (row < B) {
    float acc = 0.f;
    if (j < H1) {
      acc = b1[j];
      for (int k = 0; k < S; ++k) {
        float x = __ldg(s + (size_t)row * S + k);
        if (smin) x = fminf(fmaxf(x, __ldg(smin + k)), __ldg(smax + k));
        acc = fmaf(x, W1[(size_t)k * H1 + j], acc);
      }
    } else if (j == H1) {
      acc = 1.f;
    }
    PS[(size_t)row * H1P + j] = to_h<PREC>(acc);
  } else {
    const int n = row - B;                                   // 0 .. NT*32-1 (rows past N are zero)
    const int pitch = CH + GR_PITCH_PAD;
    const int c = j / CH, jj = j - c * CH;
    float acc = 0.f;
    if (j < H1 && n < N) {
      for (int k = 0; k < A; ++k) acc = fmaf(__ldg(a + (size_t)n * A + k), W1[(size_t)(S + k) * H1 + j], acc);
    }
    PA[((size_t)c * NT * 32 + n) * pitch + jj] = to_h<PREC>(acc);   // the 8 padding halfs per row are never read
  }
}

// The same pre-pass with 8 rows per CTA: a W1 element is loaded once and feeds 8 rows, the rows' inputs sit in shared
// memory, and the grid shrinks from (B + NT*32) x 4 tiny CTAs (20 480 at cfg4: ~9 waves of pure load latency, 26 us
// under ncu) to an eighth of that.  Measured effect on K1 + pre-pass: within run-to-run noise (0.677-0.689 ms) -- the main
// kernel's prologue (programmatic dependent launch) already hides most of the pre-pass; kept because it is less work.
#define GR_PRE_ROWS 8
template <int PREC>
__global__ void __launch_bounds__(128)
k_grid_parts8(const float* __restrict__ theta, const float* __restrict__ s, const float* __restrict__ a,
              const float* __restrict__ smin, const float* __restrict__ smax, int B, int N, int S, int A,
              int H1, int H2, int H1P, int CH, int nch, int NT, int state_groups, unsigned short* __restrict__ PS,
              unsigned short* __restrict__ PA, int* __restrict__ err) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ float xs[];                               // [GR_PRE_ROWS][K]  (K = S for states, A for actions)
  const ThetaView t = theta_view(RLC_TIN, S, A, H1, H2);
  const float* W1 = theta + t.oW1;   // [S+A][H1]
  const float* b1 = theta + t.ob1;
  const int j = blockIdx.y * 128 + threadIdx.x;
  const bool is_state = (int)blockIdx.x < state_groups;
  const int r0 = is_state ? blockIdx.x * GR_PRE_ROWS : (blockIdx.x - state_groups) * GR_PRE_ROWS;
  const int K = is_state ? S : A, rows = is_state ? B : N;
  for (int i = threadIdx.x; i < GR_PRE_ROWS * K; i += 128) {
    const int r = i / K, k = i - r * K;
    float x = 0.f;
    if (r0 + r < rows) {
      x = __ldg((is_state ? s : a) + (size_t)(r0 + r) * K + k);
      if (is_state && smin) x = fminf(fmaxf(x, __ldg(smin + k)), __ldg(smax + k));
    }
    xs[i] = x;
  }
  __syncthreads();
  if (j >= H1P) return;
  float acc[GR_PRE_ROWS];
  const float init = (is_state && j < H1) ? b1[j] : 0.f;
#pragma unroll
  for (int r = 0; r < GR_PRE_ROWS; ++r) acc[r] = init;
  if (j < H1) {
    const float* Wk = W1 + (size_t)(is_state ? 0 : S) * H1 + j;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(Wk + (size_t)k * H1);
#pragma unroll
      for (int r = 0; r < GR_PRE_ROWS; ++r) acc[r] = fmaf(xs[r * K + k], w, acc[r]);
    }
  }
  // fp16 operands: a layer-1 pre-activation beyond half of fp16's range would saturate silently (the sum PS + PA is formed
  // in fp16) -- raise the handle's error flag (code 91) so that the caller switches to precision "bf16" / "fp32"
  if (PREC == RLC_PREC_FP16 && j < H1) {
    bool bad = false;
#pragma unroll
    for (int r = 0; r < GR_PRE_ROWS; ++r) bad = bad || ((r0 + r < (is_state ? B : N)) && !(fabsf(acc[r]) <= 32000.f));
    if (bad) atomicCAS(err, 0, 91);
  }
  if (is_state) {
#pragma unroll
    for (int r = 0; r < GR_PRE_ROWS; ++r)
      if (r0 + r < B) PS[(size_t)(r0 + r) * H1P + j] = to_h<PREC>(j == H1 ? 1.f : acc[r]);
  } else {
    const int pitch = CH + GR_PITCH_PAD;
    const int c = j / CH, jj = j - c * CH;
#pragma unroll
    for (int r = 0; r < GR_PRE_ROWS; ++r) {
      const int n = r0 + r;                                   // 0 .. NT*32-1 (rows past N are zero)
      if (n < NT * 32) PA[((size_t)c * NT * 32 + n) * pitch + jj] = to_h<PREC>((j < H1 && n < N) ? acc[r] : 0.f);
    }
  }
}

template <int PREC, bool PROF>
__global__ void __launch_bounds__(GR_THREADS, 1) k_critic_umma_grid(const GridParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = um::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = um::cta_rank();
  const uint32_t pair = um::cluster_id_x();
  const uint32_t npairs = um::num_clusters_x();

  const uint32_t sW2 = base + P.sm_w2, sPA = base + P.sm_pa, sPS = base + P.sm_ps, sBar = base + P.sm_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + P.sm_bar + GB_COUNT * 8);
  float* qpart = reinterpret_cast<float*>(base_ptr + P.sm_qp);   // [2][128] partial sums
  auto bar = [&](int i) -> uint32_t { return sBar + 8u * (uint32_t)i; };

  // ---- prologue: resident W2' -> smem, barriers, TMEM ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.blob[rank]);
    uint4* dW2 = reinterpret_cast<uint4*>(base_ptr + P.sm_w2);
    const int n2 = (P.off_w1 - P.off_w2) >> 4;
    for (int i = tid; i < n2; i += GR_THREADS) dW2[i] = __ldg(src + (P.off_w2 >> 4) + i);
  }
  if (tid == 0) {
    for (int i = 0; i < GR_PA_STAGES; ++i) {
      um::mbar_init(bar(GB_PA_FULL + i), 1);
      um::mbar_init(bar(GB_PA_EMPTY + i), 4 * GR_E1G);
    }
    for (int i = 0; i < 2; ++i) {
      um::mbar_init(bar(GB_PS_FULL + i), 1);
      um::mbar_init(bar(GB_PS_EMPTY + i), 4 * GR_E1G * GR_E1C);
      um::mbar_init(bar(GB_L2_FULL + i), 1);
      um::mbar_init(bar(GB_L2_EMPTY + i), 8 * GR_E2G);
    }
    for (int i = 0; i < GR_MAX_NCH; ++i) {
      um::mbar_init(bar(GB_H1_FULL + i), 8 * GR_E1G);
      um::mbar_init(bar(GB_H1_EMPTY + i), 1);
    }
    um::fence_mbar_init();
  }
  um::fence_proxy_async();
  if (warp == 0) um::tmem_alloc2(um::smem_u32(tmem_slot), 512);
  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  um::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long pa_ = 0, pb_ = 0, pc_ = 0, pd_ = 0, pe_ = 0;   // PROF accumulators (meaning per role)
  const bool prof = PROF && P.prof != nullptr;
  const long long t_begin = prof ? clock64() : 0;
#define GPROF_T() ((PROF && prof) ? clock64() : 0)
  // event trace (pair 0, leader CTA, first 12 tiles): role r appends (code, clock - t_begin) pairs at prof[4096 + r*2048 ..]
  int tr_n = 0;
  const bool tracing = PROF && prof && pair == 0 && rank == 0;
#define GTRACE(role, code)                                                        \
  do {                                                                            \
    if (PROF && tracing && tr_n < 1000) {                                         \
      long long* tb = P.prof + 4096 + (role) * 2048;                              \
      tb[2 * tr_n] = (code);                                                      \
      tb[2 * tr_n + 1] = clock64() - t_begin;                                     \
      ++tr_n;                                                                     \
    }                                                                             \
  } while (0)

  const int ntiles = (P.num_pair_tiles > (int)pair)
                         ? (P.num_pair_tiles - (int)pair + (int)npairs - 1) / (int)npairs
                         : 0;
  const int nch = P.nch, CH = P.CH, last_w = P.H1P - (P.nch - 1) * P.CH;
  const uint32_t H1COL = (uint32_t)P.H2P;            // TMEM column of activation buffer 0
  const uint32_t bufcols = (uint32_t)(CH >> 1);      // two halves per 32-bit cell
  // CTA tile of pair-tile t: ct = 2*t + rank -> state group ct / NT, action block ct % NT
  auto tile_coords = [&](int tl, int& b0, int& n0) {
    long long ct = 2ll * ((long long)pair + (long long)tl * npairs) + rank;
    if (ct >= P.num_cta_tiles) ct = P.num_cta_tiles - 1;   // odd tail: recompute the last tile, never stored twice
    b0 = (int)(ct / P.NT) * 4;
    n0 = (int)(ct % P.NT) * 32;
  };

  if (warp == 0) {
    // =================================== MMA issuer (leader CTA) ===================================
    if (rank == 0) {
      const bool issuer = um::elect_one();
      const uint32_t fmt = (PREC == RLC_PREC_BF16) ? 1u : 0u;
      const uint32_t lbo_w2 = (uint32_t)(P.H2P / 2) * 16u;
      const uint32_t w2_lo = um::desc_lo(sW2, lbo_w2);
      const uint32_t w2_kstep = (2u * lbo_w2) >> 4;
      const uint32_t w2_half_off = (uint32_t)(P.NA / 2);
      const uint32_t idA = um::make_idesc(fmt, 256, P.NA);
      const uint32_t idB = um::make_idesc(fmt, 256, P.NB > 0 ? P.NB : 16);
      const bool two_halves = P.NB > 0;
      const uint32_t dL2A = tmem_base, dL2B = tmem_base + (uint32_t)P.NA;
      bool ok = true;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        const uint32_t tpar = (uint32_t)(tl & 1);
        for (int pass = 0; pass < (two_halves ? 2 : 1) && ok; ++pass) {
          long long t0 = GPROF_T();
          GTRACE(0, tl * 100 + pass * 10 + 8);        // waiting for accumulator half
          if (!P.micro) ok = ok && um::mbar_wait(bar(GB_L2_EMPTY + pass), tpar ^ 1u, P.err, 14 + pass);
          GTRACE(0, tl * 100 + pass * 10 + 9);        // got accumulator half
          long long t1 = GPROF_T();
          pb_ += t1 - t0;
          const bool last_pass = pass == (two_halves ? 1 : 0);
          for (int c = 0; c < nch && ok; ++c) {
            long long t2 = GPROF_T();
            if (pass == 0 && !P.micro) ok = ok && um::mbar_wait(bar(GB_H1_FULL + c), tpar, P.err, 13);
            um::tc_fence_after();
            if (issuer) GTRACE(0, tl * 100 + pass * 10 + c);   // chunk ready, issuing
            long long t3 = GPROF_T();
            pa_ += t3 - t2;
            const int ksteps = (c == nch - 1 ? last_w : CH) >> 4;
            const uint32_t a0 = tmem_base + H1COL + (uint32_t)c * bufcols;       // 8 columns per K=16 step
            const uint32_t b0 = w2_lo + (uint32_t)(c * (CH >> 3)) * (lbo_w2 >> 4) + (pass ? w2_half_off : 0u);
            const uint32_t acc0 = c > 0 ? 1u : 0u;
            if (ok && issuer) {
              const uint32_t d = pass ? dL2B : dL2A, idesc = pass ? idB : idA;
              if (ksteps == 6) {
                // full 96-wide chunk: six back-to-back tcgen05.mma, all operands precomputed (the issuing
                // thread must stay well under the ~72-80 cycles one of these MMAs takes)
                um::mma2_ts(d, a0, um::desc64(b0), idesc, acc0);
#pragma unroll
                for (int k = 1; k < 6; ++k)
                  um::mma2_ts_acc(d, a0 + 8u * (uint32_t)k, um::desc64(b0 + (uint32_t)k * w2_kstep), idesc);
              } else {
#pragma unroll 1
                for (int k = 0; k < ksteps; ++k)
                  um::mma2_ts(d, a0 + 8u * (uint32_t)k, um::desc64(b0 + (uint32_t)k * w2_kstep), idesc,
                              acc0 | (uint32_t)(k > 0));
              }
              if (last_pass) um::commit2(bar(GB_H1_EMPTY + c));      // chunk buffer free for the next tile
              if (c == nch - 1) um::commit2(bar(GB_L2_FULL + pass));
            }
            __syncwarp();
            if (issuer) GTRACE(0, tl * 100 + pass * 10 + c + 50);   // chunk issued
            pc_ += GPROF_T() - t3;
          }
        }
      }
      if (prof && issuer) {
        long long* o = P.prof + (size_t)pair * 32;
        o[0] = clock64() - t_begin; o[1] = pa_; o[2] = pb_; o[3] = pc_; o[6] = ntiles;
      }
    }
  } else if (P.micro) {
    // microbenchmark: nobody but the MMA issuer works
  } else if (warp == 1) {
    // =================================== loader (one lane) ===================================
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");      // PS/PA tables of the pre-pass are complete and visible
      bool ok = true;
      const int pitch_b = (CH + GR_PITCH_PAD) * 2;
      const uint32_t pa_tile_bytes = 32u * (uint32_t)pitch_b;
      const uint32_t ps_row_bytes = (uint32_t)P.H1P * 2u;
      const size_t pa_chunk_stride = (size_t)P.NT * 32 * pitch_b;   // bytes between chunks in PA
      uint32_t st = 0, spar = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        int b0, n0;
        tile_coords(tl, b0, n0);
        const int pb = tl & 1;
        ok = um::mbar_wait(bar(GB_PS_EMPTY + pb), (uint32_t)(((tl >> 1) & 1) ^ 1), P.err, 51);
        if (!ok) break;
        um::mbar_expect_tx(bar(GB_PS_FULL + pb), 4u * ps_row_bytes);
        for (int i = 0; i < 4; ++i) {
          const int b = (b0 + i < P.B) ? b0 + i : P.B - 1;
          um::bulk_g2s(sPS + (uint32_t)(pb * P.ps_stage_bytes) + (uint32_t)i * ps_row_bytes,
                       reinterpret_cast<const unsigned char*>(P.ps) + (size_t)b * ps_row_bytes, ps_row_bytes,
                       bar(GB_PS_FULL + pb));
        }
        for (int c = 0; c < nch && ok; ++c) {
          ok = um::mbar_wait(bar(GB_PA_EMPTY + (int)st), spar ^ 1u, P.err, 52);
          if (!ok) break;
          um::mbar_expect_tx(bar(GB_PA_FULL + (int)st), pa_tile_bytes);
          um::bulk_g2s(sPA + st * (uint32_t)P.pa_stage_bytes,
                       reinterpret_cast<const unsigned char*>(P.pa) + (size_t)c * pa_chunk_stride +
                           (size_t)n0 * pitch_b,
                       pa_tile_bytes, bar(GB_PA_FULL + (int)st));
          if (++st == GR_PA_STAGES) { st = 0; spar ^= 1u; }
        }
      }
    }
  } else if (warp >= 4 && warp < GR_W2_0) {
    // ===== epilogue 1: h1 = relu(PS[b] + PA[n]) (packed fp16) -> TMEM (layer 2's A operand) =====
    const int q4 = warp & 3;                       // TMEM lane quarter == state b0 + q4 of the tile
    const int cg = ((warp - 4) >> 2) % GR_E1G;     // which share of each chunk's columns this warp builds
    const int cgrp = ((warp - 4) >> 2) / GR_E1G;   // which chunks (global chunk index % GR_E1C == cgrp)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t h1f0 = um::mapa(bar(GB_H1_FULL), 0);
    const int pitch_b = (CH + GR_PITCH_PAD) * 2;
    bool ok = true;
    uint32_t st = 0, spar = 0;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      const int pb = tl & 1;
      const uint32_t tpar = (uint32_t)(tl & 1);
      ok = um::mbar_wait(bar(GB_PS_FULL + pb), (uint32_t)((tl >> 1) & 1), P.err, 33);
      if (!ok) break;
      const unsigned char* ps_row = base_ptr + P.sm_ps + pb * P.ps_stage_bytes + q4 * (P.H1P * 2);
      for (int c = 0; c < nch; ++c) {
        if (GR_E1C > 1 && (tl * nch + c) % GR_E1C != cgrp) {      // another group's chunk: just advance the ring cursor
          if (++st == GR_PA_STAGES) { st = 0; spar ^= 1u; }
          continue;
        }
        long long t0 = GPROF_T();
        ok = um::mbar_wait(bar(GB_PA_FULL + (int)st), spar, P.err, 31);
        long long t1 = GPROF_T();
        ok = ok && um::mbar_wait(bar(GB_H1_EMPTY + c), tpar ^ 1u, P.err, 32);
        if (!ok) break;
        um::tc_fence_after();
        if (tid == 128) GTRACE(1, tl * 100 + c);            // chunk buffer free + PA here: start building
        long long t2 = GPROF_T();
        pa_ += t1 - t0;
        pb_ += t2 - t1;
        const int w = (c == nch - 1) ? last_w : CH;
        const uint4* pa4 = reinterpret_cast<const uint4*>(base_ptr + P.sm_pa + st * P.pa_stage_bytes + lane * pitch_b);
        const uint4* ps4 = reinterpret_cast<const uint4*>(ps_row + c * CH * 2);
        const uint32_t tcol = lane_addr + H1COL + (uint32_t)c * bufcols;
#if GR_E1G == 1
        if (w == 96) {
          // full chunk: straight-line code, 12 x (2 LDS.128 + 4 HFMA2.RELU) + 3 STTM.x16
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            uint32_t o[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 x = pa4[p * 4 + g];
              const uint4 y = ps4[p * 4 + g];
              o[g * 4 + 0] = um::add_relu2<PREC>(x.x, y.x);
              o[g * 4 + 1] = um::add_relu2<PREC>(x.y, y.y);
              o[g * 4 + 2] = um::add_relu2<PREC>(x.z, y.z);
              o[g * 4 + 3] = um::add_relu2<PREC>(x.w, y.w);
            }
            um::tmem_st16p(tcol + (uint32_t)(p * 16), o);
          }
        } else {
#pragma unroll 1
          for (int p = 0; p * 16 < w; ++p) {      // 16 activations (8 packed cells) per step
            uint32_t o[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint4 x = pa4[p * 2 + g];
              const uint4 y = ps4[p * 2 + g];
              o[g * 4 + 0] = um::add_relu2<PREC>(x.x, y.x);
              o[g * 4 + 1] = um::add_relu2<PREC>(x.y, y.y);
              o[g * 4 + 2] = um::add_relu2<PREC>(x.z, y.z);
              o[g * 4 + 3] = um::add_relu2<PREC>(x.w, y.w);
            }
            um::tmem_st8(tcol + (uint32_t)(p * 8), o);
          }
        }
#else
        // my units of 16 activations (8 packed TMEM cells); stored two units at a time (tcgen05.st.x16,
        // fewer and wider TMEM stores are measurably faster) with a single-unit tail
        const int nunit = w >> 4, ub = nunit * cg / GR_E1G, ue = nunit * (cg + 1) / GR_E1G;
        constexpr int MAXU = (6 + GR_E1G - 1) / GR_E1G;
#pragma unroll
        for (int i = 0; i < MAXU; i += 2) {
          const int u = ub + i;
          if (u + 1 < ue) {
            uint32_t o[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 x = pa4[u * 2 + g];
              const uint4 y = ps4[u * 2 + g];
              o[g * 4 + 0] = um::add_relu2<PREC>(x.x, y.x);
              o[g * 4 + 1] = um::add_relu2<PREC>(x.y, y.y);
              o[g * 4 + 2] = um::add_relu2<PREC>(x.z, y.z);
              o[g * 4 + 3] = um::add_relu2<PREC>(x.w, y.w);
            }
            um::tmem_st16p(tcol + (uint32_t)(u * 8), o);
          } else if (u < ue) {
            uint32_t o[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint4 x = pa4[u * 2 + g];
              const uint4 y = ps4[u * 2 + g];
              o[g * 4 + 0] = um::add_relu2<PREC>(x.x, y.x);
              o[g * 4 + 1] = um::add_relu2<PREC>(x.y, y.y);
              o[g * 4 + 2] = um::add_relu2<PREC>(x.z, y.z);
              o[g * 4 + 3] = um::add_relu2<PREC>(x.w, y.w);
            }
            um::tmem_st8(tcol + (uint32_t)(u * 8), o);
          }
        }
#endif
        __syncwarp();
        if (lane == 0) um::mbar_arrive_local(bar(GB_PA_EMPTY + (int)st));   // PA stage consumed
        long long t3 = GPROF_T();
        um::tmem_st_wait();
        um::tc_fence_before();
        __syncwarp();
        if (lane == 0) um::mbar_arrive_cluster(h1f0 + 8u * (uint32_t)c);
        if (tid == 128) GTRACE(1, tl * 100 + c + 50);       // chunk published
        long long t4 = GPROF_T();
        pc_ += t3 - t2;
        pd_ += t4 - t3;
        if (++st == GR_PA_STAGES) { st = 0; spar ^= 1u; }
      }
      __syncwarp();
      if (lane == 0) um::mbar_arrive_local(bar(GB_PS_EMPTY + pb));
    }
    if (prof && rank == 0 && tid == 128) {
      long long* o = P.prof + (size_t)pair * 32 + 8; o[0] = pa_; o[1] = pb_; o[2] = pc_; o[3] = pd_;
    }
  } else if (warp >= GR_W2_0) {
    // ================ epilogue 2 (4 x GR_E2G warps): L2 acc -> relu -> signed sum -> q ====================
    const int q4 = warp & 3, cg = (warp - GR_W2_0) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t l2e0 = um::mapa(bar(GB_L2_EMPTY + 0), 0);
    const uint32_t l2e1 = um::mapa(bar(GB_L2_EMPTY + 1), 0);
    const int npos = __ldg(reinterpret_cast<const int*>(P.blob[0] + P.off_c0));
    const float inv_scale = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 2);
    const float b3v = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 3);
    const int rloc = q4 * 32 + lane;
    bool ok = true;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int half = 0; half < 2; ++half) {
        const int hb = half ? P.NA : 0, hn = half ? P.NB : P.NA;
        long long t0 = GPROF_T();
        if (hn > 0 || half == 0) {
          ok = um::mbar_wait(bar(GB_L2_FULL + (hn > 0 ? half : 0)), (uint32_t)(tl & 1), P.err, 41);
          if (!ok) break;
          um::tc_fence_after();
        }
        long long t1 = GPROF_T();
        pa_ += t1 - t0;
        if (tid == 32 * GR_W2_0) GTRACE(2, tl * 100 + half * 10 + 0);      // accumulator half full
        const int units = hn >> 3, u0 = units * cg / GR_E2G, u1 = units * (cg + 1) / GR_E2G;
        const int j_begin = hb + u0 * 8, j_end = hb + u1 * 8;
        for (int jb = j_begin; jb < j_end; jb += GR_VR) {
          const int w = (j_end - jb < GR_VR) ? (j_end - jb) : GR_VR;
          uint32_t v[GR_VR];
#pragma unroll
          for (int p = 0; p < (GR_VR + 31) / 32; ++p) {
            if (p * 32 + 32 <= w) um::tmem_ld32p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
            else if (p * 32 < w) {
              const int rem = w - p * 32;             // 8, 16 or 24
              if (rem >= 16) um::tmem_ld16p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
              if (rem == 8) um::tmem_ld8p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
              if (rem == 24) um::tmem_ld8p(lane_addr + (uint32_t)(jb + p * 32 + 16), v + p * 32 + 16);
            }
          }
          um::tmem_ld_wait();
          { long long tt = GPROF_T(); pb_ += tt - t1; t1 = tt; }
          if (tid == 32 * GR_W2_0) GTRACE(2, tl * 100 + half * 10 + 1);    // loaded (released right after)
          if (jb + GR_VR >= j_end) {               // last round of this half: hand it back
            um::tc_fence_before();
            __syncwarp();
            if (lane == 0) um::mbar_arrive_cluster(half ? l2e1 : l2e0);
          }
          relu_signed_round(v, w, npos - jb, a0, a1, a2, a3);
        }
        if (j_begin >= j_end) {                 // nothing to drain for this warp: still hand back
          um::tc_fence_before();
          __syncwarp();
          if (lane == 0) um::mbar_arrive_cluster(half ? l2e1 : l2e0);
        }
        pc_ += GPROF_T() - t1;
        if (tid == 32 * GR_W2_0) GTRACE(2, tl * 100 + half * 10 + 2);      // math done
      }
      if (!ok) break;
      long long t5 = GPROF_T();
      const float acc = (a0 + a1) + (a2 + a3);
      float* qp = qpart + (tl & 1) * (128 * (GR_E2G - 1));
      if (cg) qp[(cg - 1) * 128 + rloc] = acc;
      um::named_bar_sync(1, 128 * GR_E2G);
      if (!cg) {
        const long long ct = 2ll * ((long long)pair + (long long)tl * npairs) + rank;
        if (ct < P.num_cta_tiles) {
          const int b = (int)(ct / P.NT) * 4 + q4, n = (int)(ct % P.NT) * 32 + lane;
          float tot = acc;
#pragma unroll
          for (int g = 0; g < GR_E2G - 1; ++g) tot += qp[g * 128 + rloc];
          if (b < P.B && n < P.N) P.q[(size_t)b * P.N + n] = fmaf(inv_scale, tot, b3v);
        }
      }
      pd_ += GPROF_T() - t5;
    }
    if (prof && rank == 0 && tid == 32 * GR_W2_0) {
      long long* o = P.prof + (size_t)pair * 32 + 16; o[0] = pa_; o[1] = pb_; o[2] = pc_; o[3] = pd_;
    }
  }
#undef GPROF_T
#undef GTRACE

  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  if (warp == 0) um::tmem_dealloc2(tmem_base, 512);
}

// host launcher --------------------------------------------------------------------------------
static int grid_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("RLC_UMMA_GRID");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode;
}

struct GridPlan {
  int CH, nch, nbuf, sm_w2, sm_pa, sm_ps, sm_qp, sm_bar, pa_stage, ps_stage, total;
};

static bool plan_grid(const PackGeom& G, GridPlan& p) {
  const int free_cols = 512 - G.H2P;
  if (G.H1P / 2 > free_cols) return false;    // the tile's activations must fit next to the accumulator
  int ch = 96;
  const char* e = getenv("RLC_UMMA_GRID_CH");
  if (e) { const int v = atoi(e); if (v >= 16 && v <= 96 && (v & 15) == 0) ch = v; }
  if (ch > G.H1P) ch = G.H1P;
  p.CH = ch;
  p.nch = (G.H1P + ch - 1) / ch;
  if (p.nch > GR_MAX_NCH) return false;
  p.nbuf = p.nch;
  p.sm_w2 = 0;
  p.pa_stage = 32 * (ch + GR_PITCH_PAD) * 2;
  p.sm_pa = ((G.off_w1 - G.off_w2) + 127) & ~127;
  p.ps_stage = 4 * G.H1P * 2;
  p.sm_ps = p.sm_pa + GR_PA_STAGES * p.pa_stage;
  p.sm_qp = (p.sm_ps + 2 * p.ps_stage + 127) & ~127;
  p.sm_bar = p.sm_qp + 2 * 128 * 4 * (GR_E2G > 1 ? GR_E2G - 1 : 1);
  p.total = p.sm_bar + GB_COUNT * 8 + 16 + 1024;
  return true;
}

static int rlc_eval_umma_grid(rlc_handle* h, const rlc_critic* c, const PackGeom& G, rlc_pack* pk, const float* s,
                              int B, const float* a, int N, int prec, float* q_out, cudaStream_t st) {
  GridPlan gp;
  if (!plan_grid(G, gp) || (size_t)gp.total > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  const int NT = (N + 31) / 32;
  const int pitch = gp.CH + GR_PITCH_PAD;
  const size_t nps = (size_t)B * G.H1P, npa = (size_t)gp.nch * NT * 32 * pitch;
  void* ws = nullptr;
  int rc = rlc_workspace(h, (nps + npa) * sizeof(unsigned short) + 256, &ws);
  if (rc) return rc;
  unsigned short* PS = (unsigned short*)ws;
  unsigned short* PA = PS + ((nps + 63) & ~(size_t)63);       // keep PA 128-byte aligned
  if (c->S <= 256 && c->A <= 256) {
    const int sg = (B + GR_PRE_ROWS - 1) / GR_PRE_ROWS, ag = (NT * 32 + GR_PRE_ROWS - 1) / GR_PRE_ROWS;
    const dim3 blocks((unsigned)(sg + ag), (unsigned)((G.H1P + 127) / 128));
    const size_t pre_smem = (size_t)GR_PRE_ROWS * (c->S > c->A ? c->S : c->A) * sizeof(float);
    if (prec == RLC_PREC_BF16)
      k_grid_parts8<RLC_PREC_BF16><<<blocks, 128, pre_smem, st>>>(c->theta, s, a, c->smin, c->smax, B, N, c->S, c->A, c->H1,
                                                                   c->H2, G.H1P, gp.CH, gp.nch, NT, sg, PS, PA, h->err_flag);
    else
      k_grid_parts8<RLC_PREC_FP16><<<blocks, 128, pre_smem, st>>>(c->theta, s, a, c->smin, c->smax, B, N, c->S, c->A, c->H1,
                                                                   c->H2, G.H1P, gp.CH, gp.nch, NT, sg, PS, PA, h->err_flag);
    RLC_LAUNCH_CHECK(h);
  } else {
    const dim3 blocks((unsigned)(B + NT * 32), (unsigned)((G.H1P + 127) / 128));
    if (prec == RLC_PREC_BF16)
      k_grid_parts<RLC_PREC_BF16><<<blocks, 128, 0, st>>>(c->theta, s, a, c->smin, c->smax, B, N, c->S, c->A,
                                                          c->H1, c->H2, G.H1P, gp.CH, gp.nch, NT, PS, PA);
    else
      k_grid_parts<RLC_PREC_FP16><<<blocks, 128, 0, st>>>(c->theta, s, a, c->smin, c->smax, B, N, c->S, c->A,
                                                          c->H1, c->H2, G.H1P, gp.CH, gp.nch, NT, PS, PA);
    RLC_LAUNCH_CHECK(h);
  }
  GridParams P;
  memset(&P, 0, sizeof(P));
  P.q = q_out; P.B = B; P.N = N;
  P.H1P = G.H1P; P.H2P = G.H2P; P.NA = G.NA; P.NB = G.NB;
  P.CH = gp.CH; P.nch = gp.nch; P.nbuf = gp.nbuf; P.NT = NT;
  P.num_cta_tiles = (long long)((B + 3) / 4) * NT;
  P.num_pair_tiles = (int)((P.num_cta_tiles + 1) / 2);
  P.ps = (const __half*)PS; P.pa = (const __half*)PA;
  P.blob[0] = (const unsigned char*)pk->dev;
  P.blob[1] = P.blob[0] + G.blob_bytes;
  P.off_w2 = G.off_w2; P.off_w1 = G.off_w1; P.off_c0 = G.off_c0;
  P.sm_w2 = gp.sm_w2; P.sm_pa = gp.sm_pa; P.sm_ps = gp.sm_ps; P.sm_qp = gp.sm_qp; P.sm_bar = gp.sm_bar;
  P.pa_stage_bytes = gp.pa_stage; P.ps_stage_bytes = gp.ps_stage;
  P.err = h->err_flag;

  int pairs = h->num_sms / 2;
  if (pairs > P.num_pair_tiles) pairs = P.num_pair_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(GR_THREADS);
  cfg.dynamicSmemBytes = (size_t)gp.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap our prologue with k_grid_parts
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("RLC_UMMA_PDL"); pdl = (e && e[0] == '1') ? 1 : 0; }   // measured gain 0.2 %: opt-in
  cfg.numAttrs = pdl ? 2 : 1;
  static int prof_on = -1;
  static long long* prof_dev = nullptr;
  if (prof_on < 0) { const char* e = getenv("RLC_UMMA_PROF"); prof_on = (e && e[0] == '1') ? 1 : 0; }
  if (prof_on) {
    if (!prof_dev) RLC_CUDA(cudaMalloc(&prof_dev, (4096 + 3 * 2048) * sizeof(long long)));
    RLC_CUDA(cudaMemsetAsync(prof_dev, 0, (4096 + 3 * 2048) * sizeof(long long), st));
    P.prof = prof_dev;
  }
  {
    static int micro = -1;     // debug: tensor-pipe microbenchmark (results are garbage), RLC_UMMA_MICRO=1
    if (micro < 0) { const char* mi = getenv("RLC_UMMA_MICRO"); micro = mi ? atoi(mi) : 0; }
    P.micro = micro;
  }
  void (*kern)(const GridParams) =
      P.prof ? (prec == RLC_PREC_BF16 ? k_critic_umma_grid<RLC_PREC_BF16, true> : k_critic_umma_grid<RLC_PREC_FP16, true>)
             : (prec == RLC_PREC_BF16 ? k_critic_umma_grid<RLC_PREC_BF16, false> : k_critic_umma_grid<RLC_PREC_FP16, false>);
  RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gp.total));
  RLC_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  RLC_LAUNCH_CHECK(h);
  if (prof_on) {   // debug only: synchronises
    static long long hp[4096 + 3 * 2048];
    RLC_CUDA(cudaStreamSynchronize(st));
    RLC_CUDA(cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost));
    if (getenv("RLC_UMMA_TRACE")) {
      const char* names[3] = {"MMA", "EP1", "EP2"};
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < 1000; ++i) {
          const long long code = hp[4096 + r * 2048 + 2 * i], tc = hp[4096 + r * 2048 + 2 * i + 1];
          if (tc == 0 && code == 0 && i > 0) break;
          if (code / 100 >= 6 && code / 100 <= 8) fprintf(stderr, "TRACE %s %lld %lld\n", names[r], code, tc);
        }
    }
    const long long* o = hp;
    const double T = (double)o[0], nt = (double)(o[6] > 0 ? o[6] : 1);
    fprintf(stderr, "[grid prof pair0] CH %d nbuf %d tiles %lld total %.0f cyc (%.0f/tile) | MMA: waitH1 %.1f%% waitL2E %.1f%% issue %.1f%% | "
            "ep1: waitPA %.1f%% waitH1E %.1f%% build+st %.1f%% wait::st+arrive %.1f%% | ep2: waitL2F %.1f%% ld+wait %.1f%% math %.1f%% tail %.1f%%\n",
            gp.CH, gp.nbuf, o[6], T, T / nt, 100 * o[1] / T, 100 * o[2] / T, 100 * o[3] / T, 100 * o[8] / T, 100 * o[9] / T,
            100 * o[10] / T, 100 * o[11] / T, 100 * o[16] / T, 100 * o[17] / T, 100 * o[18] / T, 100 * o[19] / T);
  }
  return RLC_OK;
}
