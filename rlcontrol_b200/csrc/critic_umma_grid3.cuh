// K1-grid x3 (precision RLC_PREC_FP16X3): the STRICT tensor-core mode of the fused T-in critic evaluation for a
// shared action grid a[N,A] (ForwardKL / ReverseKL quadrature, forwardkl_network.py:104-105,160-164,263-268).
// Included by critic_umma.cu after critic_umma_grid.cuh (shares its helpers, epilogue-2 arithmetic and head folding).
//
// Why: the reference computes in fp32.  One fp16 MMA per K step rounds BOTH operands to 11 bits and lands at
// max 5.7e-3 / rms 7e-4 of the exact Q on the cfg4 workload (profiles/r02_error_decomposition.jsonl); splitting only
// one operand still leaves 2.2e-3.  Here both operands are split into fp16 hi + lo and three products are
// accumulated in the same fp32 TMEM accumulator:
//     z = h_hi.W_hi + h_lo.W_hi + h_hi.W_lo          (the dropped h_lo.W_lo term is 2^-24 relative)
//   h   = relu(PS[b] + PA[n])  formed in fp32 from fp32 tables (one fp32 add, no operand rounding before the split)
//   h_hi = r16(h), h_lo = r16(h - h_hi)            W' = 2^k |w3_j| W2[:,j] (folded head, k_pack_head), same split
// => 22-bit operands: max 9e-6 of the exact Q in the fp64 emulation, i.e. fp32-class parity, at 3x the MMA work.
//
// Structure (a CTA pair, cta_group::2, M = 256 rows = 2 x (4 states x 32 grid actions), persistent):
//   * Neither the split weights (2 x 253 KB at 400-300) fit a pair's shared memory nor the split activations
//     (2 x 208 columns) fit tensor memory next to the accumulator, so BOTH stream: the weights in K chunks of KC
//     features through a ring of shared-memory stages (cp.async.bulk of the pre-packed core-matrix layout, L2-resident,
//     the same 5 chunks every tile), the activations through rotating TMEM slots (hi | lo, KC/2 columns each).
//   * per chunk and accumulator half the MMA thread issues 3 x KC/16 tcgen05.mma (A from TMEM, B from smem).
//   * the fp32 accumulator (H2P columns) is single-buffered and split into NP column PARTS of ~100 columns (304 = 112 +
//     96 + 96), each with its own full/empty barrier pair: the last chunk of a tile finishes part p, epilogue 2 pulls it
//     into registers and hands it back while the tensor pipe is busy with the other parts, so the hand-over of part p
//     has (NP-1)/NP of a chunk's MMA time to complete (with two halves the second half's drain queued behind the first
//     half's arithmetic and the pipe idled ~1000 cycles per tile: profiles/r02_k1_strict.md).
//   * K chunks are equalised (416 = 96 + 4 x 80) so that no chunk's MMAs are shorter than building the next chunk.
// Warps: 0 MMA issuer (leader CTA) | 1 table loader (PA tiles, PS rows; fp32) | 2 weight loader | 3 relay (tells the
// leader that THIS CTA's weight stage has landed: the MMA reads both CTAs' shared memory) | 4-11 epilogue 1 (two per
// TMEM lane quarter, each building half of a chunk's columns) | 12-19 epilogue 2.
#pragma once

// Two arithmetic modes share the kernel (template parameter MODE):
//   G3_X3 (RLC_PREC_FP16X3): z = h_hi.W_hi + h_lo.W_hi + h_hi.W_lo, all fp16 (kind::f16, K = 16 per MMA) -- fp32-class, 1e-5.
//   G3_C8 (RLC_PREC_FP16C8): the two correction terms on the FP8 pipe (kind::f8f6f4, K = 32 per MMA, twice the rate):
//          z = h_hi.W_hi [fp16] + e4m3(2^9 h_lo).e4m3(2^-9 W_hi) + e5m2(h_hi).e4m3(W_lo)
//          a correction only needs ~4 bits of its own (it is 2^-12 of the product): 2e-4 max of the exact Q on cfg4
//          (oracle_np.tin_eval_rounded(head="grid3c8")), 2/3 of the tensor-pipe time of G3_X3.
#define G3_X3 0
#define G3_C8 1
#ifndef RLC_G3_PDL_DEFAULT
#define RLC_G3_PDL_DEFAULT 0   // programmatic dependent launch of K1 behind its pre-pass (RLC_G3_PDL=0|1 overrides)
#endif
#ifndef G3_C8_E1W
#define G3_C8_E1W 8          // epilogue-1 warps of the C8 kernel (8 or 12); its epilogue 2 gets the rest of 24 warps
#endif
__host__ __device__ constexpr int g3_e1w(int mode) { return mode == G3_C8 ? G3_C8_E1W : 4; }   // epilogue-1 warps (C8 builds 3 operands)
// epilogue-2 column groups (x 4 lane quarters = warps)
#ifndef G3_C8_E2G
#define G3_C8_E2G ((24 - G3_C8_E1W) / 4)
#endif
__host__ __device__ constexpr int g3_e2g(int mode) { return mode == G3_C8 ? G3_C8_E2G : 4; }
__host__ __device__ constexpr int g3_threads(int mode) { return 32 * (4 + g3_e1w(mode) + 4 * g3_e2g(mode)); }
// accumulator columns an epilogue-2 thread holds per part
#ifndef G3_C8_VR
#define G3_C8_VR (g3_e2g(G3_C8) >= 4 ? 48 : 64)
#endif
__host__ __device__ constexpr int g3_vr(int mode) { return mode == G3_C8 ? G3_C8_VR : 48; }
#define G3_E2G g3_e2g(MODE)
#define G3_VR g3_vr(MODE)
#define G3_E1W g3_e1w(MODE)
#define G3_W2_0 (4 + G3_E1W) // first epilogue-2 warp
#define G3_THREADS g3_threads(MODE)
#define G3_C8_SA 9           // C8: h_lo is scaled by 2^SA into e4m3's normal range, W_hi by 2^-SA (W' < 2^10 by construction)
#define G3_C8_SB 0           // C8: W_lo and the e5m2 copy of h are used unscaled (e5m2 has fp16's range; W_lo < 2^-2 by construction)
#define G3_RANGE_LIMIT_C8 500.f   // C8: |h_lo| 2^9 must stay below e4m3's 448: |PS|, |PA| <= 500 -> h < 1024, h_lo <= 0.25
#define G3_MAX_NP 5          // accumulator column parts
#define G3_MAX_NCH 8
#define G3_PA_PAD 4          // floats of padding per PA row: pitch (KC+4)*4 B keeps the 16-byte LDS conflict-free
#define G3_MAX_WST 4
#define G3_MAX_SLOT 3
#ifndef POL_MAX_A
#define POL_MAX_A 8          // same bound as reduce.cu's policy-fused reductions
#endif
#define G3_RANGE_LIMIT 32000.f   // |PS|, |PA| above this cannot be added and split in fp16 (max 65504): error flag 91

struct Grid3Chunks {
  int nch;
  int start[G3_MAX_NCH], width[G3_MAX_NCH];   // feature ranges, multiples of 16
};

struct Grid3Parts {
  int np;
  int base[G3_MAX_NP + 1];    // part p = accumulator columns [base[p], base[p+1]), widths multiples of 16
};

struct Grid3Params {
  float* q;
  int B, N;
  int H1P, H2P;
  Grid3Parts parts;
  int KC, nslot;
  Grid3Chunks ch;
  int NT;
  long long num_cta_tiles;
  int num_pair_tiles;
  const float* ps;            // [B][H1P] fp32, PS[b][H1] = 1 (bias carrier of layer 2)
  const float* pa;            // [nch][NT*32][KC+PAD] fp32, zero rows past N
  const unsigned char* blob[2];
  int off_hi, off_lo, off_c0;   // X3: fp16 hi | fp16 lo;  C8: fp16 hi | [e4m3 hi8 | e4m3 lo8] starting at off_lo
  int w8_bytes;                 // C8: bytes of one e4m3 weight chunk of KC features = (KC/16) * lbo
  int lbo;                    // bytes between consecutive K groups of 8 in a weight blob = (H2P/2)*16
  int w_stages, pa_stages;
  int resident_hi;            // 1: W_hi stays in shared memory for the whole kernel (sm_whi), only W_lo streams
  int sm_whi, whi_bytes;
  int sm_w, sm_pa, sm_ps, sm_qp, sm_bar, w_stage_bytes, w_half_bytes, pa_stage_bytes, ps_stage_bytes;
  int* err;
  // fused per-state reduction (rlc_critic_eval_reduce_policy): 0 = none (q only), 1 = ForwardKL, 2 = ReverseKL.  Needs the
  // state-major tile order (a CTA owns whole state groups across all NT action blocks), see tile_coords.
  int fuse, state_major, A_pol;
  int NSG;                    // number of 4-state groups = ceil(B / 4)
  const float *fw, *fU, *fJ, *fmean, *flstd, *fv;     // w[N], U[A][N], J[N] (k_grid_logterms), mean/log_std [B][A], v[B] (RKL)
  float alpha, inv_btotal;
  float *loss_b, *dmean, *dlstd;
  long long* prof;            // RLC_UMMA_PROF=1: 32 x int64 per pair (cycle accounting per role)
  int micro;                  // RLC_UMMA_MICRO (diagnostic, output garbage): 1 = MMA issuer free-runs, everyone else idle;
                              // bit flags keeping the full barrier protocol: 2 = loaders copy nothing, 4 = epilogue 1 builds
                              // nothing, 8 = epilogue 2 loads/sums nothing, 32 = epilogue 1 loads and stores without the arithmetic
};

enum {
  B3_PA_FULL = 0,     // [4]  count 1 + tx  (local)   table loader -> ep1 (PA tile + PS slices of one chunk)
  B3_PA_EMPTY = 4,    // [4]  count E1W     (local)   ep1 warps -> table loader
                      // slots 8..11 unused (the PS rows travel with the PA tile of each chunk)
  B3_W_FULL = 12,     // [4]  count 1 + tx  (local)   weight loader -> relay
  B3_W_EMPTY = 16,    // [4]  count 1       (both)    MMA commit -> weight loaders
  B3_W_READY = 20,    // [4]  count 2       (leader)  relay of each CTA -> MMA
  B3_H1_FULL = 24,    // [3]  count 2 E1W   (leader)  ep1 (E1W warps x 2 CTAs) -> MMA
  B3_H1_EMPTY = 27,   // [3]  count 1       (both)    MMA commit -> ep1
  B3_L2_FULL = 30,    // [5]  count 1       (both)    MMA commit -> ep2 (one per accumulator part)
  B3_L2_EMPTY = 35,   // [5]  count 8 E2G   (leader)  ep2 (4 E2G warps x 2 CTAs) -> MMA
  B3_COUNT = 40
};

// Pre-pass, fp32 tables (8 rows per CTA as k_grid_parts8):  PS[b][j] = b1[j] + sum_k clip(s[b][k]) W1[k][j]  (j < H1),
// PS[b][H1] = 1;  PA[c][n][jj] = sum_k a[n][k] W1[S+k][start_c + jj].  Same fp32 FMA order as the fp16 pre-pass.  Both tables
// are stored multiplied by 1/2 (exact), see ep1_unit.
__global__ void __launch_bounds__(128)
k_grid3_parts(const float* __restrict__ theta, const float* __restrict__ s, const float* __restrict__ a,
              const float* __restrict__ smin, const float* __restrict__ smax, int B, int N, int S, int A, int H1, int H2,
              int H1P, int KC, Grid3Chunks ch, int NT, int state_groups, float* __restrict__ PS,
              float* __restrict__ PA, int* __restrict__ err, float limit) {
  extern __shared__ float xs[];                               // [GR_PRE_ROWS][K]
  // programmatic dependent launch: the evaluation kernel behind this pre-pass may become resident and run its prologue
  // (resident W_hi copy, barrier set-up, TMEM allocation: ~12 us) now; its table loader waits (griddepcontrol.wait) for
  // this grid to complete before it touches PS/PA.  A no-op when the next launch does not ask for it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
  bool bad = false;
  if (is_state) {
#pragma unroll
    for (int r = 0; r < GR_PRE_ROWS; ++r)
      if (r0 + r < B) {
        const float v = (j == H1) ? 1.f : acc[r];
        bad = bad || !(fabsf(v) <= limit);
        PS[(size_t)(r0 + r) * H1P + j] = 0.5f * v;        // tables are stored HALVED (exact): relu(x) = x/2 + |x/2|
      }
  } else {
    int c = 0;
    while (c + 1 < ch.nch && !(j >= ch.start[c] && j < ch.start[c] + ch.width[c])) ++c;
    const int jj = j - ch.start[c];
    const int pitch = KC + G3_PA_PAD;
#pragma unroll
    for (int r = 0; r < GR_PRE_ROWS; ++r) {
      const int n = r0 + r;                                   // 0 .. NT*32-1 (rows past N are zero)
      if (n < NT * 32) {
        const float v = (j < H1 && n < N) ? acc[r] : 0.f;
        bad = bad || !(fabsf(v) <= limit);
        PA[((size_t)c * NT * 32 + n) * pitch + jj] = 0.5f * v;
      }
    }
  }
  if (bad) atomicCAS(err, 0, 91);
}

// Weight pack for the split mode: W' = fl32(sw_j * W2[k][j]) (bias row k = H1: sw_j * b2_j), sw_j = scale |w3_j| with the
// column permutation and `scale` written by k_pack_head + k_pack_scale3;  hi = r16(W'), lo = r16(W' - hi).  Same
// no-swizzle K-major core-matrix layout and pair split as k_pack_umma's W2 region.
__global__ void k_pack_x3(const float* __restrict__ theta, PackGeom G, Grid3Parts parts, int mode, unsigned char* blob0,
                          unsigned char* blob1) {
  const ThetaView t = theta_view(RLC_TIN, G.S, G.A, G.H1, G.H2);
  const float* W2 = theta + t.oW2;   // [H1][H2]
  const float* b2 = theta + t.ob2;
  const float* w3 = theta + t.ow3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)G.H1P * G.H2P) return;
  const int k = (int)(gid / G.H2P), n = (int)(gid % G.H2P);
  // column n of accumulator part p: the pair splits every part in two, CTA `rank` holds rows [base/2, base/2 + width/2)
  int p = 0;
  while (p + 1 < parts.np && n >= parts.base[p + 1]) ++p;
  const int hN = (parts.base[p + 1] - parts.base[p]) / 2, m = n - parts.base[p];
  const int rank = m / hN, l = parts.base[p] / 2 + m % hN;
  float v = 0.f;
  const int j = reinterpret_cast<const int*>(blob0 + G.off_w3)[n];   // written by k_pack_head
  if (j >= 0 && k <= G.H1) {
    const float sw = reinterpret_cast<const float*>(blob0 + G.off_c0)[1] * fabsf(w3[j]);
    v = sw * (k < G.H1 ? W2[(long long)k * G.H2 + j] : b2[j]);       // one fp32 rounding
  }
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half hi = __float2half_rn(v);
  unsigned char* blob = rank ? blob1 : blob0;
  const long long lbo = (G.H2P / 2) * 16;
  const long long off = (long long)(k / 8) * lbo + (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 8) * 2;
  *reinterpret_cast<__half*>(blob + G.off_w2 + off) = hi;
  if (mode == G3_X3) {
    *reinterpret_cast<__half*>(blob + G.off_w1 + off) = __float2half_rn(v - __half2float(hi));   // off_w1 = lo region
  } else {
    // 8-bit operands: same core-matrix layout with 16 K-elements per 16-byte row; [hi8 | lo8] after the fp16 region
    const long long off8 = (long long)(k / 16) * lbo + (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 16);
    const long long sz8 = (long long)(G.H1P / 16) * lbo;
    blob[G.off_w1 + off8] = (unsigned char)__nv_cvt_float_to_fp8(ldexpf(__half2float(hi), -G3_C8_SA), __NV_SATFINITE, __NV_E4M3);
    blob[G.off_w1 + sz8 + off8] =
        (unsigned char)__nv_cvt_float_to_fp8(ldexpf(v - __half2float(hi), G3_C8_SB), __NV_SATFINITE, __NV_E4M3);
  }
}

// Extra power-of-two column scale of the split mode: 2^e2 with 2^e2 * max(|W2|, |b2|) in [2^8, 2^9), so that W' < 2^10
// and its lo part (<= 2^-12 W') stays a normal fp16 number.  Multiplies the scale k_pack_head wrote (exact).
__global__ void __launch_bounds__(1024) k_pack_scale3(const float* __restrict__ theta, PackGeom G, unsigned char* blob0,
                                                       unsigned char* blob1) {
  const ThetaView t = theta_view(RLC_TIN, G.S, G.A, G.H1, G.H2);
  const float* W2 = theta + t.oW2;
  const long long n = (long long)G.H1 * G.H2 + G.H2;          // W2 and b2 are adjacent in theta
  float mx = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) mx = fmaxf(mx, fabsf(W2[i]));
  __shared__ float red[32];
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  mx = warp_max(red[threadIdx.x]);
  if (threadIdx.x != 0) return;
  float s2 = 1.f;
  if (mx > 0.f && isfinite(mx)) {
    int e = 0;
    (void)frexpf(mx, &e);              // mx = m * 2^e, m in [0.5,1)
    s2 = ldexpf(1.f, 9 - e);
  }
  for (int r = 0; r < 2; ++r) {
    float* c0 = reinterpret_cast<float*>((r ? blob1 : blob0) + G.off_c0);
    const float sc = c0[1] * s2;
    c0[1] = sc;
    c0[2] = 1.f / sc;
  }
}

namespace um {
// One accumulator part of one K chunk: KS K-steps x 3 operand-term products, back to back with precomputed operands.
template <int KS>
__device__ __forceinline__ void issue_half3(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t bh, uint32_t bl,
                                            uint32_t kstep, uint32_t idesc, uint32_t acc0) {
  mma2_ts(d, a_hi, desc64(bh), idesc, acc0);
  mma2_ts_acc(d, a_lo, desc64(bh), idesc);
  mma2_ts_acc(d, a_hi, desc64(bl), idesc);
#pragma unroll
  for (int k = 1; k < KS; ++k) {
    mma2_ts_acc(d, a_hi + 8u * (uint32_t)k, desc64(bh + (uint32_t)k * kstep), idesc);
    mma2_ts_acc(d, a_lo + 8u * (uint32_t)k, desc64(bh + (uint32_t)k * kstep), idesc);
    mma2_ts_acc(d, a_hi + 8u * (uint32_t)k, desc64(bl + (uint32_t)k * kstep), idesc);
  }
}
// D[tmem] += A[tmem] * B[smem]^T on the FP8 pipe (kind::f8f6f4, K = 32 per instruction; operand formats in idesc)
__device__ __forceinline__ void mma2_ts_f8(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_ab(uint32_t afmt, uint32_t bfmt, int M, int N) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// One accumulator part of one K chunk in C8 mode: KS K=16 steps of the fp16 product, then KS/2 K=32 steps of each FP8
// correction (a_lo8 x b_hi8: e4m3 x e4m3;  a_hi8 x b_lo8: e5m2 x e4m3).
template <int KS>
__device__ __forceinline__ void issue_part_c8(uint32_t d, uint32_t a_hi, uint32_t a_lo8, uint32_t a_hi8, uint32_t b16,
                                              uint32_t b_hi8, uint32_t b_lo8, uint32_t kstep, uint32_t id16, uint32_t id_t2,
                                              uint32_t id_t3, uint32_t acc0) {
  mma2_ts(d, a_hi, desc64(b16), id16, acc0);
#pragma unroll
  for (int k = 1; k < KS; ++k) mma2_ts_acc(d, a_hi + 8u * (uint32_t)k, desc64(b16 + (uint32_t)k * kstep), id16);
#pragma unroll
  for (int k = 0; k < KS / 2; ++k) mma2_ts_f8(d, a_lo8 + 8u * (uint32_t)k, desc64(b_hi8 + (uint32_t)k * kstep), id_t2);
#pragma unroll
  for (int k = 0; k < KS / 2; ++k) mma2_ts_f8(d, a_hi8 + 8u * (uint32_t)k, desc64(b_lo8 + (uint32_t)k * kstep), id_t3);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
}  // namespace um

// 2 * sign * relu(z) summed over NU units of 8 columns: relu(z) = (z + |z|) / 2 exactly, so the ReLU is an FADD with a free
// |.| source modifier on the FMA pipe instead of an FMNMX on the (half-rate, shared with epilogue 1's min/max, permutes and
// conversions) ALU pipe; the factor 2 is folded into the power-of-two output scale.
template <int NU>
__device__ __forceinline__ void relu2_signed_sum(const uint32_t* v, int kpos, float& a0, float& a1, float& a2, float& a3) {
#pragma unroll
  for (int p = 0; p < NU; ++p) {
    const float sg = (p < kpos) ? 1.f : -1.f;
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float z0 = __uint_as_float(v[p * 8 + e + 0]), z1 = __uint_as_float(v[p * 8 + e + 1]);
      const float z2 = __uint_as_float(v[p * 8 + e + 2]), z3 = __uint_as_float(v[p * 8 + e + 3]);
      a0 = fmaf(z0 + fabsf(z0), sg, a0);
      a1 = fmaf(z1 + fabsf(z1), sg, a1);
      a2 = fmaf(z2 + fabsf(z2), sg, a2);
      a3 = fmaf(z3 + fabsf(z3), sg, a3);
    }
  }
}

// relu_signed_round for at most G3_VR = 64 columns held in registers (all indices static); accumulates 2x the signed sum.
template <int VR>
__device__ __forceinline__ void relu_signed_round8(const uint32_t* v, int w, int rel, float& a0, float& a1, float& a2,
                                                   float& a3) {
  if (rel <= 0 || rel >= w || (rel & 7) == 0) {
    const int kpos = rel <= 0 ? 0 : (rel >= w ? 8 : (rel >> 3));
    if constexpr (VR >= 64) {
      if ((w >> 3) == 8) { relu2_signed_sum<8>(v, kpos, a0, a1, a2, a3); return; }
      if ((w >> 3) == 7) { relu2_signed_sum<7>(v, kpos, a0, a1, a2, a3); return; }
    }
    switch (w >> 3) {
      case 6: relu2_signed_sum<6>(v, kpos, a0, a1, a2, a3); break;
      case 5: relu2_signed_sum<5>(v, kpos, a0, a1, a2, a3); break;
      case 4: relu2_signed_sum<4>(v, kpos, a0, a1, a2, a3); break;
      case 3: relu2_signed_sum<3>(v, kpos, a0, a1, a2, a3); break;
      case 2: relu2_signed_sum<2>(v, kpos, a0, a1, a2, a3); break;
      default: relu2_signed_sum<1>(v, kpos, a0, a1, a2, a3); break;
    }
  } else {
#pragma unroll
    for (int p = 0; p < VR / 8; ++p) {
      if (p * 8 < w) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          a0 = fmaf(__uint_as_float(v[p * 8 + e]) + fabsf(__uint_as_float(v[p * 8 + e])), (p * 8 + e >= rel) ? -1.f : 1.f, a0);
      }
    }
  }
}

// Epilogue 1, one unit of 16 activations of one row: h = relu(PS + PA) in fp32 (one add), hi = fp16(h) packed in 8 cells;
// X3: lo[0..7] = fp16(h - hi) cells;  C8: lo[0..3] = e4m3(2^SA (h - hi)) cells, lo[4..7] = e5m2(hi) cells.
template <int MODE>
__device__ __forceinline__ void ep1_unit(const float4* __restrict__ pa4, const float4* __restrict__ ps4, int u, uint32_t* hi,
                                         uint32_t* lo) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 x = pa4[u * 4 + g];
    const float4 y = ps4[u * 4 + g];
    // packed fp32 adds / multiplies (FADD2 / FMUL2): half the instructions of the scalar forms, same roundings
    const float2 s01 = __fadd2_rn(make_float2(x.x, x.y), make_float2(y.x, y.y));
    const float2 s23 = __fadd2_rn(make_float2(x.z, x.w), make_float2(y.z, y.w));
    // the tables hold PS/2 and PA/2 (exact), so s = (PS + PA)/2 and relu(PS + PA) = s + |s| EXACTLY: an FADD with a free |.|
    // source modifier on the FMA pipe instead of an FMNMX on the half-rate ALU pipe (which the conversions and permutes need)
    const float2 v01 = make_float2(s01.x + fabsf(s01.x), s01.y + fabsf(s01.y));
    const float2 v23 = make_float2(s23.x + fabsf(s23.x), s23.y + fabsf(s23.y));
    const __half2 h01 = __float22half2_rn(v01), h23 = __float22half2_rn(v23);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    hi[g * 2 + 0] = *reinterpret_cast<const uint32_t*>(&h01);
    hi[g * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
    const float2 d01 = __fadd2_rn(v01, make_float2(-f01.x, -f01.y)), d23 = __fadd2_rn(v23, make_float2(-f23.x, -f23.y));   // exact
    if (MODE == G3_C8) {
      constexpr float SA = (float)(1 << G3_C8_SA);
      const uint32_t l01 = __nv_cvt_float2_to_fp8x2(__fmul2_rn(d01, make_float2(SA, SA)), __NV_SATFINITE, __NV_E4M3);
      const uint32_t l23 = __nv_cvt_float2_to_fp8x2(__fmul2_rn(d23, make_float2(SA, SA)), __NV_SATFINITE, __NV_E4M3);
      lo[g] = l01 | (l23 << 16);                      // four consecutive K elements per 32-bit cell, lowest first
      // e5m2 IS the upper byte of fp16 (same sign and exponent fields, 2 of the 10 mantissa bits): round h_hi to it with
      // one packed integer add (+half an e5m2 ulp; h >= 0 and finite, so no carry crosses a half) and pick the four upper
      // bytes -- no FP8 conversion instruction (they are the slow ones) and no scaling of h
      lo[4 + g] = __byte_perm(hi[g * 2 + 0] + 0x00800080u, hi[g * 2 + 1] + 0x00800080u, 0x7531);
    } else {
      const __half2 l01 = __float22half2_rn(d01), l23 = __float22half2_rn(d23);
      lo[g * 2 + 0] = *reinterpret_cast<const uint32_t*>(&l01);
      lo[g * 2 + 1] = *reinterpret_cast<const uint32_t*>(&l23);
    }
  }
}

// FA: 0 = evaluation only; > 0 = the per-state policy reduction fused into epilogue 2, compiled for action dimensions <= FA
// (the reduction's running sums live in registers, so the plain kernel must not carry them)
template <bool PROF, int MODE, int FA>
__global__ void __launch_bounds__(G3_THREADS, 1) k_critic_umma_grid3(const Grid3Params P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = um::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = um::cta_rank();
  const uint32_t pair = um::cluster_id_x();
  const uint32_t npairs = um::num_clusters_x();

  const uint32_t sW = base + P.sm_w, sPA = base + P.sm_pa, sBar = base + P.sm_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + P.sm_bar + B3_COUNT * 8);
  float* qpart = reinterpret_cast<float*>(base_ptr + P.sm_qp);   // [2][E2G-1][128] partial sums
  auto bar = [&](int i) -> uint32_t { return sBar + 8u * (uint32_t)i; };

  if (P.resident_hi) {        // resident W_hi (this CTA's half) -> smem, once
    const uint4* src = reinterpret_cast<const uint4*>(P.blob[rank] + P.off_hi);
    uint4* dst = reinterpret_cast<uint4*>(base_ptr + P.sm_whi);
    for (int i = tid; i < (P.whi_bytes >> 4); i += G3_THREADS) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      um::mbar_init(bar(B3_PA_FULL + i), 1);
      um::mbar_init(bar(B3_PA_EMPTY + i), G3_E1W);
      um::mbar_init(bar(B3_W_FULL + i), 1);
      um::mbar_init(bar(B3_W_EMPTY + i), 1);
      um::mbar_init(bar(B3_W_READY + i), 2);
    }
    for (int i = 0; i < G3_MAX_NP; ++i) {
      um::mbar_init(bar(B3_L2_FULL + i), 1);
      um::mbar_init(bar(B3_L2_EMPTY + i), 8 * G3_E2G);
    }
    for (int i = 0; i < G3_MAX_SLOT; ++i) {
      um::mbar_init(bar(B3_H1_FULL + i), 2 * G3_E1W);
      um::mbar_init(bar(B3_H1_EMPTY + i), 1);
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

  long long pa_ = 0, pb_ = 0, pc_ = 0, pd_ = 0;   // PROF accumulators (meaning per role)
  const bool prof = PROF && P.prof != nullptr;
  const long long t_begin = prof ? clock64() : 0;
#define G3T() ((PROF && prof) ? clock64() : 0)
  // event trace (RLC_UMMA_TRACE, pair 0 leader CTA, tiles 8..11): role r appends (tile*1000 + code, clock) at prof[4096 + r*2048 ..]
  int tr_n = 0;
  const bool tracing = PROF && prof && pair == 0 && rank == 0;
#define G3TR(role, tl, code)                                                      \
  do {                                                                            \
    if (PROF && tracing && (tl) >= 8 && (tl) < 12 && tr_n < 1000) {               \
      long long* tb = P.prof + 4096 + (role) * 2048;                              \
      tb[2 * tr_n] = (tl) * 1000 + (code);                                        \
      tb[2 * tr_n + 1] = clock64() - t_begin;                                     \
      ++tr_n;                                                                     \
    }                                                                             \
  } while (0)

  // tiles of this pair.  Round-robin order: pair-tile t = pair + tl * npairs covers CTA tiles 2t, 2t+1 (any B).  State-major
  // order (fused reduction): CTA g = 2*pair + rank owns state groups g, g + 2*npairs, ... and walks each group's NT action
  // blocks consecutively, so that a state's N evaluations all pass through the same epilogue-2 threads.
  const int Gc = 2 * (int)npairs;
  const int ntiles = P.state_major
                         ? ((P.NSG > 2 * (int)pair) ? ((P.NSG - 2 * (int)pair + Gc - 1) / Gc) * P.NT : 0)
                         : ((P.num_pair_tiles > (int)pair) ? (P.num_pair_tiles - (int)pair + (int)npairs - 1) / (int)npairs : 0);
  const int nch = P.ch.nch, KC = P.KC;
  const uint32_t SLOT0 = (uint32_t)P.H2P;             // TMEM column of activation slot 0: [hi KC/2 | lo KC/2]
  const uint32_t nslot = (uint32_t)P.nslot, wst_n = (uint32_t)P.w_stages, past_n = (uint32_t)P.pa_stages;
  auto tile_coords = [&](int tl, int& b0, int& n0) -> bool {       // returns false for a padding tile (computed, never stored)
    if (P.state_major) {
      int sg = 2 * (int)pair + (int)rank + (tl / P.NT) * Gc;
      const bool valid = sg < P.NSG;
      if (!valid) sg = P.NSG - 1;
      b0 = sg * 4;
      n0 = (tl % P.NT) * 32;
      return valid;
    }
    long long ct = 2ll * ((long long)pair + (long long)tl * npairs) + rank;
    const bool valid = ct < P.num_cta_tiles;
    if (!valid) ct = P.num_cta_tiles - 1;                  // odd tail: recompute the last tile, never stored twice
    b0 = (int)(ct / P.NT) * 4;
    n0 = (int)(ct % P.NT) * 32;
    return valid;
  };

  // One warp's share of one activation chunk: h = relu(PS[b] + PA[n]) in fp32 -> the chunk's A operands in its TMEM slot.
  // TMEM lane quarter == state b0 + q4 of the tile, lane == grid action n0 + lane: a 16-byte PA load covers 32 distinct
  // rows (4 full wavefronts), a PS load is one broadcast wavefront.  (Measured alternative: lane <-> (state lane/8,
  // action 8*q4 + lane%8) reads 8 PA rows per warp but a 128-bit load still costs one wavefront per quarter-warp, and the
  // PS load stops being a single broadcast: 8 wavefronts per pair instead of 5 -- slower, 1.83 vs 1.76 ms.)
  // k of nk: the chunk's 16-feature units go round-robin over the nk epilogue-1 warps of a lane quarter.
  // (Measured alternative, round 2: the epilogue-2 warps -- idle ~60 % of a tile, waiting for the accumulator -- building
  // shares too, either of every chunk or only of the mid-tile chunks with pre-arrivals for the others: every chunk's
  // hand-shake then waits for the slowest of 24 warps, some of them busy draining: 1.93 - 2.76 ms against 1.45.)
  const uint32_t b_lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t b_h1f0 = um::mapa(bar(B3_H1_FULL), 0);
  auto build_chunk = [&](int tl, int c, int k, int nk, uint32_t& st, uint32_t& ppar, uint32_t& slot, uint32_t& spar) -> bool {
    const int q4 = warp & 3;
    const int pitch_b = (KC + G3_PA_PAD) * 4;
    const uint32_t alo = (uint32_t)(KC >> 1);
    long long t0 = G3T();
    if (!um::mbar_wait(bar(B3_PA_FULL + (int)st), ppar, P.err, 31)) return false;
    long long t1 = G3T();
    const int nunit = P.ch.width[c] >> 4;
    const float4* pa4 = reinterpret_cast<const float4*>(base_ptr + P.sm_pa + st * P.pa_stage_bytes + lane * pitch_b);
    const float4* ps4 = reinterpret_cast<const float4*>(base_ptr + P.sm_pa + st * P.pa_stage_bytes + 32 * pitch_b +
                                                        q4 * (KC * 4));
    const uint32_t tcol = b_lane_addr + SLOT0 + slot * (uint32_t)KC;
    if (!um::mbar_wait(bar(B3_H1_EMPTY + (int)slot), spar ^ 1u, P.err, 32)) return false;
    um::tc_fence_after();
    long long t2 = G3T();
    pa_ += t1 - t0; pb_ += t2 - t1;
    if (tid == 128) G3TR(1, tl, c * 10 + 1);        // slot free + PA here: start building
    // (Measured alternative for C8: computing the warp's share into registers BEFORE the slot is free and only storing
    // afterwards -- 48 live registers per thread at the 72-register budget of 896 threads: spills, 1.58 vs 1.48 ms.)
#pragma unroll 1
    for (int u = k; u < nunit && !(P.micro & 4); u += nk) {   // 16 activations -> 8 packed hi cells + the lo operand cells
      uint32_t hi[8], lo[8];
      if (P.micro & 32) {     // timing decomposition only: the loads and the TMEM stores without the arithmetic
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);   // (& 64: without the PA loads, & 128: without the PS loads, & 256: no stores)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 x = make_float4(1.f, 2.f, 3.f, 4.f), y = x;
          if (!(P.micro & 64)) x = pa4[u * 4 + g];
          if (!(P.micro & 128)) y = ps4[u * 4 + g];
          t.x += x.x + y.x; t.y += x.y + y.y; t.z += x.z + y.z; t.w += x.w + y.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { hi[i] = __float_as_uint((i & 1) ? t.x + t.y : t.z + t.w); lo[i] = hi[i]; }
        if ((P.micro & 256) && hi[0] != 0x12345u) continue;
      } else
      ep1_unit<MODE>(pa4, ps4, u, hi, lo);
      um::tmem_st8(tcol + (uint32_t)(u * 8), hi);
      if (MODE == G3_C8) {
        um::tmem_st4(tcol + alo + (uint32_t)(u * 4), lo);
        um::tmem_st4(tcol + alo + (uint32_t)(KC >> 2) + (uint32_t)(u * 4), lo + 4);
      } else {
        um::tmem_st8(tcol + alo + (uint32_t)(u * 8), lo);
      }
    }
    __syncwarp();
    if (lane == 0) um::mbar_arrive_local(bar(B3_PA_EMPTY + (int)st));   // PA stage consumed
    long long t3 = G3T();
    um::tmem_st_wait();
    um::tc_fence_before();
    __syncwarp();
    if (lane == 0) um::mbar_arrive_cluster(b_h1f0 + 8u * slot);
    if (tid == 128) G3TR(1, tl, c * 10 + 2);        // chunk published
    pc_ += t3 - t2; pd_ += G3T() - t3;
    if (++st == past_n) { st = 0; ppar ^= 1u; }
    if (++slot == nslot) { slot = 0; spar ^= 1u; }
    return true;
  };

  if (warp == 0) {
    // =================================== MMA issuer (leader CTA) ===================================
    if (rank == 0) {
      const bool issuer = um::elect_one();
      const uint32_t lbo = (uint32_t)P.lbo;
      const uint32_t kstep = (2u * lbo) >> 4;                  // one K=16 step of a weight chunk (>>4 of bytes)
      const uint32_t lo_off = (uint32_t)P.w_half_bytes >> 4;   // lo chunk inside a weight stage (streaming both)
      const uint32_t whi_lo = um::desc_lo(base + (uint32_t)P.sm_whi, lbo);   // resident W_hi
      const int np = P.parts.np;
      const uint32_t alo = (uint32_t)(KC >> 1);
      bool ok = true;
      uint32_t slot = 0, spar = 0, ws = 0, wpar = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        const uint32_t tpar = (uint32_t)(tl & 1);
        for (int c = 0; c < nch && ok; ++c) {
          long long t0 = G3T();
          if (P.micro != 1) ok = ok && um::mbar_wait(bar(B3_H1_FULL + (int)slot), spar, P.err, 13);
          long long t1 = G3T();
          if (issuer) G3TR(0, tl, c * 10 + 0);          // H1 chunk ready
          if (P.micro != 1) ok = ok && um::mbar_wait(bar(B3_W_READY + (int)ws), wpar, P.err, 16);
          um::tc_fence_after();
          long long t2 = G3T();
          if (issuer) G3TR(0, tl, c * 10 + 1);          // weight stage ready
          pa_ += t1 - t0;
          pb_ += t2 - t1;
          const int ksteps = P.ch.width[c] >> 4;
          const uint32_t a_hi = tmem_base + SLOT0 + slot * (uint32_t)KC;    // 8 columns per K=16 step
          const uint32_t a_lo = a_hi + alo;
          const uint32_t b_st = um::desc_lo(sW + ws * (uint32_t)P.w_stage_bytes, lbo);
          const uint32_t b_hi = P.resident_hi ? whi_lo + (uint32_t)(P.ch.start[c] >> 3) * (lbo >> 4) : b_st;
          const uint32_t b_lo = P.resident_hi ? b_st : b_st + lo_off;     // X3: fp16 lo chunk; C8: e4m3 hi8 chunk, lo8 after it
          const uint32_t b_lo8 = b_lo + ((uint32_t)P.w8_bytes >> 4);
          const uint32_t a_hi8 = a_lo + (uint32_t)(KC >> 2);              // C8 slot: [hi16 KC/2 | lo8 KC/4 | hi8 KC/4] columns
          const uint32_t acc0 = c > 0 ? 1u : 0u;
#pragma unroll 1
          for (int p = 0; p < np && ok; ++p) {
            if (c == 0 && P.micro != 1) {
              long long t3 = G3T();
              ok = ok && um::mbar_wait(bar(B3_L2_EMPTY + p), tpar ^ 1u, P.err, 60 + p);
              um::tc_fence_after();
              if (p) pd_ += G3T() - t3; else pc_ += G3T() - t3;
              if (issuer) G3TR(0, tl, 100 + p);         // accumulator part p handed back
            }
            if (ok && issuer) {
              // per accumulator part: TMEM column, instruction descriptor (N = part width), B row offset (>>4 of bytes)
              const int pb0 = P.parts.base[p], pw = P.parts.base[p + 1] - pb0;
              const uint32_t d = tmem_base + (uint32_t)pb0, idesc = um::make_idesc(0u, 256, pw);
              const uint32_t bh = b_hi + (uint32_t)(pb0 >> 1), bl = b_lo + (uint32_t)(pb0 >> 1);
              if (MODE == G3_C8) {
                const uint32_t id_t2 = um::make_idesc_ab(0u, 0u, 256, pw), id_t3 = um::make_idesc_ab(1u, 0u, 256, pw);
                const uint32_t bl8 = b_lo8 + (uint32_t)(pb0 >> 1);
                switch (ksteps) {
                  case 6: um::issue_part_c8<6>(d, a_hi, a_lo, a_hi8, bh, bl, bl8, kstep, idesc, id_t2, id_t3, acc0); break;
                  case 4: um::issue_part_c8<4>(d, a_hi, a_lo, a_hi8, bh, bl, bl8, kstep, idesc, id_t2, id_t3, acc0); break;
                  default: um::issue_part_c8<2>(d, a_hi, a_lo, a_hi8, bh, bl, bl8, kstep, idesc, id_t2, id_t3, acc0); break;
                }
              } else
              switch (ksteps) {
                case 6: um::issue_half3<6>(d, a_hi, a_lo, bh, bl, kstep, idesc, acc0); break;
                case 5: um::issue_half3<5>(d, a_hi, a_lo, bh, bl, kstep, idesc, acc0); break;
                case 4: um::issue_half3<4>(d, a_hi, a_lo, bh, bl, kstep, idesc, acc0); break;
                default:
#pragma unroll 1
                  for (int k = 0; k < ksteps; ++k) {
                    um::mma2_ts(d, a_hi + 8u * (uint32_t)k, um::desc64(bh + (uint32_t)k * kstep), idesc, acc0 | (uint32_t)(k > 0));
                    um::mma2_ts_acc(d, a_lo + 8u * (uint32_t)k, um::desc64(bh + (uint32_t)k * kstep), idesc);
                    um::mma2_ts_acc(d, a_hi + 8u * (uint32_t)k, um::desc64(bl + (uint32_t)k * kstep), idesc);
                  }
                  break;
              }
              if (c == nch - 1) um::commit2(bar(B3_L2_FULL + p));   // part p is handed over before part p+1 is issued
            }
            __syncwarp();
          }
          if (ok && issuer) {
            um::commit2(bar(B3_H1_EMPTY + (int)slot));
            um::commit2(bar(B3_W_EMPTY + (int)ws));
          }
          __syncwarp();
          if (issuer) G3TR(0, tl, c * 10 + 2);          // chunk issued
          if (++slot == nslot) { slot = 0; spar ^= 1u; }
          if (++ws == wst_n) { ws = 0; wpar ^= 1u; }
        }
      }
      if (prof && issuer) {
        long long* o = P.prof + (size_t)pair * 32;
        o[0] = clock64() - t_begin; o[1] = pa_; o[2] = pb_; o[3] = pc_; o[4] = pd_; o[6] = ntiles;
      }
    }
  } else if (P.micro == 1) {
    // microbenchmark: nobody but the MMA issuer works
  } else if (warp == 1) {
    // =================================== table loader (one lane) ===================================
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");   // the PS/PA tables of the pre-pass are complete and visible
      bool ok = true;
      const int pitch_b = (KC + G3_PA_PAD) * 4;
      const uint32_t pa_tile_bytes = 32u * (uint32_t)pitch_b;
      const size_t pa_chunk_stride = (size_t)P.NT * 32 * pitch_b;   // bytes between chunks in PA
      uint32_t st = 0, spar = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        int b0, n0;
        tile_coords(tl, b0, n0);
        for (int c = 0; c < nch && ok; ++c) {
          ok = um::mbar_wait(bar(B3_PA_EMPTY + (int)st), spar ^ 1u, P.err, 52);
          if (!ok) break;
          // one stage = the chunk's PA tile (32 actions) + the chunk's slice of the 4 PS rows (4 x width floats)
          const uint32_t ps_slice = (uint32_t)P.ch.width[c] * 4u;
          const uint32_t dst = sPA + st * (uint32_t)P.pa_stage_bytes;
          if (P.micro & 2) {
            um::mbar_arrive_local(bar(B3_PA_FULL + (int)st));
          } else {
            um::mbar_expect_tx(bar(B3_PA_FULL + (int)st), pa_tile_bytes + 4u * ps_slice);
            um::bulk_g2s(dst, reinterpret_cast<const unsigned char*>(P.pa) + (size_t)c * pa_chunk_stride + (size_t)n0 * pitch_b,
                         pa_tile_bytes, bar(B3_PA_FULL + (int)st));
            for (int i = 0; i < 4; ++i) {
              const int b = (b0 + i < P.B) ? b0 + i : P.B - 1;
              um::bulk_g2s(dst + pa_tile_bytes + (uint32_t)i * (uint32_t)(KC * 4),
                           reinterpret_cast<const unsigned char*>(P.ps) + ((size_t)b * P.H1P + (size_t)P.ch.start[c]) * 4,
                           ps_slice, bar(B3_PA_FULL + (int)st));
            }
          }
          if (++st == past_n) { st = 0; spar ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ============ weight loader (one lane): this CTA's half of chunk c, hi then lo, every tile ============
    if (lane == 0) {
      bool ok = true;
      const unsigned char* whi = P.blob[rank] + P.off_hi;
      const unsigned char* wlo = P.blob[rank] + P.off_lo;
      uint32_t ws = 0, wpar = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        for (int c = 0; c < nch && ok; ++c) {
          ok = um::mbar_wait(bar(B3_W_EMPTY + (int)ws), wpar ^ 1u, P.err, 53);
          if (!ok) break;
          const uint32_t bytes = (uint32_t)(P.ch.width[c] >> 3) * (uint32_t)P.lbo;
          const size_t src = (size_t)(P.ch.start[c] >> 3) * (size_t)P.lbo;
          const uint32_t dst = sW + ws * (uint32_t)P.w_stage_bytes;
          if (P.micro & 2) {
            um::mbar_arrive_local(bar(B3_W_FULL + (int)ws));
          } else if (MODE == G3_C8) {
            // [fp16 hi chunk unless resident] [e4m3 hi8 chunk] [e4m3 lo8 chunk]; the 8-bit regions follow each other in the blob
            const uint32_t b8 = bytes >> 1;
            const size_t src8 = src >> 1, sz8 = (size_t)(P.H1P >> 4) * (size_t)P.lbo;
            const uint32_t d8 = dst + (P.resident_hi ? 0u : (uint32_t)P.w_half_bytes);
            um::mbar_expect_tx(bar(B3_W_FULL + (int)ws), (P.resident_hi ? 0u : bytes) + 2u * b8);
            if (!P.resident_hi) um::bulk_g2s(dst, whi + src, bytes, bar(B3_W_FULL + (int)ws));
            um::bulk_g2s(d8, wlo + src8, b8, bar(B3_W_FULL + (int)ws));
            um::bulk_g2s(d8 + (uint32_t)P.w8_bytes, wlo + sz8 + src8, b8, bar(B3_W_FULL + (int)ws));
          } else if (P.resident_hi) {
            um::mbar_expect_tx(bar(B3_W_FULL + (int)ws), bytes);
            um::bulk_g2s(dst, wlo + src, bytes, bar(B3_W_FULL + (int)ws));
          } else {
            um::mbar_expect_tx(bar(B3_W_FULL + (int)ws), 2u * bytes);
            um::bulk_g2s(dst, whi + src, bytes, bar(B3_W_FULL + (int)ws));
            um::bulk_g2s(dst + (uint32_t)P.w_half_bytes, wlo + src, bytes, bar(B3_W_FULL + (int)ws));
          }
          if (++ws == wst_n) { ws = 0; wpar ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ============ relay (one lane): my weight stage has landed -> tell the leader's MMA thread ============
    if (lane == 0) {
      bool ok = true;
      const uint32_t ready0 = um::mapa(bar(B3_W_READY), 0);
      uint32_t ws = 0, wpar = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        for (int c = 0; c < nch && ok; ++c) {
          ok = um::mbar_wait(bar(B3_W_FULL + (int)ws), wpar, P.err, 54);
          if (!ok) break;
          um::mbar_arrive_cluster(ready0 + 8u * ws);
          if (++ws == wst_n) { ws = 0; wpar ^= 1u; }
        }
      }
    } else if (PROF && prof && lane == 1 && pair == 0 && rank == 0) {
      // cycle-accounting monitor: completion time of every chunk's MMAs (tcgen05.commit on H1_EMPTY) for the first tiles
      bool ok = true;
      uint32_t slot = 0, spar = 0;
      int n = 0;
      for (int tl = 0; tl < ntiles && ok && n < 480; ++tl) {
        for (int c = 0; c < nch && ok && n < 480; ++c) {
          ok = um::mbar_wait(bar(B3_H1_EMPTY + (int)slot), spar, P.err, 55);
          P.prof[3072 + n++] = clock64() - t_begin;
          G3TR(3, tl, c * 10 + 9);                        // chunk's MMAs complete
          if (++slot == nslot) { slot = 0; spar ^= 1u; }
        }
      }
    }
  } else if (warp < G3_W2_0) {
    // ===== epilogue 1: h = relu(PS[b] + PA[n]) in fp32 -> fp16 hi | lo -> TMEM slot (layer 2's A operands) =====
    const int cg = (warp - 4) >> 2;                // this warp's index among the builders of its lane quarter
    bool ok = true;
    uint32_t st = 0, ppar = 0, slot = 0, spar = 0;
    for (int tl = 0; tl < ntiles && ok; ++tl)
      for (int c = 0; c < nch && ok; ++c) ok = build_chunk(tl, c, cg, G3_E1W / 4, st, ppar, slot, spar);
    if (prof && rank == 0 && tid == 128) {
      long long* o = P.prof + (size_t)pair * 32 + 8; o[0] = pa_; o[1] = pb_; o[2] = pc_; o[3] = pd_;
    }
  } else {
    // ================ epilogue 2 (8 warps): L2 acc -> relu -> signed sum -> q (arithmetic of K1-grid) ===============
    const int q4 = warp & 3, cg = (warp - G3_W2_0) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t l2e0 = um::mapa(bar(B3_L2_EMPTY), 0);
    const int npos = __ldg(reinterpret_cast<const int*>(P.blob[0] + P.off_c0));
    const float inv_scale = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 2);
    const float b3v = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 3);
    const int rloc = q4 * 32 + lane;
    const int np = P.parts.np;
    bool ok = true;
    // fused reduction state of this thread's (state, lane) across the state group's action blocks
    constexpr int FAN = FA > 0 ? FA : 1;
    float fm = -CUDART_INF_F, fz = 0.f, facc = 0.f, fgm[FAN], fgl[FAN];
    float cmu = 0.f, ch = 0.f, cgs = 0.f, cc0 = 0.f, cvb = 0.f;
#pragma unroll
    for (int k = 0; k < FAN; ++k) { fgm[k] = 0.f; fgl[k] = 0.f; }
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int p = 0; p < np; ++p) {
        const int hb = P.parts.base[p], hn = P.parts.base[p + 1] - hb;
        long long t0 = G3T();
        ok = um::mbar_wait(bar(B3_L2_FULL + p), (uint32_t)(tl & 1), P.err, 41);
        if (!ok) break;
        um::tc_fence_after();
        long long t1 = G3T();
        pa_ += t1 - t0;
        if (tid == 32 * G3_W2_0) G3TR(2, tl, p * 10 + 0);    // part full
        // my columns of this part: units of 8 split over the column groups (<= G3_VR columns, guaranteed by the plan)
        const int units = hn >> 3, u0 = units * cg / G3_E2G, u1 = units * (cg + 1) / G3_E2G;
        const int jb = hb + u0 * 8, w = (u1 - u0) * 8;
        uint32_t v[G3_VR];
        const uint32_t t0c = lane_addr + (uint32_t)jb;
        int nu = (P.micro & 8) ? 0 : (w >> 3);
        if constexpr (G3_VR >= 64) {
          if (nu == 8) { um::tmem_ld32p(t0c, v); um::tmem_ld32p(t0c + 32, v + 32); nu = 0; }
          else if (nu == 7) { um::tmem_ld32p(t0c, v); um::tmem_ld16p(t0c + 32, v + 32); um::tmem_ld8p(t0c + 48, v + 48); nu = 0; }
        }
        switch (nu) {                                                        // static register indices for every width
          case 6: um::tmem_ld32p(t0c, v); um::tmem_ld16p(t0c + 32, v + 32); break;
          case 5: um::tmem_ld32p(t0c, v); um::tmem_ld8p(t0c + 32, v + 32); break;
          case 4: um::tmem_ld32p(t0c, v); break;
          case 3: um::tmem_ld16p(t0c, v); um::tmem_ld8p(t0c + 16, v + 16); break;
          case 2: um::tmem_ld16p(t0c, v); break;
          case 1: um::tmem_ld8p(t0c, v); break;
          default: break;
        }
        um::tmem_ld_wait();
        um::tc_fence_before();
        __syncwarp();
        if (lane == 0) um::mbar_arrive_cluster(l2e0 + 8u * (uint32_t)p);     // part p handed back
        { long long tt = G3T(); pb_ += tt - t1; t1 = tt; }
        if (tid == 32 * G3_W2_0) G3TR(2, tl, p * 10 + 1);    // loaded + released
        if (w > 0 && !(P.micro & 8)) relu_signed_round8<G3_VR>(v, w, npos - jb, a0, a1, a2, a3);
        { long long tt = G3T(); pc_ += tt - t1; }
        if (tid == 32 * G3_W2_0) G3TR(2, tl, p * 10 + 2);    // math done
      }
      if (!ok) break;
      const float acc = (a0 + a1) + (a2 + a3);
      float* qp = qpart + (tl & 1) * (128 * (G3_E2G - 1));
      if (cg) qp[(cg - 1) * 128 + rloc] = acc;
      um::named_bar_sync(1, 128 * G3_E2G);
      if (!cg) {
        int b0, n0;
        const bool valid = tile_coords(tl, b0, n0);
        const int b = b0 + q4, n = n0 + lane;
        float tot = acc;
#pragma unroll
        for (int g = 0; g < G3_E2G - 1; ++g) tot += qp[g * 128 + rloc];
        const float qv = fmaf(0.5f * inv_scale, tot, b3v);       // tot = 2 x the signed sum
        const bool live = valid && b < P.B && n < P.N;
        if (live && P.q) P.q[(size_t)b * P.N + n] = qv;
        if (FA > 0 && P.fuse) {
          // ---- fused per-state reduction (the arithmetic of k_policy_reduce, reduce.cu; forwardkl_network.py:165-194 with
          // get_logprob :324-351 in place; ReverseKL reversekl_network.py:181-203).  This thread sees q(s_b, a_n) for its
          // state and n = lane, lane + 32, ...: ForwardKL keeps an ONLINE softmax (running max, rescaled sums).
          const int A = P.A_pol, nb = P.state_major ? (tl % P.NT) : 0;
          if (nb == 0) {                                   // new state group: per-state constants (lane k holds dimension k)
            fz = 0.f; facc = 0.f; fm = -CUDART_INF_F;
#pragma unroll
            for (int k = 0; k < FAN; ++k) { fgm[k] = 0.f; fgl[k] = 0.f; }
            const int bb = b < P.B ? b : P.B - 1;
            float ls = 0.f;
            if (lane < A) {
              ls = __ldg(P.flstd + (size_t)bb * A + lane);
              const float sd = expf(ls);
              cmu = __ldg(P.fmean + (size_t)bb * A + lane);
              ch = (A == 1) ? 0.5f / (sd * sd) : 0.5f / sd;
              cgs = (A == 1) ? 1.f / (sd * sd) : 0.5f / sd;
            }
            float c = (A == 1) ? -ls : -0.5f * ls;                 // A == 1: Normal(mean, std); A > 1: std as covariance (:350)
            c = (lane < A) ? c : 0.f;
            c = warp_sum(c);
            cc0 = c - 0.9189385332046727f * (float)A;
            cvb = (P.fuse == 2) ? __ldg(P.fv + bb) : 0.f;
          }
          float lp = cc0, wn = 0.f, d[FAN];
          if (n < P.N) { lp -= __ldg(P.fJ + n); wn = __ldg(P.fw + n); }
#pragma unroll
          for (int k = 0; k < FAN; ++k) {
            const float mu_k = __shfl_sync(0xffffffffu, cmu, k), h_k = __shfl_sync(0xffffffffu, ch, k);
            d[k] = 0.f;
            if (k < A && n < P.N) {
              d[k] = __ldg(P.fU + (size_t)k * P.N + n) - mu_k;
              lp = fmaf(-h_k * d[k], d[k], lp);
            }
          }
          float g = 0.f;                                  // weight of this (state, action) pair in the gradient sums
          if (P.fuse == 1) {
            const float x = qv * (1.f / P.alpha);
            const float mn = live ? fmaxf(fm, x) : fm;
            const float sc = (mn == fm) ? 1.f : expf(fm - mn);     // rescale of the running sums (exp(-inf) = 0 the first time)
            const float e = live ? expf(x - mn) * wn : 0.f;
            fm = mn;
            fz = fmaf(fz, sc, e);
            facc = fmaf(facc, sc, e * lp);
#pragma unroll
            for (int k = 0; k < FAN; ++k) { fgm[k] *= sc; fgl[k] *= sc; }
            g = e;
          } else if (live) {
            const float pe = expf(lp), inner = (qv - cvb) - P.alpha * lp;
            facc = fmaf(-pe * inner, wn, facc);
            g = (-pe * (inner - P.alpha)) * wn;
          }
#pragma unroll
          for (int k = 0; k < FAN; ++k) {
            const float h_k = __shfl_sync(0xffffffffu, ch, k), gs_k = __shfl_sync(0xffffffffu, cgs, k);
            if (k < A) {
              fgm[k] = fmaf(g, 2.f * h_k * d[k], fgm[k]);
              fgl[k] = fmaf(g, (A == 1) ? (gs_k * d[k] * d[k] - 1.f) : (gs_k * d[k] * d[k] - 0.5f), fgl[k]);
            }
          }
          if (nb == P.NT - 1 || !P.state_major) {                  // the state's last action block: combine the 32 lanes
            float sc = 1.f, zt = 1.f;
            if (P.fuse == 1) {
              const float M = warp_max(fm);
              sc = (fm == -CUDART_INF_F) ? 0.f : expf(fm - M);
              zt = warp_sum(fz * sc);
            }
            const float at = warp_sum(facc * sc);
            const float post = (P.fuse == 1) ? -1.f / zt : 1.f;    // FKL: loss = -sum p lp, g = -p / B_total
            if (lane == 0 && valid && b < P.B) P.loss_b[b] = (P.fuse == 1) ? -at / zt : at;
#pragma unroll
            for (int k = 0; k < FAN; ++k) {
              if (k < A) {
                const float gm_t = warp_sum(fgm[k] * sc) * post * P.inv_btotal, gl_t = warp_sum(fgl[k] * sc) * post * P.inv_btotal;
                if (lane == 0 && valid && b < P.B) {
                  if (P.dmean) P.dmean[(size_t)b * A + k] = gm_t;
                  if (P.dlstd) P.dlstd[(size_t)b * A + k] = gl_t;
                }
              }
            }
          }
        }
      }
    }
    if (prof && rank == 0 && tid == 32 * G3_W2_0) {
      long long* o = P.prof + (size_t)pair * 32 + 16; o[0] = pa_; o[1] = pb_; o[2] = pc_;
    }
  }
#undef G3T
#undef G3TR

  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  if (warp == 0) um::tmem_dealloc2(tmem_base, 512);
}

// host side ------------------------------------------------------------------------------------
struct Grid3Plan {
  int KC, nslot, w_stages, pa_stages;
  Grid3Chunks ch;
  Grid3Parts parts;
  int resident_hi, sm_whi, whi_bytes;
  int sm_w, sm_pa, sm_ps, sm_qp, sm_bar, w_stage, w_half, pa_stage, ps_stage, total;
};

// Column granularity of the C8 accumulator parts.  The PTX shape table allows N in steps of 16 at cta_group::2 for
// kind::f8f6f4 as for kind::f16 (the multiple-of-32 rule is a CuTe static_assert, not the hardware's): 300 -> 304 instead
// of 320 columns, 5 % less tensor-pipe, accumulator-drain and epilogue-2 work.  RLC_G3_C8_N16=0 restores steps of 32.
#ifndef RLC_G3_C8_N16_DEFAULT
#define RLC_G3_C8_N16_DEFAULT 1
#endif
static int g3_c8_unit() {
  static int u = 0;
  if (!u) { const char* e = getenv("RLC_G3_C8_N16"); u = (e ? e[0] == '1' : RLC_G3_C8_N16_DEFAULT) ? 16 : 32; }
  return u;
}

static bool make_geom3(const rlc_critic* c, PackGeom& G, int mode = G3_X3) {
  if (!make_geom(c, G, 1)) return false;
  if (mode == G3_C8) {          // kind::f8f6f4: K = 32 per MMA
    G.H1P = (c->H1 + 1 + 31) & ~31;
    G.H2P = (c->H2 + g3_c8_unit() - 1) & ~(g3_c8_unit() - 1);
    if (G.H2P > 480) return false;
  }
  const int sz = (G.H1P / 8) * (G.H2P / 2) * 16;
  G.CH = 0; G.nch = 0;                 // the split pack does not depend on the chunking (no W1 region)
  G.off_w2 = 0;                        // hi
  G.off_w1 = sz;                       // lo
  G.off_w3 = 2 * sz;                   // int inv[H2P] (k_pack_head)
  G.off_nb2 = G.off_w3 + G.H2P * 4;
  G.off_c0 = G.off_nb2;
  G.blob_bytes = G.off_c0 + 16;
  return true;
}

// Accumulator column parts: as few as possible, widths multiples of 16 as even as possible (304 -> 160 + 144).
static bool make_parts3(int H2P, Grid3Parts& pt, int mode = G3_X3) {
  memset(&pt, 0, sizeof(pt));
  const int U = mode == G3_C8 ? g3_c8_unit() : 16;  // part widths: multiples of 16 (see g3_c8_unit)
  if (H2P % U) return false;
  const int units = H2P / U;
  int np = (H2P + 191) / 192;                     // fewest parts an epilogue-2 thread can hold (wider MMAs are more efficient)
  {
    const char* e = getenv("RLC_G3_NP");          // tuning knob: number of accumulator parts
    if (e) { const int v = atoi(e); if (v >= 1 && v <= G3_MAX_NP) np = v; }
  }
  if (np > units) np = units;
  // more parts when an epilogue-2 thread could not hold its share of the widest one
  while (np < units && np < G3_MAX_NP && U * ((units + np - 1) / np) > g3_e2g(mode) * g3_vr(mode)) ++np;
  if (np < 1 || np > G3_MAX_NP) return false;
  int b = 0;
  for (int p = 0; p < np; ++p) {
    const int w = U * (units / np + (p < units % np ? 1 : 0));
    if (w > g3_e2g(mode) * g3_vr(mode) || w > 256) return false;
    pt.base[p] = b;
    b += w;
  }
  pt.base[np] = b;
  pt.np = np;
  return b == H2P;
}

static bool plan_grid3(const PackGeom& G, size_t smem_limit, Grid3Plan& p, int mode = G3_X3) {
  memset(&p, 0, sizeof(p));
  if (!make_parts3(G.H2P, p.parts, mode)) return false;
  const int CU = mode == G3_C8 ? 32 : 16;           // chunk widths: multiples of one MMA's K
  int kc = ((512 - G.H2P) / 2) & ~(CU - 1);
  if (kc > 96) kc = 96;
  {
    const char* e = getenv("RLC_G3_KC");          // tuning knob: chunk width
    if (e) { const int v = atoi(e); if (v >= CU && v <= kc && (v % CU) == 0) kc = v; }
  }
  if (kc > G.H1P) kc = G.H1P;
  if (kc < CU) return false;
  p.KC = kc;
  p.nslot = (512 - G.H2P) / kc;
  if (p.nslot > G3_MAX_SLOT) p.nslot = G3_MAX_SLOT;
  if (p.nslot < 2) return false;
  // chunks: ceil(H1P / KC) of them, the 16-feature units spread evenly (416 -> 96 + 4 x 80): no chunk's MMAs are
  // much shorter than the time epilogue 1 needs to build the next one
  const int units = G.H1P / CU, n = (G.H1P + kc - 1) / kc;
  if (n > G3_MAX_NCH) return false;
  // (wide chunks first or last makes no measurable difference: 1.448 vs 1.449 ms)
  for (int i = 0, f = 0; i < n; ++i) {
    const int w = CU * (units / n + (i < units % n ? 1 : 0));
    p.ch.start[i] = f; p.ch.width[i] = w; f += w;
  }
  p.ch.nch = n;
  const int lbo = (G.H2P / 2) * 16;
  p.w_half = (kc / 8) * lbo;
  p.pa_stage = 32 * (kc + G3_PA_PAD) * 4 + 4 * kc * 4;     // PA tile + the chunk's slice of the tile's 4 PS rows
  p.ps_stage = 0;
  const int tail = 128 * 4 * 2 * (g3_e2g(mode) - 1) + B3_COUNT * 8 + 16 + 1024;
  p.whi_bytes = (G.H1P / 8) * lbo;
  // Preferred: W_hi resident, only W_lo streams (half the bulk-copy writes into shared memory, whose bandwidth the
  // tensor core's operand reads need); otherwise both stream.  Stage counts: as deep as fits, weights first.
  static const int combos[4][2] = {{3, 3}, {3, 2}, {2, 3}, {2, 2}};   // {weight stages, PA stages}
  int mode_res = 1;
  {
    const char* e = getenv("RLC_G3_RESIDENT");      // tuning knob: 0 = stream both halves of the split weights
    if (e && e[0] == '0') mode_res = 0;
  }
  bool found = false;
  for (int res = mode_res; res >= 0 && !found; --res) {
    // per stage: X3 = [fp16 hi unless resident | fp16 lo];  C8 = [fp16 hi unless resident | e4m3 hi8 | e4m3 lo8] (2 x w_half / 2)
    const int wst = res ? p.w_half : 2 * p.w_half;
    for (int i = 0; i < 4 && !found; ++i) {
      long long need = (long long)(res ? p.whi_bytes : 0) + (long long)combos[i][0] * wst +
                       (long long)combos[i][1] * p.pa_stage;
      need = ((need + 127) & ~127ll) + tail;           // exactly the carve below
      if (need <= (long long)smem_limit) {
        p.resident_hi = res; p.w_stage = wst; p.w_stages = combos[i][0]; p.pa_stages = combos[i][1];
        found = true;
      }
    }
  }
  if (!found) return false;
  p.sm_whi = 0;
  p.sm_w = p.resident_hi ? p.whi_bytes : 0;
  p.sm_pa = p.sm_w + p.w_stages * p.w_stage;
  p.sm_ps = 0;
  p.sm_qp = (p.sm_pa + p.pa_stages * p.pa_stage + 127) & ~127;
  p.sm_bar = p.sm_qp + 2 * 128 * 4 * (g3_e2g(mode) - 1);
  p.total = p.sm_bar + B3_COUNT * 8 + 16 + 1024;
  return (size_t)p.total <= smem_limit;
}

static int launch_pack_x3(rlc_handle* h, const float* theta, const PackGeom& G, int mode, unsigned char* b0,
                          unsigned char* b1, cudaStream_t st) {
  Grid3Parts pt;
  if (!make_parts3(G.H2P, pt, mode)) return RLC_ERR_UNSUPPORTED;
  const long long n3 = (long long)G.H1P * G.H2P;
  k_pack_x3<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(theta, G, pt, mode, b0, b1);
  (void)h;
  return RLC_OK;
}

static inline int g3_mode_of(int prec) { return prec == RLC_PREC_FP16C8 ? G3_C8 : G3_X3; }

static int rlc_eval_umma_grid3(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                               int prec, float* q_out, cudaStream_t st, const rlc_fuse_args* fuse);

int rlc_eval_umma_grid3_fused(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N, int prec,
                              float* q_out, const rlc_fuse_args* f, cudaStream_t st) {
  if (h->sm_major != 10 || c->topology != RLC_TIN || (prec != RLC_PREC_FP16X3 && prec != RLC_PREC_FP16C8)) return RLC_ERR_UNSUPPORTED;
  return rlc_eval_umma_grid3(h, c, s, B, a, N, prec, q_out, st, f);
}

bool rlc_umma3_supported(const rlc_handle* h, const rlc_critic* c, int prec) {
  if (h->sm_major != 10 || c->topology != RLC_TIN) return false;
  PackGeom G;
  Grid3Plan gp;
  const int mode = g3_mode_of(prec);
  return make_geom3(c, G, mode) && plan_grid3(G, h->smem_optin, gp, mode);
}

static int rlc_eval_umma_grid3(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                               int prec, float* q_out, cudaStream_t st, const rlc_fuse_args* fuse = nullptr) {
  PackGeom G;
  Grid3Plan gp;
  const int mode = g3_mode_of(prec);
  if (!make_geom3(c, G, mode) || !plan_grid3(G, h->smem_optin, gp, mode)) return RLC_ERR_UNSUPPORTED;
  if ((long long)B * N >= (1ll << 31)) return RLC_ERR_UNSUPPORTED;
  // the fused reduction needs the state-major tile order: every CTA must own at least two whole state groups, or the
  // persistent grid would idle (small minibatches keep the round-robin order and the separate reduction kernel)
  if (fuse && ((B + 3) / 4 < 2 * h->num_sms || fuse->A < 1 || fuse->A > POL_MAX_A || fuse->A != c->A)) return RLC_ERR_UNSUPPORTED;
  rlc_pack* pk = nullptr;
  int rc = get_pack(h, c, prec, G, st, &pk);
  if (rc) return rc;
  const int NT = (N + 31) / 32;
  const int pitch = gp.KC + G3_PA_PAD;
  const size_t nps = ((size_t)B * G.H1P + 63) & ~(size_t)63, npa = ((size_t)gp.ch.nch * NT * 32 * pitch + 63) & ~(size_t)63;
  const size_t nuj = fuse ? (size_t)N * (fuse->A + 1) : 0;
  void* ws = nullptr;
  rc = rlc_workspace(h, (nps + npa + nuj) * sizeof(float) + 256, &ws);
  if (rc) return rc;
  float* PS = (float*)ws;
  float* PA = PS + nps;                                        // 256-byte aligned
  float* UJ = PA + npa;
  if (fuse) {
    rc = rlc_launch_grid_logterms(h, fuse->grid, N, fuse->A, fuse->action_scale, UJ, UJ + (size_t)N * fuse->A, st);
    if (rc) return rc;
  }
  {
    const int sg = (B + GR_PRE_ROWS - 1) / GR_PRE_ROWS, ag = (NT * 32 + GR_PRE_ROWS - 1) / GR_PRE_ROWS;
    const dim3 blocks((unsigned)(sg + ag), (unsigned)((G.H1P + 127) / 128));
    const size_t pre_smem = (size_t)GR_PRE_ROWS * (c->S > c->A ? c->S : c->A) * sizeof(float);
    k_grid3_parts<<<blocks, 128, pre_smem, st>>>(c->theta, s, a, c->smin, c->smax, B, N, c->S, c->A, c->H1, c->H2,
                                                  G.H1P, gp.KC, gp.ch, NT, sg, PS, PA, h->err_flag,
                                                  mode == G3_C8 ? G3_RANGE_LIMIT_C8 : G3_RANGE_LIMIT);
    RLC_LAUNCH_CHECK(h);
  }
  Grid3Params P;
  memset(&P, 0, sizeof(P));
  P.q = q_out; P.B = B; P.N = N;
  P.H1P = G.H1P; P.H2P = G.H2P; P.parts = gp.parts;
  P.KC = gp.KC; P.nslot = gp.nslot; P.ch = gp.ch; P.NT = NT;
  P.num_cta_tiles = (long long)((B + 3) / 4) * NT;
  P.num_pair_tiles = (int)((P.num_cta_tiles + 1) / 2);
  P.ps = PS; P.pa = PA;
  P.blob[0] = (const unsigned char*)pk->dev;
  P.blob[1] = P.blob[0] + G.blob_bytes;
  P.off_hi = G.off_w2; P.off_lo = G.off_w1; P.off_c0 = G.off_c0;
  P.lbo = (G.H2P / 2) * 16;
  P.w8_bytes = gp.w_half / 2;
  P.w_stages = gp.w_stages; P.pa_stages = gp.pa_stages;
  P.resident_hi = gp.resident_hi; P.sm_whi = gp.sm_whi; P.whi_bytes = gp.whi_bytes;
  P.sm_w = gp.sm_w; P.sm_pa = gp.sm_pa; P.sm_ps = gp.sm_ps; P.sm_qp = gp.sm_qp; P.sm_bar = gp.sm_bar;
  P.w_stage_bytes = gp.w_stage; P.w_half_bytes = gp.w_half; P.pa_stage_bytes = gp.pa_stage; P.ps_stage_bytes = gp.ps_stage;
  P.err = h->err_flag;
  P.NSG = (B + 3) / 4;
  if (fuse) {
    P.fuse = fuse->mode; P.state_major = 1; P.A_pol = fuse->A;
    P.fw = fuse->w; P.fU = UJ; P.fJ = UJ + (size_t)N * fuse->A; P.fmean = fuse->mean; P.flstd = fuse->log_std; P.fv = fuse->v;
    P.alpha = fuse->alpha; P.inv_btotal = 1.f / (float)fuse->B_total;
    P.loss_b = fuse->loss_b; P.dmean = fuse->dmean; P.dlstd = fuse->dlog_std;
  }

  int pairs = h->num_sms / 2;
  if (pairs > P.num_pair_tiles) pairs = P.num_pair_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(g3_threads(mode));
  cfg.dynamicSmemBytes = (size_t)gp.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap our prologue with k_grid3_parts
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("RLC_G3_PDL"); pdl = e ? (e[0] == '1' ? 1 : 0) : RLC_G3_PDL_DEFAULT; }
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  static int prof_on = -1, micro = -1;
  static long long* prof_dev = nullptr;
  if (prof_on < 0) { const char* e = getenv("RLC_UMMA_PROF"); prof_on = (e && e[0] == '1') ? 1 : 0; }
  if (micro < 0) { const char* e = getenv("RLC_UMMA_MICRO"); micro = e ? atoi(e) : 0; }
  P.micro = micro;
  if (prof_on) {
    if (!prof_dev) RLC_CUDA(cudaMalloc(&prof_dev, (4096 + 4 * 2048) * sizeof(long long)));
    RLC_CUDA(cudaMemsetAsync(prof_dev, 0, (4096 + 4 * 2048) * sizeof(long long), st));
    P.prof = prof_dev;
  }
  void (*kern)(const Grid3Params) = nullptr;
  if (fuse) {
    const int fa = fuse->A <= 1 ? 1 : fuse->A <= 2 ? 2 : fuse->A <= 4 ? 4 : fuse->A <= 6 ? 6 : 8;
    P.prof = nullptr;
#define G3_FUSED(M_)                                                                                            \
    (fa == 1 ? k_critic_umma_grid3<false, M_, 1> : fa == 2 ? k_critic_umma_grid3<false, M_, 2>                   \
     : fa == 4 ? k_critic_umma_grid3<false, M_, 4> : fa == 6 ? k_critic_umma_grid3<false, M_, 6>                 \
                                                             : k_critic_umma_grid3<false, M_, 8>)
    kern = mode == G3_C8 ? G3_FUSED(G3_C8) : G3_FUSED(G3_X3);
#undef G3_FUSED
  } else {
    kern = mode == G3_C8 ? (P.prof ? k_critic_umma_grid3<true, G3_C8, 0> : k_critic_umma_grid3<false, G3_C8, 0>)
                         : (P.prof ? k_critic_umma_grid3<true, G3_X3, 0> : k_critic_umma_grid3<false, G3_X3, 0>);
  }
  RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gp.total));
  RLC_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  RLC_LAUNCH_CHECK(h);
  if (prof_on) {   // debug only: synchronises
    static long long hp[4096 + 4 * 2048];   // [0,3072): per-pair accounting; [3072,3552): chunk completions; [4096,..): event trace
    RLC_CUDA(cudaStreamSynchronize(st));
    RLC_CUDA(cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost));
    const long long* o = hp;
    const double T = (double)o[0], nt = (double)(o[6] > 0 ? o[6] : 1);
    fprintf(stderr, "[grid3 prof pair0] KC %d slots %d wst %d%s past %d tiles %lld total %.0f cyc (%.0f/tile) | MMA: waitH1 %.1f%% waitW %.1f%% "
            "waitL2E A %.1f%% B %.1f%% | ep1: waitPA %.1f%% waitH1E %.1f%% build+st %.1f%% wait::st+arrive %.1f%% | ep2: waitL2F %.1f%% ld %.1f%% math %.1f%%\n",
            gp.KC, gp.nslot, gp.w_stages, gp.resident_hi ? " (lo only)" : "", gp.pa_stages, o[6], T, T / nt, 100 * o[1] / T, 100 * o[2] / T, 100 * o[3] / T, 100 * o[4] / T,
            100 * o[8] / T, 100 * o[9] / T, 100 * o[10] / T, 100 * o[11] / T, 100 * o[16] / T, 100 * o[17] / T, 100 * o[18] / T);
    if (getenv("RLC_UMMA_TRACE") && atoi(getenv("RLC_UMMA_TRACE")) >= 2) {
      const char* names[4] = {"MMA", "EP1", "EP2", "DONE"};
      for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 1000; ++i) {
          const long long code = hp[4096 + r * 2048 + 2 * i], tc = hp[4096 + r * 2048 + 2 * i + 1];
          if (tc == 0 && code == 0) break;
          fprintf(stderr, "TRACE %lld %s %lld\n", tc, names[r], code);
        }
    }
    if (getenv("RLC_UMMA_TRACE")) {
      const int nc = gp.ch.nch;
      for (int tl = 8; tl < 16; ++tl) {
        fprintf(stderr, "[grid3 trace] tile %d chunk-completion deltas:", tl);
        for (int c = 0; c < nc; ++c) fprintf(stderr, " %lld", hp[3072 + tl * nc + c] - hp[3072 + tl * nc + c - 1]);
        fprintf(stderr, "  (tile %lld)\n", hp[3072 + tl * nc + nc - 1] - hp[3072 + (tl - 1) * nc + nc - 1]);
      }
    }
  }
  return RLC_OK;
}
