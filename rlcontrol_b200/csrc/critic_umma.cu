// K1: fused T-in critic evaluation on 5th-gen tensor cores (sm_100a).
//
//   q[b,n] = w3 . relu(W2 relu(W1 [s_b ; a_n ; 1]) + b2) + b3      for B*N rows, state-major
//
// One persistent CTA *pair* (cta_group::2, 256 rows per pair-tile) per two SMs.  Per CTA:
//   * fp16/bf16 weights pre-packed in the UMMA no-swizzle K-major core-matrix layout stay resident
//     in shared memory for the whole kernel (W2 is split by output feature across the pair, so
//     the 400x300 layer fits: 121.6 KB per CTA);
//   * the input tile X = [s ; a ; 1] (bias folded as a ones column) is assembled by 3 producer
//     warps straight from s[B,S] and a[N,A] / a[B,N,A] -- the stacked [B*N,S+A] tensor of
//     forwardkl_network.py:160-164 never exists;
//   * layer 1 runs as tcgen05.mma (M=256,K=K1P) into a double-buffered TMEM accumulator, in
//     N-chunks of <=80 columns; 4 epilogue warps pull each chunk with tcgen05.ld, apply ReLU,
//     convert to fp16 and store it as the next K-slice of layer 2's A operand in a 3-stage
//     shared-memory ring -- activations never touch HBM;
//   * layer 2 accumulates [256 x H2P] in TMEM over the ring; 4 more warps drain it with
//     tcgen05.ld and fuse bias + ReLU + the w3 dot, writing one float per row.
// Synchronisation is mbarrier-only (tcgen05.commit multicast to both CTAs; consumer arrivals
// are sent to the leader CTA's barriers through shared::cluster addresses).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <math_constants.h>

#include "common.cuh"

#define UM_THREADS 384
#define UM_NST 3           // H1 ring stages
#define UM_WAIT_LIMIT (1u << 26)

struct UmmaParams {
  // problem
  const float* s;
  const float* a;
  const float* smin;
  const float* smax;
  float* q;
  long long R;  // B*N
  int N, S, A, act_per_state;
  // network (padded)
  int K1P, KC1;        // layer-1 K padded to 16, number of 8-wide K chunks
  int H1P, KC2;        // layer-2 K (=H1 padded to 16), number of 8-wide K chunks
  int H2P;             // layer-2 N padded to 16
  int NA, NB;          // layer-2 N split (NA + NB = H2P), both % 16 == 0, NB may be 0
  int nch;             // number of layer-1 chunks: chunk c covers features [c*CH, min((c+1)*CH, H1P))
  int CH;              // chunk width (multiple of 16) == TMEM columns per L1 buffer
  int nb1;             // TS variant: number of L1 chunk buffers in TMEM (2 or 3)
  // packed blob for each cta rank (device), and byte offsets inside it
  const unsigned char* blob[2];
  int off_w2, off_w1, off_w3, off_nb2, off_c0, blob_bytes;
  // smem carve (bytes from the 1024-aligned base)
  int sm_w2, sm_w1, sm_x, sm_h1, sm_par, sm_bar, x_stage_bytes, h1_stage_bytes;
  int num_pair_tiles;
  int* err;     // device error flag (0 = ok)
  long long* prof;  // optional cycle accounting (RLC_UMMA_PROF=1), 32 x int64 per pair
  const uint4* sp;  // TS variant: packed state part of X, [B][KC1] 16-byte chunks
  const uint4* ap;  // TS variant: packed action part of X, [KC1][N] (shared grid only)
};

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
namespace um {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// bounded wait: returns false (and raises the error flag) instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > UM_WAIT_LIMIT) {
      atomicCAS(err, 0, code);
      return false;
    }
    if ((spins & 0xffff) == 0 && *(volatile int*)err != 0) return false;
  }
  return true;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = 256 over the CTA pair
__device__ __forceinline__ void mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                     uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive::one on the barrier at this smem offset in BOTH CTAs once all prior MMAs completed
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t make_idesc(int fmt, int M, int N) {
  // c_format F32 (1) @4 ; a_format @7 ; b_format @10 ; a/b K-major (0) ; N>>3 @17 ; M>>4 @24
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int PREC>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (PREC == RLC_PREC_BF16) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
  } else {
    __half2 t = __floats2half2_rn(lo, hi);
    t = __hmin2(t, __half2half2(__ushort_as_half((unsigned short)0x7BFF)));  // saturate at 65504
    t = __hmax2(t, __half2half2(__ushort_as_half((unsigned short)0xFBFF)));
    return *reinterpret_cast<uint32_t*>(&t);
  }
}
template <int PREC>
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
  uint32_t r;  // one F2FP: relu, saturate to the largest finite value, round-to-nearest-even, pack
  if (PREC == RLC_PREC_BF16)
    asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace um

// ------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------
// barrier slots (8 bytes each) inside sm_bar
enum {
  BAR_X_FULL = 0,       // [2]   count 6  (leader)   3 producer warps x 2 CTAs -> MMA
  BAR_X_EMPTY = 2,      // [2]   count 1  (both)     MMA commit -> producers
  BAR_L1_FULL = 4,      // [2]   count 1  (both)     MMA commit -> ep1
  BAR_L1_EMPTY = 6,     // [2]   count 8  (leader)   ep1 -> MMA
  BAR_H1_FULL = 8,      // [NST] count 8  (leader)   ep1 -> MMA
  BAR_H1_EMPTY = 11,    // [NST] count 1  (both)     MMA commit -> ep1
  BAR_L2_FULL = 14,     // [1]   count 1  (both)     MMA commit -> ep2
  BAR_L2_EMPTY = 15,    // [2]   count 8  (leader)   ep2 -> MMA (half A, half B)
  BAR_COUNT = 17
};

namespace um {
// The shared-memory descriptor of every operand here has the same upper word: SBO = 128 B between
// 8-row groups, descriptor version 1 (sm_100), no swizzle.  Only the lower word moves.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo) {
  constexpr uint32_t hi = (128u >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}
}  // namespace um

template <int PREC>
__global__ void __launch_bounds__(UM_THREADS, 1) k_critic_umma(const UmmaParams P) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned base (same offset in both CTAs of the pair: same kernel, same carve)
  const uint32_t raw_addr = um::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = um::cta_rank();
  const uint32_t pair = um::cluster_id_x();
  const uint32_t npairs = um::num_clusters_x();

  const uint32_t sW2 = base + P.sm_w2, sW1 = base + P.sm_w1, sX = base + P.sm_x,
                 sH1 = base + P.sm_h1, sBar = base + P.sm_bar;
  const float* w3s = reinterpret_cast<const float*>(base_ptr + P.sm_par);
  const float* nb2s = w3s + P.H2P;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + P.sm_bar + BAR_COUNT * 8);
  auto bar = [&](int i) -> uint32_t { return sBar + 8u * (uint32_t)i; };

  // ---- prologue: resident weights -> smem, barriers, TMEM ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.blob[rank]);
    uint4* dW2 = reinterpret_cast<uint4*>(base_ptr + P.sm_w2);
    const int n2 = (P.off_w1 - P.off_w2) >> 4;
    for (int i = tid; i < n2; i += UM_THREADS) dW2[i] = __ldg(src + (P.off_w2 >> 4) + i);
    uint4* dW1 = reinterpret_cast<uint4*>(base_ptr + P.sm_w1);
    const int n1 = (P.off_w3 - P.off_w1) >> 4;
    for (int i = tid; i < n1; i += UM_THREADS) dW1[i] = __ldg(src + (P.off_w1 >> 4) + i);
    uint4* dP = reinterpret_cast<uint4*>(base_ptr + P.sm_par);
    const int np = (P.off_c0 - P.off_w3) >> 4;
    for (int i = tid; i < np; i += UM_THREADS) dP[i] = __ldg(src + (P.off_w3 >> 4) + i);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      um::mbar_init(bar(BAR_X_FULL + i), 6);
      um::mbar_init(bar(BAR_X_EMPTY + i), 1);
      um::mbar_init(bar(BAR_L1_FULL + i), 1);
      um::mbar_init(bar(BAR_L1_EMPTY + i), 8);
      um::mbar_init(bar(BAR_L2_EMPTY + i), 8);
    }
    for (int i = 0; i < UM_NST; ++i) {
      um::mbar_init(bar(BAR_H1_FULL + i), 8);
      um::mbar_init(bar(BAR_H1_EMPTY + i), 1);
    }
    um::mbar_init(bar(BAR_L2_FULL), 1);
    um::fence_mbar_init();
  }
  um::fence_proxy_async();  // generic-proxy weight stores -> visible to the tensor core (async proxy)
  if (warp == 0) um::tmem_alloc2(um::smem_u32(tmem_slot), 512);
  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  um::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ntiles = (P.num_pair_tiles > (int)pair)
                         ? (P.num_pair_tiles - (int)pair + (int)npairs - 1) / (int)npairs
                         : 0;
  const int nch = P.nch;
  const uint32_t L1COL = (uint32_t)P.H2P;  // TMEM column of L1 accumulator buffer 0

  if (warp == 0) {
    // =================================== MMA issuer (leader CTA) ===================================
    // One warp waits, one elected lane issues.  Everything the issue loop needs is hoisted into
    // registers (descriptor words, per-K-step increments, instruction descriptors): between two
    // tcgen05.mma there is one 32-bit add per operand.
    if (rank == 0) {
      const bool issuer = um::elect_one();
      const uint32_t fmt = (PREC == RLC_PREC_BF16) ? 1u : 0u;
      const uint32_t lbo_x = 128u * 16u;                        // X / H1 tiles: 128 rows per K chunk
      const uint32_t lbo_w1 = (uint32_t)(P.H1P / 2) * 16u;
      const uint32_t lbo_w2 = (uint32_t)(P.H2P / 2) * 16u;
      const uint32_t x_lo0 = um::desc_lo(sX, lbo_x), x_lo1 = um::desc_lo(sX + P.x_stage_bytes, lbo_x);
      const uint32_t h1_lo0 = um::desc_lo(sH1, lbo_x);
      const uint32_t h1_stage16 = (uint32_t)P.h1_stage_bytes >> 4;
      const uint32_t w1_lo = um::desc_lo(sW1, lbo_w1), w2_lo = um::desc_lo(sW2, lbo_w2);
      const uint32_t a_kstep = (2u * lbo_x) >> 4;               // one K=16 step of an A tile
      const uint32_t w1_kstep = (2u * lbo_w1) >> 4, w2_kstep = (2u * lbo_w2) >> 4;
      const uint32_t w2_half_off = (uint32_t)(P.NA / 2);        // rows of half B start here (>>4 of bytes)
      const uint32_t idA = um::make_idesc(fmt, 256, P.NA);
      const uint32_t idB = um::make_idesc(fmt, 256, P.NB > 0 ? P.NB : 16);
      const uint32_t id1_full = um::make_idesc(fmt, 256, P.CH);
      const uint32_t id1_last = um::make_idesc(fmt, 256, (P.H1P - (nch - 1) * P.CH));
      const int k1steps = P.K1P / 16;
      const int CH = P.CH, last_w = (P.H1P - (nch - 1) * P.CH);
      const bool two_halves = P.NB > 0;
      const uint32_t dL2A = tmem_base, dL2B = tmem_base + (uint32_t)P.NA;

      bool ok = true;
      // layer-2 work for the unit issued one step earlier (software pipeline of depth 1)
      uint32_t stg = 0, h1_par = 0;                             // H1 ring stage / parity of unit v
      auto issue_l2 = [&](int tlv, int cv) {
        ok = ok && um::mbar_wait(bar(BAR_H1_FULL + (int)stg), h1_par, P.err, 13);
        const int ksteps = (cv == nch - 1 ? last_w : CH) >> 4;
        const uint32_t a0 = h1_lo0 + stg * h1_stage16;
        const uint32_t b0 = w2_lo + (uint32_t)(cv * (CH >> 3)) * (lbo_w2 >> 4);
        const uint32_t acc0 = cv > 0 ? 1u : 0u;
        if (cv == 0) ok = ok && um::mbar_wait(bar(BAR_L2_EMPTY + 0), (uint32_t)((tlv & 1) ^ 1), P.err, 14);
        um::tc_fence_after();
        if (ok && issuer) {
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k)
            um::mma2(dL2A, um::desc64(a0 + (uint32_t)k * a_kstep), um::desc64(b0 + (uint32_t)k * w2_kstep),
                     idA, acc0 | (uint32_t)(k > 0));
        }
        if (two_halves) {
          if (cv == 0) {
            ok = ok && um::mbar_wait(bar(BAR_L2_EMPTY + 1), (uint32_t)((tlv & 1) ^ 1), P.err, 15);
            um::tc_fence_after();
          }
          if (ok && issuer) {
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k)
              um::mma2(dL2B, um::desc64(a0 + (uint32_t)k * a_kstep),
                       um::desc64(b0 + w2_half_off + (uint32_t)k * w2_kstep), idB, acc0 | (uint32_t)(k > 0));
          }
        }
        if (ok && issuer) {
          um::commit2(bar(BAR_H1_EMPTY + (int)stg));
          if (cv == nch - 1) um::commit2(bar(BAR_L2_FULL));
        }
        __syncwarp();
        if (++stg == UM_NST) { stg = 0; h1_par ^= 1u; }
      };

      uint32_t u = 0;
      int ptl = 0, pc = 0;
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        const int xs = tl & 1;
        for (int c = 0; c < nch && ok; ++c, ++u) {
          // ---- layer 1, chunk c of tile tl ----
          if (c == 0) ok = ok && um::mbar_wait(bar(BAR_X_FULL + xs), (uint32_t)((tl >> 1) & 1), P.err, 11);
          const uint32_t lb = u & 1u;
          ok = ok && um::mbar_wait(bar(BAR_L1_EMPTY + (int)lb), ((u >> 1) & 1u) ^ 1u, P.err, 12);
          um::tc_fence_after();
          if (ok && issuer) {
            const uint32_t d = tmem_base + L1COL + lb * (uint32_t)CH;
            const uint32_t a0 = xs ? x_lo1 : x_lo0;
            const uint32_t b0 = w1_lo + (uint32_t)(c * (CH >> 1));
            const uint32_t idesc = (c == nch - 1) ? id1_last : id1_full;
#pragma unroll 1
            for (int k = 0; k < k1steps; ++k)
              um::mma2(d, um::desc64(a0 + (uint32_t)k * a_kstep), um::desc64(b0 + (uint32_t)k * w1_kstep), idesc,
                       (uint32_t)(k > 0));
            um::commit2(bar(BAR_L1_FULL + (int)lb));
            if (c == nch - 1) um::commit2(bar(BAR_X_EMPTY + xs));
          }
          __syncwarp();
          // ---- layer 2 over the K-slice produced from the previous unit ----
          if (u > 0) issue_l2(ptl, pc);
          ptl = tl; pc = c;
        }
      }
      if (u > 0 && ok) issue_l2(ptl, pc);
    }
  } else if (warp < 4) {
    // =================================== X producers (3 warps) ===================================
    // thread <-> row: the [s_b ; a_n ; 1] row is assembled in registers and written as K1P/8
    // 16-byte core-matrix rows.  Rows past R are zero.
    const int pt = tid - 32;  // 0..95
    const uint32_t xfull_leader0 = um::mapa(bar(BAR_X_FULL + 0), 0);
    const uint32_t xfull_leader1 = um::mapa(bar(BAR_X_FULL + 1), 0);
    bool ok = true;
    const int S = P.S, K1 = P.S + P.A, KC1 = P.KC1;
    const uint32_t N = (uint32_t)P.N;
    const bool clip = P.smin != nullptr;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      const int xs = tl & 1;
      ok = um::mbar_wait(bar(BAR_X_EMPTY + xs), (uint32_t)(((tl >> 1) & 1) ^ 1), P.err, 21);
      if (!ok) break;
      const long long tile = (long long)pair + (long long)tl * npairs;
      const long long row0 = tile * 256 + (long long)rank * 128;
      unsigned char* xbase = base_ptr + P.sm_x + xs * P.x_stage_bytes;
      for (int r = pt; r < 128; r += 96) {
        const long long row = row0 + r;
        const bool live = row < P.R;
        const float* sp = P.s;
        const float* ap = P.a;
        if (live) {
          const uint32_t b = (uint32_t)row / N;   // R = B*N < 2^31 rows is enforced on the host
          sp += (size_t)b * S;
          ap += (size_t)(P.act_per_state ? (uint32_t)row : ((uint32_t)row - b * N)) * P.A;
        }
        for (int kc = 0; kc < KC1; ++kc) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = kc * 8 + e;
            float t = 0.f;
            if (live) {
              if (k < S) {
                t = __ldg(sp + k);
                if (clip) t = fminf(fmaxf(t, __ldg(P.smin + k)), __ldg(P.smax + k));
              } else if (k < K1) {
                t = __ldg(ap + (k - S));
              } else if (k == K1) {
                t = 1.f;
              }
            }
            v[e] = t;
          }
          uint4 pk;
          pk.x = um::pack2<PREC>(v[0], v[1]);
          pk.y = um::pack2<PREC>(v[2], v[3]);
          pk.z = um::pack2<PREC>(v[4], v[5]);
          pk.w = um::pack2<PREC>(v[6], v[7]);
          *reinterpret_cast<uint4*>(xbase + kc * 2048 + r * 16) = pk;
        }
      }
      um::fence_proxy_async();
      __syncwarp();
      if (lane == 0) um::mbar_arrive_cluster(xs ? xfull_leader1 : xfull_leader0);
    }
  } else if (warp < 8) {
    // =================================== epilogue 1: L1 acc -> relu -> fp16 -> H1 ring ===========
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t l1e0 = um::mapa(bar(BAR_L1_EMPTY + 0), 0), l1e1 = um::mapa(bar(BAR_L1_EMPTY + 1), 0);
    const uint32_t h1f0 = um::mapa(bar(BAR_H1_FULL), 0);   // the cluster window is linear: +8 per stage
    bool ok = true;
    uint32_t u = 0, stg = 0, h1_par = 0;
    const int CH = P.CH, last_w = (P.H1P - (nch - 1) * P.CH);
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      for (int c = 0; c < nch; ++c, ++u) {
        const uint32_t lb = u & 1u;
        ok = um::mbar_wait(bar(BAR_L1_FULL + (int)lb), (u >> 1) & 1u, P.err, 31);
        ok = ok && um::mbar_wait(bar(BAR_H1_EMPTY + (int)stg), h1_par ^ 1u, P.err, 32);
        if (!ok) break;
        um::tc_fence_after();
        unsigned char* hbase = base_ptr + P.sm_h1 + stg * P.h1_stage_bytes + row * 16;
        const uint32_t tcol = lane_addr + L1COL + lb * (uint32_t)CH;
        const int w = (c == nch - 1) ? last_w : CH;
        for (int j0 = 0; j0 < w; j0 += 16) {
          uint32_t v[16];
          um::tmem_ld16(tcol + (uint32_t)j0, v);
          um::tmem_ld_wait();
          uint4 p0, p1;
          p0.x = um::pack2_relu<PREC>(__uint_as_float(v[0]), __uint_as_float(v[1]));
          p0.y = um::pack2_relu<PREC>(__uint_as_float(v[2]), __uint_as_float(v[3]));
          p0.z = um::pack2_relu<PREC>(__uint_as_float(v[4]), __uint_as_float(v[5]));
          p0.w = um::pack2_relu<PREC>(__uint_as_float(v[6]), __uint_as_float(v[7]));
          p1.x = um::pack2_relu<PREC>(__uint_as_float(v[8]), __uint_as_float(v[9]));
          p1.y = um::pack2_relu<PREC>(__uint_as_float(v[10]), __uint_as_float(v[11]));
          p1.z = um::pack2_relu<PREC>(__uint_as_float(v[12]), __uint_as_float(v[13]));
          p1.w = um::pack2_relu<PREC>(__uint_as_float(v[14]), __uint_as_float(v[15]));
          *reinterpret_cast<uint4*>(hbase + (j0 / 8) * 2048) = p0;
          *reinterpret_cast<uint4*>(hbase + (j0 / 8 + 1) * 2048) = p1;
        }
        um::tc_fence_before();
        um::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          um::mbar_arrive_cluster(lb ? l1e1 : l1e0);
          um::mbar_arrive_cluster(h1f0 + 8u * stg);
        }
        if (++stg == UM_NST) { stg = 0; h1_par ^= 1u; }
      }
    }
  } else {
    // =================================== epilogue 2: L2 acc -> bias/relu/w3 dot -> q ============
    const int q4 = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t l2e0 = um::mapa(bar(BAR_L2_EMPTY + 0), 0);
    const uint32_t l2e1 = um::mapa(bar(BAR_L2_EMPTY + 1), 0);
    const float c0 = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0));
    bool ok = true;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      ok = um::mbar_wait(bar(BAR_L2_FULL), (uint32_t)(tl & 1), P.err, 41);
      if (!ok) break;
      um::tc_fence_after();
      float acc = 0.f;
      for (int half = 0; half < 2; ++half) {
        const int j_begin = half ? P.NA : 0, j_end = half ? P.H2P : P.NA;
        for (int j0 = j_begin; j0 < j_end; j0 += 16) {
          uint32_t v[16];
          um::tmem_ld16(lane_addr + (uint32_t)j0, v);
          um::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w3s + j0 + e);
            const float4 nb = *reinterpret_cast<const float4*>(nb2s + j0 + e);
            acc = fmaf(w.x, fmaxf(__uint_as_float(v[e + 0]), nb.x), acc);
            acc = fmaf(w.y, fmaxf(__uint_as_float(v[e + 1]), nb.y), acc);
            acc = fmaf(w.z, fmaxf(__uint_as_float(v[e + 2]), nb.z), acc);
            acc = fmaf(w.w, fmaxf(__uint_as_float(v[e + 3]), nb.w), acc);
          }
        }
        um::tc_fence_before();
        __syncwarp();
        if (lane == 0) um::mbar_arrive_cluster(half ? l2e1 : l2e0);
      }
      const long long tile = (long long)pair + (long long)tl * npairs;
      const long long row = tile * 256 + (long long)rank * 128 + q4 * 32 + lane;
      if (row < P.R) P.q[row] = acc + c0;
    }
  }

  // ---- teardown: nobody leaves while the peer may still read our smem / TMEM ----
  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  if (warp == 0) um::tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// K1 (TS variant, default): layer-2's A operand is read from TENSOR MEMORY.
//
// Epilogue 1 converts a layer-1 accumulator chunk (fp32, CH columns) to fp16 *in place* (CH/2
// columns, two values per 32-bit cell) with tcgen05.ld -> F2FP.RELU -> tcgen05.st, and layer 2 is
// issued as tcgen05.mma [d_tmem], [a_tmem], b_desc: the ReLU'd activations never touch shared
// memory, which removes ~100 KB of st.shared + 200 KB of tensor-core operand reads per 128-row
// tile from the 128 B/clk shared-memory port, the proxy fence, and one barrier (the chunk buffer
// is recycled by MMA issue order: L1(u+NB1) is issued after L2(u), and tcgen05.mma executes in
// order).  TMEM: [0,H2P) layer-2 accumulator, then NB1 chunk buffers of CH columns.
//
// 16 warps per CTA: 0 = MMA issuer (leader CTA), 1-3 = X producers (OR of two pre-packed 16-byte
// tables: state part SP[b][kc] and action part AP[kc][n]), 4-7 = epilogue 1, 8-15 = epilogue 2
// (two warps per TMEM lane quarter, each draining half of the columns of each accumulator half, so
// the accumulator is handed back to the MMA issuer after one tcgen05.ld round trip).
// ------------------------------------------------------------------------------------------
#define TS_THREADS 512
enum {
  TB_X_FULL = 0,     // [2]   count 6  (leader)   3 producer warps x 2 CTAs -> MMA
  TB_X_EMPTY = 2,    // [2]   count 1  (both)     MMA commit -> producers
  TB_L1_FULL = 4,    // [3]   count 1  (both)     MMA commit -> ep1
  TB_H1_FULL = 7,    // [3]   count 8  (leader)   ep1 (4 warps x 2 CTAs) -> MMA
  TB_L2_FULL = 10,   // [1]   count 1  (both)     MMA commit -> ep2
  TB_L2_EMPTY = 11,  // [2]   count 16 (leader)   ep2 (8 warps x 2 CTAs) -> MMA (half A, half B)
  TB_COUNT = 13
};

namespace um {
__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] += A[tmem] * B[smem]^T (accumulate always: the predicate is a constant, nothing to compute)
__device__ __forceinline__ void mma2_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
               "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                 "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
}  // namespace um

// acc += sign_p * relu(z) over NU units of 8 accumulator columns held in registers; units p < kpos belong
// to the non-negative-w3 group (+), the rest to the negative group (-).  Straight-line: per unit one
// compare/select for the sign, 8 FMNMX, 8 FFMA on four independent accumulators.
template <int NU>
__device__ __forceinline__ void relu_signed_sum(const uint32_t* v, int kpos, float& a0, float& a1, float& a2,
                                                float& a3) {
#pragma unroll
  for (int p = 0; p < NU; ++p) {
    const float sg = (p < kpos) ? 1.f : -1.f;
    a0 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 0]), 0.f), sg, a0);
    a1 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 1]), 0.f), sg, a1);
    a2 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 2]), 0.f), sg, a2);
    a3 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 3]), 0.f), sg, a3);
    a0 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 4]), 0.f), sg, a0);
    a1 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 5]), 0.f), sg, a1);
    a2 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 6]), 0.f), sg, a2);
    a3 = fmaf(fmaxf(__uint_as_float(v[p * 8 + 7]), 0.f), sg, a3);
  }
}

// One round of up to 96 columns (w, multiple of 8) whose first `rel` columns belong to the + group.
// Fast path: no unit straddles the group boundary (sign per unit, straight-line code specialised on the
// unit count); the one round per row that does straddle takes the per-column path.
__device__ __forceinline__ void relu_signed_round(const uint32_t* v, int w, int rel, float& a0, float& a1,
                                                  float& a2, float& a3) {
  if (rel <= 0 || rel >= w || (rel & 7) == 0) {
    const int kpos = rel <= 0 ? 0 : (rel >= w ? 12 : (rel >> 3));
    switch (w >> 3) {
      case 12: relu_signed_sum<12>(v, kpos, a0, a1, a2, a3); break;
      case 11: relu_signed_sum<11>(v, kpos, a0, a1, a2, a3); break;
      case 10: relu_signed_sum<10>(v, kpos, a0, a1, a2, a3); break;
      case 9: relu_signed_sum<9>(v, kpos, a0, a1, a2, a3); break;
      case 8: relu_signed_sum<8>(v, kpos, a0, a1, a2, a3); break;
      case 7: relu_signed_sum<7>(v, kpos, a0, a1, a2, a3); break;
      case 6: relu_signed_sum<6>(v, kpos, a0, a1, a2, a3); break;
      case 5: relu_signed_sum<5>(v, kpos, a0, a1, a2, a3); break;
      case 4: relu_signed_sum<4>(v, kpos, a0, a1, a2, a3); break;
      case 3: relu_signed_sum<3>(v, kpos, a0, a1, a2, a3); break;
      case 2: relu_signed_sum<2>(v, kpos, a0, a1, a2, a3); break;
      default: relu_signed_sum<1>(v, kpos, a0, a1, a2, a3); break;
    }
  } else {
#pragma unroll
    for (int p = 0; p < 12; ++p) {
      if (p * 8 < w) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          a0 = fmaf(fmaxf(__uint_as_float(v[p * 8 + e]), 0.f), (p * 8 + e >= rel) ? -1.f : 1.f, a0);
      }
    }
  }
}

// Pre-pass: 16-byte X chunks.  SP[b][kc] = fp16/bf16 of the (clipped) state entries that fall in
// K chunk kc, plus the ones column at k = S+A; AP[kc][n] = the action entries (shared grid only).
// A row of X is SP[b][kc] | AP[kc][n] (the two parts occupy disjoint halfwords; +0.0 = 0x0000).
template <int PREC>
__global__ void k_xparts(const float* __restrict__ s, const float* __restrict__ a,
                         const float* __restrict__ smin, const float* __restrict__ smax, int B, int N,
                         int S, int A, int KC1, int shared_actions, uint4* __restrict__ SP,
                         uint4* __restrict__ AP) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nsp = (long long)B * KC1, nap = shared_actions ? (long long)N * KC1 : 0;
  if (gid >= nsp + nap) return;
  float v[8];
  const int K1 = S + A;
  if (gid < nsp) {
    const int b = (int)(gid / KC1), kc = (int)(gid % KC1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float t = 0.f;
      if (k < S) {
        t = __ldg(s + (size_t)b * S + k);
        if (smin) t = fminf(fmaxf(t, __ldg(smin + k)), __ldg(smax + k));
      } else if (k == K1) {
        t = 1.f;
      }
      v[e] = t;
    }
  } else {
    const long long g = gid - nsp;
    const int kc = (int)(g / N), n = (int)(g % N);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      v[e] = (k >= S && k < K1) ? __ldg(a + (size_t)n * A + (k - S)) : 0.f;
    }
  }
  uint4 pk;
  pk.x = um::pack2<PREC>(v[0], v[1]);
  pk.y = um::pack2<PREC>(v[2], v[3]);
  pk.z = um::pack2<PREC>(v[4], v[5]);
  pk.w = um::pack2<PREC>(v[6], v[7]);
  if (gid < nsp) SP[gid] = pk;
  else AP[gid - nsp] = pk;
}

template <int PREC, bool PROF>
__global__ void __launch_bounds__(TS_THREADS, 1) k_critic_umma_ts(const UmmaParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = um::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = um::cta_rank();
  const uint32_t pair = um::cluster_id_x();
  const uint32_t npairs = um::num_clusters_x();

  const uint32_t sW2 = base + P.sm_w2, sW1 = base + P.sm_w1, sX = base + P.sm_x, sBar = base + P.sm_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + P.sm_bar + TB_COUNT * 8);
  float* qpart = reinterpret_cast<float*>(base_ptr + P.sm_h1);   // [2][128] partial dot products
  auto bar = [&](int i) -> uint32_t { return sBar + 8u * (uint32_t)i; };

  // ---- prologue: resident weights -> smem, barriers, TMEM ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.blob[rank]);
    uint4* dW2 = reinterpret_cast<uint4*>(base_ptr + P.sm_w2);
    const int n2 = (P.off_w1 - P.off_w2) >> 4;
    for (int i = tid; i < n2; i += TS_THREADS) dW2[i] = __ldg(src + (P.off_w2 >> 4) + i);
    uint4* dW1 = reinterpret_cast<uint4*>(base_ptr + P.sm_w1);
    const int n1 = (P.off_w3 - P.off_w1) >> 4;
    for (int i = tid; i < n1; i += TS_THREADS) dW1[i] = __ldg(src + (P.off_w1 >> 4) + i);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      um::mbar_init(bar(TB_X_FULL + i), 6);
      um::mbar_init(bar(TB_X_EMPTY + i), 1);
      um::mbar_init(bar(TB_L2_EMPTY + i), 16);
    }
    for (int i = 0; i < 3; ++i) {
      um::mbar_init(bar(TB_L1_FULL + i), 1);
      um::mbar_init(bar(TB_H1_FULL + i), 8);
    }
    um::mbar_init(bar(TB_L2_FULL), 1);
    um::fence_mbar_init();
  }
  um::fence_proxy_async();
  if (warp == 0) um::tmem_alloc2(um::smem_u32(tmem_slot), 512);
  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  um::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // optional cycle accounting (RLC_UMMA_PROF=1): 8 x int64 per role, 32 per pair
  long long pw0 = 0, pw1 = 0, pw2 = 0, pi0 = 0, pi1 = 0;
  const bool prof = PROF && P.prof != nullptr;
  const long long t_begin = prof ? clock64() : 0;
#define PROF_T() ((PROF && prof) ? clock64() : 0)

  const int ntiles = (P.num_pair_tiles > (int)pair)
                         ? (P.num_pair_tiles - (int)pair + (int)npairs - 1) / (int)npairs
                         : 0;
  const int nch = P.nch, CH = P.CH, last_w = P.H1P - (P.nch - 1) * P.CH;
  const uint32_t NB1 = (uint32_t)P.nb1;
  const uint32_t L1COL = (uint32_t)P.H2P;

  if (warp == 0) {
    // =================================== MMA issuer (leader CTA) ===================================
    if (rank == 0) {
      const bool issuer = um::elect_one();
      const uint32_t fmt = (PREC == RLC_PREC_BF16) ? 1u : 0u;
      const uint32_t lbo_x = 128u * 16u;
      const uint32_t lbo_w1 = (uint32_t)(P.H1P / 2) * 16u;
      const uint32_t lbo_w2 = (uint32_t)(P.H2P / 2) * 16u;
      const uint32_t x_lo0 = um::desc_lo(sX, lbo_x), x_lo1 = um::desc_lo(sX + P.x_stage_bytes, lbo_x);
      const uint32_t w1_lo = um::desc_lo(sW1, lbo_w1), w2_lo = um::desc_lo(sW2, lbo_w2);
      const uint32_t a_kstep = (2u * lbo_x) >> 4;
      const uint32_t w1_kstep = (2u * lbo_w1) >> 4, w2_kstep = (2u * lbo_w2) >> 4;
      const uint32_t w2_half_off = (uint32_t)(P.NA / 2);
      const uint32_t idA = um::make_idesc(fmt, 256, P.NA);
      const uint32_t idB = um::make_idesc(fmt, 256, P.NB > 0 ? P.NB : 16);
      const uint32_t id1_full = um::make_idesc(fmt, 256, CH);
      const uint32_t id1_last = um::make_idesc(fmt, 256, last_w);
      const int k1steps = P.K1P / 16;
      const bool two_halves = P.NB > 0;
      const uint32_t dL2A = tmem_base, dL2B = tmem_base + (uint32_t)P.NA;
      const int total = ntiles * nch;

      bool ok = true;
      // cursor of the layer-1 issue (runs NB1-1 units ahead of the layer-2 cursor)
      int tl1 = 0, c1 = 0;
      uint32_t lb1 = 0;
      auto issue_l1 = [&]() {
        const int xs = tl1 & 1;
        long long t0 = PROF_T();
        if (c1 == 0) ok = ok && um::mbar_wait(bar(TB_X_FULL + xs), (uint32_t)((tl1 >> 1) & 1), P.err, 11);
        um::tc_fence_after();
        long long t1 = PROF_T();
        pw0 += t1 - t0;
        if (ok && issuer) {
          const uint32_t d = tmem_base + L1COL + lb1 * (uint32_t)CH;
          const uint32_t a0 = xs ? x_lo1 : x_lo0;
          const uint32_t b0 = w1_lo + (uint32_t)(c1 * (CH >> 1));
          const uint32_t idesc = (c1 == nch - 1) ? id1_last : id1_full;
#pragma unroll 1
          for (int k = 0; k < k1steps; ++k)
            um::mma2(d, um::desc64(a0 + (uint32_t)k * a_kstep), um::desc64(b0 + (uint32_t)k * w1_kstep), idesc,
                     (uint32_t)(k > 0));
          um::commit2(bar(TB_L1_FULL + (int)lb1));
          if (c1 == nch - 1) um::commit2(bar(TB_X_EMPTY + xs));
        }
        __syncwarp();
        pi0 += PROF_T() - t1;
        if (++lb1 == NB1) lb1 = 0;
        if (++c1 == nch) { c1 = 0; ++tl1; }
      };
      int issued1 = 0;
      const int ahead = (int)NB1 - 1;
      for (; issued1 < ahead && issued1 < total && ok; ++issued1) issue_l1();

      uint32_t lb2 = 0, par2 = 0;   // chunk buffer / parity of the layer-2 cursor
      for (int tl = 0; tl < ntiles && ok; ++tl) {
        for (int c = 0; c < nch && ok; ++c) {
          if (issued1 < total) { issue_l1(); ++issued1; }
          // ---- layer 2 over K-slice c: A = fp16 activations in TMEM, B = W2 slice in smem ----
          long long t0 = PROF_T();
          ok = ok && um::mbar_wait(bar(TB_H1_FULL + (int)lb2), par2, P.err, 13);
          long long t1 = PROF_T();
          pw1 += t1 - t0;
          const int ksteps = (c == nch - 1 ? last_w : CH) >> 4;
          const uint32_t a0 = tmem_base + L1COL + lb2 * (uint32_t)CH;       // 8 columns per K=16 step
          const uint32_t b0 = w2_lo + (uint32_t)(c * (CH >> 3)) * (lbo_w2 >> 4);
          const uint32_t acc0 = c > 0 ? 1u : 0u;
          if (c == 0) ok = ok && um::mbar_wait(bar(TB_L2_EMPTY + 0), (uint32_t)((tl & 1) ^ 1), P.err, 14);
          um::tc_fence_after();
          long long t2 = PROF_T();
          pw2 += t2 - t1;
          if (ok && issuer) {
            if (ksteps == 6) {      // full chunk: back-to-back issue, operands precomputed
              um::mma2_ts(dL2A, a0, um::desc64(b0), idA, acc0);
#pragma unroll
              for (int k = 1; k < 6; ++k)
                um::mma2_ts_acc(dL2A, a0 + 8u * (uint32_t)k, um::desc64(b0 + (uint32_t)k * w2_kstep), idA);
            } else {
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k)
                um::mma2_ts(dL2A, a0 + 8u * (uint32_t)k, um::desc64(b0 + (uint32_t)k * w2_kstep), idA,
                            acc0 | (uint32_t)(k > 0));
            }
          }
          if (two_halves) {
            if (c == 0) {
              long long t3 = PROF_T();
              ok = ok && um::mbar_wait(bar(TB_L2_EMPTY + 1), (uint32_t)((tl & 1) ^ 1), P.err, 15);
              um::tc_fence_after();
              long long t4 = PROF_T();
              pw2 += t4 - t3;
              t2 += t4 - t3;
            }
            if (ok && issuer) {
              if (ksteps == 6) {
                um::mma2_ts(dL2B, a0, um::desc64(b0 + w2_half_off), idB, acc0);
#pragma unroll
                for (int k = 1; k < 6; ++k)
                  um::mma2_ts_acc(dL2B, a0 + 8u * (uint32_t)k, um::desc64(b0 + w2_half_off + (uint32_t)k * w2_kstep), idB);
              } else {
#pragma unroll 1
                for (int k = 0; k < ksteps; ++k)
                  um::mma2_ts(dL2B, a0 + 8u * (uint32_t)k, um::desc64(b0 + w2_half_off + (uint32_t)k * w2_kstep),
                              idB, acc0 | (uint32_t)(k > 0));
              }
            }
          }
          if (ok && issuer && c == nch - 1) um::commit2(bar(TB_L2_FULL));
          __syncwarp();
          pi1 += PROF_T() - t2;
          if (++lb2 == NB1) { lb2 = 0; par2 ^= 1u; }
        }
      }
      if (prof && issuer) {
        long long* o = P.prof + (size_t)pair * 32;
        o[0] = clock64() - t_begin; o[1] = pw0; o[2] = pw1; o[3] = pw2; o[4] = pi0; o[5] = pi1; o[6] = ntiles;
      }
    }
  } else if (warp < 4) {
    // =================================== X producers (3 warps) ===================================
    // X[r][kc] = SP[b][kc] | AP[kc][n]  (two 16-byte loads and an OR per chunk); per-state actions
    // are packed inline from a[B,N,A].
    const int pt = tid - 32;  // 0..95
    const uint32_t xfull_leader0 = um::mapa(bar(TB_X_FULL + 0), 0);
    const uint32_t xfull_leader1 = um::mapa(bar(TB_X_FULL + 1), 0);
    bool ok = true;
    const int S = P.S, K1 = P.S + P.A, KC1 = P.KC1, A = P.A;
    const uint32_t N = (uint32_t)P.N;
    const uint4* __restrict__ SP = P.sp;
    const uint4* __restrict__ AP = P.ap;
    const int kc_a0 = S >> 3, kc_a1 = (K1 - 1) >> 3;   // K chunks that contain action entries
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      const int xs = tl & 1;
      long long t0 = PROF_T();
      ok = um::mbar_wait(bar(TB_X_EMPTY + xs), (uint32_t)(((tl >> 1) & 1) ^ 1), P.err, 21);
      if (!ok) break;
      long long t1 = PROF_T();
      pw0 += t1 - t0;
      const long long tile = (long long)pair + (long long)tl * npairs;
      const long long row0 = tile * 256 + (long long)rank * 128;
      unsigned char* xbase = base_ptr + P.sm_x + xs * P.x_stage_bytes;
      for (int r = pt; r < 128; r += 96) {
        const long long row = row0 + r;
        uint4* dst = reinterpret_cast<uint4*>(xbase + r * 16);
        if (row < P.R) {
          const uint32_t b = (uint32_t)row / N, n = (uint32_t)row - b * N;
          const uint4* sp = SP + (size_t)b * KC1;
          if (!P.act_per_state) {
            for (int kc = 0; kc < KC1; ++kc) {
              uint4 x = __ldg(sp + kc);
              const uint4 y = __ldg(AP + (size_t)kc * N + n);
              x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w;
              dst[kc * 128] = x;
            }
          } else {
            const float* ap = P.a + (size_t)(uint32_t)row * A;
            for (int kc = 0; kc < KC1; ++kc) {
              uint4 x = __ldg(sp + kc);
              if (kc >= kc_a0 && kc <= kc_a1) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int k = kc * 8 + e;
                  v[e] = (k >= S && k < K1) ? __ldg(ap + (k - S)) : 0.f;
                }
                x.x |= um::pack2<PREC>(v[0], v[1]);
                x.y |= um::pack2<PREC>(v[2], v[3]);
                x.z |= um::pack2<PREC>(v[4], v[5]);
                x.w |= um::pack2<PREC>(v[6], v[7]);
              }
              dst[kc * 128] = x;
            }
          }
        } else {
          for (int kc = 0; kc < KC1; ++kc) dst[kc * 128] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      um::fence_proxy_async();
      __syncwarp();
      if (lane == 0) um::mbar_arrive_cluster(xs ? xfull_leader1 : xfull_leader0);
      pi0 += PROF_T() - t1;
    }
    if (prof && rank == 0 && tid == 32) { long long* o = P.prof + (size_t)pair * 32 + 8; o[0] = pw0; o[1] = pi0; }
  } else if (warp < 8) {
    // ============ epilogue 1: L1 accumulator chunk -> relu -> fp16, in place in TMEM ============
    // All tcgen05.ld of the chunk are issued back to back (<= 96 columns), one wait, convert,
    // then tcgen05.st into the first half of the same columns.
    const int q4 = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t h1f0 = um::mapa(bar(TB_H1_FULL), 0);
    bool ok = true;
    uint32_t lb = 0, par = 0;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      for (int c = 0; c < nch; ++c) {
        long long t0 = PROF_T();
        ok = um::mbar_wait(bar(TB_L1_FULL + (int)lb), par, P.err, 31);
        if (!ok) break;
        um::tc_fence_after();
        long long t1 = PROF_T();
        pw0 += t1 - t0;
        const uint32_t tcol = lane_addr + L1COL + lb * (uint32_t)CH;
        const int w = (c == nch - 1) ? last_w : CH;
        uint32_t v[96];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          if (p * 32 + 32 <= w) um::tmem_ld32p(tcol + (uint32_t)(p * 32), v + p * 32);
          else if (p * 32 < w) um::tmem_ld16p(tcol + (uint32_t)(p * 32), v + p * 32);
        }
        um::tmem_ld_wait();
        { long long tt = PROF_T(); pw1 += tt - t1; }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          if (p * 32 < w) {
            uint32_t o[16];
#pragma unroll
            for (int e = 0; e < 16; ++e)
              o[e] = um::pack2_relu<PREC>(__uint_as_float(v[p * 32 + 2 * e]), __uint_as_float(v[p * 32 + 2 * e + 1]));
            if (p * 32 + 32 <= w) um::tmem_st16p(tcol + (uint32_t)(p * 16), o);
            else um::tmem_st8(tcol + (uint32_t)(p * 16), o);
          }
        }
        long long t6 = PROF_T();
        um::tmem_st_wait();
        long long t7 = PROF_T();
        pw2 += t6 - t1;      // ld+wait + cvt + st issue
        pi1 += t7 - t6;      // wait::st
        um::tc_fence_before();
        __syncwarp();
        if (lane == 0) um::mbar_arrive_cluster(h1f0 + 8u * lb);
        pi0 += PROF_T() - t1;
        if (++lb == NB1) { lb = 0; par ^= 1u; }
      }
    }
    if (prof && rank == 0 && tid == 128) { long long* o = P.prof + (size_t)pair * 32 + 16; o[0] = pw0; o[1] = pi0; o[2] = pw1; o[3] = pw2; o[4] = pi1; }
  } else {
    // ================ epilogue 2 (8 warps): L2 acc -> relu -> signed sum -> q ====================
    // warp w: TMEM lane quarter w&3, column half (w-8)>>2 of each accumulator half.  The columns
    // are pulled into registers with back-to-back tcgen05.ld, the accumulator half is released to
    // the MMA issuer right after the wait, the arithmetic runs from registers.  The output head
    // is folded into W2 (see k_pack_head): q = b3 + 2^-k (sum_{j<npos} relu - sum_{j>=npos} relu),
    // so there are no per-column constants to fetch.
    const int q4 = warp & 3, chalf = (warp - 8) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t l2e0 = um::mapa(bar(TB_L2_EMPTY + 0), 0);
    const uint32_t l2e1 = um::mapa(bar(TB_L2_EMPTY + 1), 0);
    const int npos = __ldg(reinterpret_cast<const int*>(P.blob[0] + P.off_c0));
    const float inv_scale = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 2);
    const float b3v = __ldg(reinterpret_cast<const float*>(P.blob[0] + P.off_c0) + 3);
    const int rloc = q4 * 32 + lane;
    bool ok = true;
    for (int tl = 0; tl < ntiles && ok; ++tl) {
      long long t0 = PROF_T();
      ok = um::mbar_wait(bar(TB_L2_FULL), (uint32_t)(tl & 1), P.err, 41);
      if (!ok) break;
      um::tc_fence_after();
      long long t1 = PROF_T();
      pw0 += t1 - t0;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int half = 0; half < 2; ++half) {
        const int hb = half ? P.NA : 0, hn = half ? P.NB : P.NA;      // this accumulator half
        // my columns: [j_begin, j_end) -- split of the half in units of 8 columns
        const int units = hn >> 3, u0 = chalf ? (units + 1) / 2 : 0, u1 = chalf ? units : (units + 1) / 2;
        const int j_begin = hb + u0 * 8, j_end = hb + u1 * 8;
        for (int jb = j_begin; jb < j_end; jb += 96) {
          const int w = (j_end - jb < 96) ? (j_end - jb) : 96;
          uint32_t v[96];
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            if (p * 32 + 32 <= w) um::tmem_ld32p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
            else if (p * 32 < w) {
              const int rem = w - p * 32;             // 8, 16 or 24
              if (rem >= 16) um::tmem_ld16p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
              if (rem == 8) um::tmem_ld8p(lane_addr + (uint32_t)(jb + p * 32), v + p * 32);
              if (rem == 24) um::tmem_ld8p(lane_addr + (uint32_t)(jb + p * 32 + 16), v + p * 32 + 16);
            }
          }
          um::tmem_ld_wait();
          if (half == 0) { long long tt = PROF_T(); pw1 += tt - t1; }
          if (jb + 96 >= j_end) {               // last round of this half: hand it back
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
      }
      // combine the two column halves of each row through shared memory
      const float acc = (a0 + a1) + (a2 + a3);
      float* qp = qpart + (tl & 1) * 128;
      if (chalf) qp[rloc] = acc;
      long long t5 = PROF_T();
      um::named_bar_sync(1, 256);
      pw2 += PROF_T() - t5;
      if (!chalf) {
        const long long tile = (long long)pair + (long long)tl * npairs;
        const long long row = tile * 256 + (long long)rank * 128 + rloc;
        if (row < P.R) P.q[row] = fmaf(inv_scale, acc + qp[rloc], b3v);
      }
      pi0 += PROF_T() - t1;
    }
    if (prof && rank == 0 && tid == 256) { long long* o = P.prof + (size_t)pair * 32 + 24; o[0] = pw0; o[1] = pi0; o[2] = pw1; o[3] = pw2; }
  }
#undef PROF_T

  um::tc_fence_before();
  __syncthreads();
  um::cluster_sync();
  if (warp == 0) um::tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// Weight pre-pack: theta (fp32 canonical) -> per-rank blobs in the UMMA core-matrix layout
// ------------------------------------------------------------------------------------------
struct PackGeom {
  int S, A, H1, H2, K1P, H1P, H2P, NA, NB, nch, CH, nb1, ts;
  int off_w2, off_w1, off_w3, off_nb2, off_c0, blob_bytes;
};

template <int PREC>
__device__ __forceinline__ unsigned short to_h(float x) {
  if (PREC == RLC_PREC_BF16) {
    __nv_bfloat16 t = __float2bfloat16_rn(x);
    return *reinterpret_cast<unsigned short*>(&t);
  }
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  __half t = __float2half_rn(x);
  return *reinterpret_cast<unsigned short*>(&t);
}

// TS variant, output head folded into layer 2 (so that epilogue 2 needs no per-column constants):
//   q = b3 + sum_j w3_j relu(z_j + b2_j)
//     = b3 + 2^-k ( sum_{w3_j >= 0} relu(z'_j) - sum_{w3_j < 0} relu(z'_j) ),
//   z'_j = h1 . (2^k |w3_j| W2[:,j]) + 2^k |w3_j| b2_j
// The columns of W2 are scaled by 2^k |w3_j| (k: max_j 2^k |w3_j| in [1,2), exact) and stably
// partitioned by the sign of w3_j (non-negative first); the bias rides on an always-one layer-1
// feature (index H1: zero W1 column, bias 1).  This kernel computes the partition and k.
// blob area off_w3: int inv[H2P] (destination column -> source column, -1 = zero column);
// blob slot off_c0: {int npos; float scale; float inv_scale; float b3}.
__global__ void k_pack_head(const float* __restrict__ theta, PackGeom G, unsigned char* blob0,
                            unsigned char* blob1) {
  const ThetaView t = theta_view(RLC_TIN, G.S, G.A, G.H1, G.H2);
  const float* w3 = theta + t.ow3;
  const float* b3 = theta + t.ob3;
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;   // one warp: ballot-based stable partition
  const int lane = threadIdx.x;
  int* inv0 = reinterpret_cast<int*>(blob0 + G.off_w3);
  int* inv1 = reinterpret_cast<int*>(blob1 + G.off_w3);
  int npos = 0;
  float mx = 0.f;
  for (int j0 = 0; j0 < G.H2; j0 += 32) {
    const int j = j0 + lane;
    const float w = (j < G.H2) ? w3[j] : 0.f;
    npos += __popc(__ballot_sync(0xffffffffu, j < G.H2 && !(w < 0.f)));
    mx = fmaxf(mx, fabsf(w));
  }
  mx = warp_max(mx);
  int p = 0, n = npos;
  const unsigned lt = (1u << lane) - 1u;
  for (int j0 = 0; j0 < G.H2; j0 += 32) {
    const int j = j0 + lane;
    const bool live = j < G.H2;
    const bool pos = live && !(w3[j] < 0.f);
    const unsigned mp = __ballot_sync(0xffffffffu, pos), mn = __ballot_sync(0xffffffffu, live && !pos);
    if (live) {
      const int d = pos ? p + __popc(mp & lt) : n + __popc(mn & lt);
      inv0[d] = j;
      inv1[d] = j;
    }
    p += __popc(mp);
    n += __popc(mn);
  }
  for (int d = G.H2 + lane; d < G.H2P; d += 32) { inv0[d] = -1; inv1[d] = -1; }
  if (lane != 0) return;
  int e = 0;
  float scale = 1.f;
  if (mx > 0.f && isfinite(mx)) {
    (void)frexpf(mx, &e);              // mx = m * 2^e, m in [0.5,1)  ->  2^(1-e) * mx in [1,2)
    scale = ldexpf(1.f, 1 - e);
  }
  for (int r = 0; r < 2; ++r) {
    unsigned char* blob = r ? blob1 : blob0;
    reinterpret_cast<int*>(blob + G.off_c0)[0] = npos;
    reinterpret_cast<float*>(blob + G.off_c0)[1] = scale;
    reinterpret_cast<float*>(blob + G.off_c0)[2] = 1.f / scale;
    reinterpret_cast<float*>(blob + G.off_c0)[3] = b3[0];
  }
}

template <int PREC>
__global__ void k_pack_umma(const float* __restrict__ theta, PackGeom G, unsigned char* blob0,
                            unsigned char* blob1) {
  const ThetaView t = theta_view(RLC_TIN, G.S, G.A, G.H1, G.H2);
  const float* W1 = theta + t.oW1;   // [S+A][H1]
  const float* b1 = theta + t.ob1;
  const float* W2 = theta + t.oW2;   // [H1][H2]
  const float* b2 = theta + t.ob2;
  const float* w3 = theta + t.ow3;
  const float* b3 = theta + t.ob3;
  const int K1 = G.S + G.A;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nW2 = (long long)G.H1P * G.H2P;
  const long long nW1 = (long long)G.K1P * G.H1P;
  if (gid < nW2) {
    const int k = (int)(gid / G.H2P), n = (int)(gid % G.H2P);
    // which instruction / rank / local row
    int rank, l;
    if (n < G.NA) {
      const int hN = G.NA / 2;
      rank = n / hN;
      l = n % hN;
    } else {
      const int hN = G.NB / 2, m = n - G.NA;
      rank = m / hN;
      l = G.NA / 2 + m % hN;
    }
    float v = 0.f;
    if (G.ts) {
      const int j = reinterpret_cast<const int*>(blob0 + G.off_w3)[n];   // written by k_pack_head
      if (j >= 0 && k <= G.H1) {
        const float sw = reinterpret_cast<const float*>(blob0 + G.off_c0)[1] * fabsf(w3[j]);
        v = sw * (k < G.H1 ? W2[(long long)k * G.H2 + j] : b2[j]);       // one fp32 rounding, then fp16
      }
    } else {
      v = (k < G.H1 && n < G.H2) ? W2[(long long)k * G.H2 + n] : 0.f;
    }
    unsigned char* blob = rank ? blob1 : blob0;
    const long long off = (long long)G.off_w2 + (long long)(k / 8) * ((G.H2P / 2) * 16) +
                          (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<unsigned short*>(blob + off) = to_h<PREC>(v);
  } else if (gid < nW2 + nW1) {
    const long long g = gid - nW2;
    const int k = (int)(g / G.H1P), j = (int)(g % G.H1P);
    const int c = j / G.CH, c0f = c * G.CH;
    const int cw = (G.H1P - c0f < G.CH) ? (G.H1P - c0f) : G.CH;
    const int within = j - c0f, hw = cw / 2;
    const int rank = within / hw;
    const int l = c0f / 2 + within % hw;
    float v = 0.f;
    if (j < G.H1) {
      if (k < K1) v = W1[(long long)k * G.H1 + j];
      else if (k == K1) v = b1[j];
    } else if (G.ts && j == G.H1 && k == K1) {
      v = 1.f;                                    // the always-one feature that carries layer 2's bias
    }
    unsigned char* blob = rank ? blob1 : blob0;
    const long long off = (long long)G.off_w1 + (long long)(k / 8) * ((G.H1P / 2) * 16) +
                          (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<unsigned short*>(blob + off) = to_h<PREC>(v);
  } else if (!G.ts && gid < nW2 + nW1 + G.H2P) {
    const int j = (int)(gid - nW2 - nW1);
    const float w = (j < G.H2) ? w3[j] : 0.f;
    const float nb = (j < G.H2) ? -b2[j] : 0.f;
    for (int r = 0; r < 2; ++r) {
      unsigned char* blob = r ? blob1 : blob0;
      reinterpret_cast<float*>(blob + G.off_w3)[j] = w;
      reinterpret_cast<float*>(blob + G.off_nb2)[j] = nb;
    }
  } else if (!G.ts && gid == nW2 + nW1 + G.H2P) {
    float c0 = b3[0];
    for (int j = 0; j < G.H2; ++j) c0 = fmaf(w3[j], b2[j], c0);
    *reinterpret_cast<float*>(blob0 + G.off_c0) = c0;
    *reinterpret_cast<float*>(blob1 + G.off_c0) = c0;
  }
}

static bool make_geom(const rlc_critic* c, PackGeom& G, int ts) {
  memset(&G, 0, sizeof(G));
  G.S = c->S; G.A = c->A; G.H1 = c->H1; G.H2 = c->H2; G.ts = ts;
  G.K1P = (c->S + c->A + 1 + 15) & ~15;
  G.H1P = (c->H1 + (ts ? 1 : 0) + 15) & ~15;   // TS: + the always-one feature (bias of layer 2)
  G.H2P = (c->H2 + 15) & ~15;
  if (G.K1P > 64 || G.H2P > 480 || G.H2P < 32) return false;
  // layer-2 N split: one instruction if <= 256 else two halves rounded to 16
  if (G.H2P <= 256) { G.NA = G.H2P; G.NB = 0; }
  else { G.NA = ((G.H2P / 2) + 15) & ~15; G.NB = G.H2P - G.NA; }
  {
    const char* e = getenv("RLC_UMMA_NA");      // tuning knob: columns of the first layer-2 MMA
    if (e) {
      const int na = atoi(e);
      if (na >= 16 && (na & 15) == 0 && na <= 256 && na <= G.H2P && G.H2P - na <= 256) { G.NA = na; G.NB = G.H2P - na; }
    }
  }
  if (G.NA > 256 || G.NB > 256) return false;
  const int free_cols = 512 - G.H2P;
  int ch;
  if (ts) {
    // three chunk buffers when they can be >= 32 columns wide, else two
    // widest chunks first (fewer barrier round trips, longer MMAs); a third buffer only if free
    G.nb1 = 2;
    ch = (free_cols / 2) & ~15;
    if (ch > 96) ch = 96;                       // epilogue 1 holds one chunk in registers
    if (ch == 96 && free_cols >= 3 * 96) G.nb1 = 3;
    const char* e = getenv("RLC_UMMA_TS_CH");   // tuning knob: "<nb1>x<ch>"
    if (e) {
      int nb = 0, cw = 0;
      if (sscanf(e, "%dx%d", &nb, &cw) == 2 && (nb == 2 || nb == 3) && cw >= 16 && (cw & 15) == 0 &&
          cw <= 96 && nb * cw <= free_cols) { G.nb1 = nb; ch = cw; }
    }
  } else {
    G.nb1 = 2;
    ch = (free_cols / 2) & ~15;
    if (ch > 80) ch = 80;
  }
  if (ch < 16) return false;
  if (ch > G.H1P) ch = G.H1P;
  G.CH = ch;
  G.nch = (G.H1P + ch - 1) / ch;
  G.off_w2 = 0;
  G.off_w1 = G.off_w2 + (G.H1P / 8) * (G.H2P / 2) * 16;
  G.off_w3 = G.off_w1 + (G.K1P / 8) * (G.H1P / 2) * 16;
  G.off_nb2 = G.off_w3 + G.H2P * 4;
  G.off_c0 = G.off_nb2 + G.H2P * 4;
  G.blob_bytes = G.off_c0 + 16;
  return true;
}

struct SmemPlan {
  int sm_w2, sm_w1, sm_x, sm_h1, sm_par, sm_bar, x_stage, h1_stage, total;
};

static SmemPlan plan_smem(const PackGeom& G) {
  SmemPlan p;
  p.sm_w2 = 0;
  p.sm_w1 = p.sm_w2 + (G.off_w1 - G.off_w2);
  p.sm_par = p.sm_w1 + (G.off_w3 - G.off_w1);
  p.sm_x = (p.sm_par + (G.off_c0 - G.off_w3) + 127) & ~127;
  p.x_stage = (G.K1P / 8) * 2048;
  p.sm_h1 = p.sm_x + 2 * p.x_stage;
  p.h1_stage = G.ts ? 0 : (G.CH / 8) * 2048;     // TS variant keeps the activations in TMEM
  p.sm_bar = p.sm_h1 + (G.ts ? 1024 : UM_NST * p.h1_stage);   // TS: [2][128] fp32 partial dots
  p.total = p.sm_bar + BAR_COUNT * 8 + 16 + 1024;  // + alignment slack
  return p;
}

// 1 = TS variant (layer-2 A operand from tensor memory, default), 0 = SS variant (shared-memory ring)
static int umma_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("RLC_UMMA_MODE");
    mode = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;
  }
  return mode;
}

bool rlc_umma_supported(const rlc_handle* h, const rlc_critic* c, int B, int N) {
  if (h->sm_major != 10 || c->topology != RLC_TIN) return false;
  PackGeom G;
  if (!make_geom(c, G, umma_mode())) return false;
  const SmemPlan p = plan_smem(G);
  if ((size_t)p.total > h->smem_optin) return false;
  (void)B; (void)N;
  return true;
}

struct Grid3Parts;
static int launch_pack_x3(rlc_handle* h, const float* theta, const PackGeom& G, int mode, unsigned char* b0,
                          unsigned char* b1, cudaStream_t st);
__global__ void k_pack_scale3(const float* __restrict__ theta, PackGeom G, unsigned char* blob0, unsigned char* blob1);

static int get_pack(rlc_handle* h, const rlc_critic* c, int prec, const PackGeom& G,
                    cudaStream_t st, rlc_pack** out) {
  rlc_pack* slot = nullptr;
  for (int i = 0; i < RLC_MAX_PACKS; ++i) {
    rlc_pack& p = h->packs[i];
    if (p.theta == c->theta && p.prec == prec && p.S == c->S && p.A == c->A && p.H1 == c->H1 &&
        p.H2 == c->H2 && p.ch == G.CH && p.dev) { slot = &p; break; }
  }
  if (!slot) {
    slot = &h->packs[h->pack_rr];
    h->pack_rr = (h->pack_rr + 1) % RLC_MAX_PACKS;
    if (slot->dev && slot->bytes < (size_t)2 * G.blob_bytes) {
      const int rcr = rlc_retire_block(h, slot->dev);     // kept alive for graphs captured with the old pack
      if (rcr) return rcr;
      slot->dev = nullptr;
    }
    if (!slot->dev) {
      if (cudaMalloc(&slot->dev, (size_t)2 * G.blob_bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return RLC_ERR_ALLOC;
      }
      slot->bytes = (size_t)2 * G.blob_bytes;
    }
    slot->theta = c->theta; slot->prec = prec; slot->topology = c->topology;
    slot->S = c->S; slot->A = c->A; slot->H1 = c->H1; slot->H2 = c->H2; slot->ch = G.CH;
    slot->valid = false;
  }
  if (!slot->valid) {
    unsigned char* b0 = (unsigned char*)slot->dev;
    unsigned char* b1 = b0 + G.blob_bytes;
    const long long n = (long long)G.H1P * G.H2P + (long long)G.K1P * G.H1P + G.H2P + 1;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (G.ts) {
      k_pack_head<<<1, 32, 0, st>>>(c->theta, G, b0, b1);
      RLC_LAUNCH_CHECK(h);
    }
    if (prec == RLC_PREC_FP16X3 || prec == RLC_PREC_FP16C8) {   // split modes: hi | lo weight blobs, extra power-of-two column scale
      k_pack_scale3<<<1, 1024, 0, st>>>(c->theta, G, b0, b1);
      RLC_LAUNCH_CHECK(h);
      const int rc3 = launch_pack_x3(h, c->theta, G, prec == RLC_PREC_FP16C8 ? 1 : 0, b0, b1, st);
      if (rc3) return rc3;
    } else if (prec == RLC_PREC_BF16) k_pack_umma<RLC_PREC_BF16><<<blocks, 256, 0, st>>>(c->theta, G, b0, b1);
    else k_pack_umma<RLC_PREC_FP16><<<blocks, 256, 0, st>>>(c->theta, G, b0, b1);
    RLC_LAUNCH_CHECK(h);
    slot->valid = true;
  }
  *out = slot;
  return RLC_OK;
}

#include "critic_umma_grid.cuh"
#include "critic_umma_grid3.cuh"

int rlc_eval_umma(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                  int act_mode, int prec, float* q_out, cudaStream_t st) {
  if (prec == RLC_PREC_FP16X3 || prec == RLC_PREC_FP16C8) {   // split modes: shared grids only, never a silent downgrade
    if (act_mode != RLC_ACT_SHARED) return RLC_ERR_UNSUPPORTED;
    return rlc_eval_umma_grid3(h, c, s, B, a, N, prec, q_out, st);
  }
  PackGeom G;
  const int ts = umma_mode();
  if (!make_geom(c, G, ts)) return RLC_ERR_UNSUPPORTED;
  if ((long long)B * N >= (1ll << 31)) return RLC_ERR_UNSUPPORTED;  // 32-bit row arithmetic in the producers
  const SmemPlan sp = plan_smem(G);
  if ((size_t)sp.total > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  rlc_pack* pk = nullptr;
  int rc = get_pack(h, c, prec, G, st, &pk);
  if (rc) return rc;
  if (ts && act_mode == RLC_ACT_SHARED && grid_mode()) {
    rc = rlc_eval_umma_grid(h, c, G, pk, s, B, a, N, prec, q_out, st);
    if (rc != RLC_ERR_UNSUPPORTED) return rc;       // geometry the grid kernel cannot take: generic TS kernel
  }
  int* err = h->err_flag;

  UmmaParams P;
  memset(&P, 0, sizeof(P));
  P.s = s; P.a = a; P.smin = c->smin; P.smax = c->smax; P.q = q_out;
  P.R = (long long)B * N; P.N = N; P.S = c->S; P.A = c->A;
  P.act_per_state = act_mode == RLC_ACT_PER_STATE;
  P.K1P = G.K1P; P.KC1 = G.K1P / 8; P.H1P = G.H1P; P.KC2 = G.H1P / 8; P.H2P = G.H2P;
  P.NA = G.NA; P.NB = G.NB; P.nch = G.nch; P.CH = G.CH; P.nb1 = G.nb1;
  P.blob[0] = (const unsigned char*)pk->dev;
  P.blob[1] = P.blob[0] + G.blob_bytes;
  P.off_w2 = G.off_w2; P.off_w1 = G.off_w1; P.off_w3 = G.off_w3; P.off_nb2 = G.off_nb2;
  P.off_c0 = G.off_c0; P.blob_bytes = G.blob_bytes;
  P.sm_w2 = sp.sm_w2; P.sm_w1 = sp.sm_w1; P.sm_x = sp.sm_x; P.sm_h1 = sp.sm_h1;
  P.sm_par = sp.sm_par; P.sm_bar = sp.sm_bar; P.x_stage_bytes = sp.x_stage;
  P.h1_stage_bytes = sp.h1_stage;
  P.num_pair_tiles = (int)((P.R + 255) / 256);
  P.err = err;

  int pairs = h->num_sms / 2;
  if (pairs > P.num_pair_tiles) pairs = P.num_pair_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(ts ? TS_THREADS : UM_THREADS);
  cfg.dynamicSmemBytes = (size_t)sp.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (ts) {
    const int KC1 = G.K1P / 8;
    const int shared = act_mode == RLC_ACT_SHARED;
    const size_t nsp = (size_t)B * KC1, nap = shared ? (size_t)N * KC1 : 0;
    void* ws = nullptr;
    rc = rlc_workspace(h, (nsp + nap) * sizeof(uint4), &ws);
    if (rc) return rc;
    uint4* SP = (uint4*)ws;
    uint4* AP = SP + nsp;
    const unsigned blocks = (unsigned)((nsp + nap + 255) / 256);
    if (prec == RLC_PREC_BF16)
      k_xparts<RLC_PREC_BF16><<<blocks, 256, 0, st>>>(s, a, c->smin, c->smax, B, N, c->S, c->A, KC1, shared, SP, AP);
    else
      k_xparts<RLC_PREC_FP16><<<blocks, 256, 0, st>>>(s, a, c->smin, c->smax, B, N, c->S, c->A, KC1, shared, SP, AP);
    RLC_LAUNCH_CHECK(h);
    P.sp = SP; P.ap = AP;
  }
  static int prof_on = -1;
  static long long* prof_dev = nullptr;
  if (prof_on < 0) { const char* e = getenv("RLC_UMMA_PROF"); prof_on = (e && e[0] == '1') ? 1 : 0; }
  if (prof_on && ts) {
    if (!prof_dev) RLC_CUDA(cudaMalloc(&prof_dev, 128 * 32 * sizeof(long long)));
    RLC_CUDA(cudaMemsetAsync(prof_dev, 0, 128 * 32 * sizeof(long long), st));
    P.prof = prof_dev;
  }
  void (*kern)(const UmmaParams) =
      ts ? (P.prof ? (prec == RLC_PREC_BF16 ? k_critic_umma_ts<RLC_PREC_BF16, true> : k_critic_umma_ts<RLC_PREC_FP16, true>)
                   : (prec == RLC_PREC_BF16 ? k_critic_umma_ts<RLC_PREC_BF16, false> : k_critic_umma_ts<RLC_PREC_FP16, false>))
         : (prec == RLC_PREC_BF16 ? k_critic_umma<RLC_PREC_BF16> : k_critic_umma<RLC_PREC_FP16>);
  RLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
  RLC_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  if (prof_on && ts) {   // debug only: synchronises
    static long long hostp[128 * 32];
    RLC_CUDA(cudaStreamSynchronize(st));
    RLC_CUDA(cudaMemcpy(hostp, prof_dev, sizeof(hostp), cudaMemcpyDeviceToHost));
    const long long* o = hostp;  // pair 0
    const double T = (double)o[0], nt = (double)(o[6] > 0 ? o[6] : 1);
    fprintf(stderr, "[umma prof pair0] tiles %lld total %.0f cyc (%.0f/tile) | MMA: waitX %.1f%% waitH1 %.1f%% waitL2E %.1f%% "
            "issueL1 %.1f%% issueL2 %.1f%% | producer wait %.1f%% work %.1f%% | ep1 wait %.1f%% work %.1f%% | "
            "ep2 wait %.1f%% work %.1f%% || ep1 ld+wait %.1f%% thru-st-issue %.1f%% wait::st %.1f%% | ep2 first ld+wait %.1f%% pairbar %.1f%%\n", o[6], T, T / nt,
            100 * o[1] / T, 100 * o[2] / T, 100 * o[3] / T,
            100 * o[4] / T, 100 * o[5] / T, 100 * o[8] / T, 100 * o[9] / T, 100 * o[16] / T, 100 * o[17] / T,
            100 * o[24] / T, 100 * o[25] / T, 100 * o[18] / T, 100 * o[19] / T, 100 * o[20] / T, 100 * o[26] / T, 100 * o[27] / T);
  }
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_umma_mode(const rlc_critic* c, int act_mode) {
  if (!critic_ok(c) || c->topology != RLC_TIN) return -1;
  PackGeom G;
  const int ts = umma_mode();
  if (!make_geom(c, G, ts)) return -1;
  if (!ts) return 0;
  GridPlan gp;
  if (act_mode == RLC_ACT_SHARED && grid_mode() && plan_grid(G, gp)) return 3;
  return 1;
}

extern "C" int rlc_umma_mode_prec(const rlc_critic* c, int act_mode, int precision) {
  if (precision != RLC_PREC_FP16X3 && precision != RLC_PREC_FP16C8) return rlc_umma_mode(c, act_mode);
  if (!critic_ok(c) || c->topology != RLC_TIN || act_mode != RLC_ACT_SHARED) return -1;
  PackGeom G;
  Grid3Plan gp;
  const int mode = g3_mode_of(precision);
  return (make_geom3(c, G, mode) && plan_grid3(G, 232448, gp, mode)) ? (mode == G3_C8 ? 5 : 4) : -1;
}

// Debug/diagnostic: last error flag raised by a bounded wait inside the kernel (0 = none).
extern "C" int rlc_umma_last_error(rlc_handle* h, void* stream) {
  if (!h || !h->err_flag) return 0;
  int v = 0;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaMemcpy(&v, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v != 0) cudaMemset(h->err_flag, 0, sizeof(int));
  return v;
}
