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

#include "common.cuh"

#define UM_THREADS 384
#define UM_NST 3           // H1 ring stages
#define UM_MAXCH 8         // max layer-1 chunks
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
  int nch;             // number of layer-1 chunks
  int ch0[UM_MAXCH];   // first feature of chunk
  int chw[UM_MAXCH];   // width of chunk (multiple of 16, <= CH)
  int CH;              // max chunk width == TMEM columns per L1 buffer
  // packed blob for each cta rank (device), and byte offsets inside it
  const unsigned char* blob[2];
  int off_w2, off_w1, off_w3, off_nb2, off_c0, blob_bytes;
  // smem carve (bytes from the 1024-aligned base)
  int sm_w2, sm_w1, sm_x, sm_h1, sm_par, sm_bar, x_stage_bytes, h1_stage_bytes;
  int num_pair_tiles;
  int* err;     // device error flag (0 = ok)
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
      const uint32_t id1_last = um::make_idesc(fmt, 256, P.chw[nch - 1]);
      const int k1steps = P.K1P / 16;
      const int CH = P.CH, last_w = P.chw[nch - 1];
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
    const int CH = P.CH, last_w = P.chw[nch - 1];
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
// Weight pre-pack: theta (fp32 canonical) -> per-rank blobs in the UMMA core-matrix layout
// ------------------------------------------------------------------------------------------
struct PackGeom {
  int S, A, H1, H2, K1P, H1P, H2P, NA, NB, nch, CH;
  int ch0[UM_MAXCH], chw[UM_MAXCH];
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
    const float v = (k < G.H1 && n < G.H2) ? W2[(long long)k * G.H2 + n] : 0.f;
    unsigned char* blob = rank ? blob1 : blob0;
    const long long off = (long long)G.off_w2 + (long long)(k / 8) * ((G.H2P / 2) * 16) +
                          (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<unsigned short*>(blob + off) = to_h<PREC>(v);
  } else if (gid < nW2 + nW1) {
    const long long g = gid - nW2;
    const int k = (int)(g / G.H1P), j = (int)(g % G.H1P);
    int c = 0;
    while (c + 1 < G.nch && j >= G.ch0[c + 1]) ++c;
    const int within = j - G.ch0[c], hw = G.chw[c] / 2;
    const int rank = within / hw;
    const int l = G.ch0[c] / 2 + within % hw;
    float v = 0.f;
    if (j < G.H1) {
      if (k < K1) v = W1[(long long)k * G.H1 + j];
      else if (k == K1) v = b1[j];
    }
    unsigned char* blob = rank ? blob1 : blob0;
    const long long off = (long long)G.off_w1 + (long long)(k / 8) * ((G.H1P / 2) * 16) +
                          (long long)(l / 8) * 128 + (l % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<unsigned short*>(blob + off) = to_h<PREC>(v);
  } else if (gid < nW2 + nW1 + G.H2P) {
    const int j = (int)(gid - nW2 - nW1);
    const float w = (j < G.H2) ? w3[j] : 0.f;
    const float nb = (j < G.H2) ? -b2[j] : 0.f;
    for (int r = 0; r < 2; ++r) {
      unsigned char* blob = r ? blob1 : blob0;
      reinterpret_cast<float*>(blob + G.off_w3)[j] = w;
      reinterpret_cast<float*>(blob + G.off_nb2)[j] = nb;
    }
  } else if (gid == nW2 + nW1 + G.H2P) {
    float c0 = b3[0];
    for (int j = 0; j < G.H2; ++j) c0 = fmaf(w3[j], b2[j], c0);
    *reinterpret_cast<float*>(blob0 + G.off_c0) = c0;
    *reinterpret_cast<float*>(blob1 + G.off_c0) = c0;
  }
}

static bool make_geom(const rlc_critic* c, PackGeom& G) {
  memset(&G, 0, sizeof(G));
  G.S = c->S; G.A = c->A; G.H1 = c->H1; G.H2 = c->H2;
  G.K1P = (c->S + c->A + 1 + 15) & ~15;
  G.H1P = (c->H1 + 15) & ~15;
  G.H2P = (c->H2 + 15) & ~15;
  if (G.K1P > 64 || G.H2P > 480 || G.H2P < 32) return false;
  // layer-2 N split: one instruction if <= 256 else two halves rounded to 16
  if (G.H2P <= 256) { G.NA = G.H2P; G.NB = 0; }
  else { G.NA = ((G.H2P / 2) + 15) & ~15; G.NB = G.H2P - G.NA; }
  if (G.NA > 256 || G.NB > 256) return false;
  int ch = ((512 - G.H2P) / 2) & ~15;
  if (ch > 80) ch = 80;
  if (ch < 16) return false;
  G.CH = ch;
  G.nch = (G.H1P + ch - 1) / ch;
  if (G.nch > UM_MAXCH) return false;
  for (int i = 0; i < G.nch; ++i) {
    G.ch0[i] = i * ch;
    G.chw[i] = (G.H1P - i * ch < ch) ? (G.H1P - i * ch) : ch;
  }
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
  p.h1_stage = (G.CH / 8) * 2048;
  p.sm_bar = p.sm_h1 + UM_NST * p.h1_stage;
  p.total = p.sm_bar + BAR_COUNT * 8 + 16 + 1024;  // + alignment slack
  return p;
}

bool rlc_umma_supported(const rlc_handle* h, const rlc_critic* c, int B, int N) {
  if (h->sm_major != 10 || c->topology != RLC_TIN) return false;
  PackGeom G;
  if (!make_geom(c, G)) return false;
  const SmemPlan p = plan_smem(G);
  if ((size_t)p.total > h->smem_optin) return false;
  (void)B; (void)N;
  return true;
}

static int get_pack(rlc_handle* h, const rlc_critic* c, int prec, const PackGeom& G,
                    cudaStream_t st, rlc_pack** out) {
  rlc_pack* slot = nullptr;
  for (int i = 0; i < RLC_MAX_PACKS; ++i) {
    rlc_pack& p = h->packs[i];
    if (p.theta == c->theta && p.prec == prec && p.S == c->S && p.A == c->A && p.H1 == c->H1 &&
        p.H2 == c->H2 && p.dev) { slot = &p; break; }
  }
  if (!slot) {
    slot = &h->packs[h->pack_rr];
    h->pack_rr = (h->pack_rr + 1) % RLC_MAX_PACKS;
    if (slot->dev && slot->bytes < (size_t)2 * G.blob_bytes) {
      RLC_CUDA(cudaDeviceSynchronize());
      cudaFree(slot->dev);
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
    slot->S = c->S; slot->A = c->A; slot->H1 = c->H1; slot->H2 = c->H2;
    slot->valid = false;
  }
  if (!slot->valid) {
    unsigned char* b0 = (unsigned char*)slot->dev;
    unsigned char* b1 = b0 + G.blob_bytes;
    const long long n = (long long)G.H1P * G.H2P + (long long)G.K1P * G.H1P + G.H2P + 1;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (prec == RLC_PREC_BF16) k_pack_umma<RLC_PREC_BF16><<<blocks, 256, 0, st>>>(c->theta, G, b0, b1);
    else k_pack_umma<RLC_PREC_FP16><<<blocks, 256, 0, st>>>(c->theta, G, b0, b1);
    RLC_LAUNCH_CHECK(h);
    slot->valid = true;
  }
  *out = slot;
  return RLC_OK;
}

int rlc_eval_umma(rlc_handle* h, const rlc_critic* c, const float* s, int B, const float* a, int N,
                  int act_mode, int prec, float* q_out, cudaStream_t st) {
  PackGeom G;
  if (!make_geom(c, G)) return RLC_ERR_UNSUPPORTED;
  if ((long long)B * N >= (1ll << 31)) return RLC_ERR_UNSUPPORTED;  // 32-bit row arithmetic in the producers
  const SmemPlan sp = plan_smem(G);
  if ((size_t)sp.total > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  rlc_pack* pk = nullptr;
  int rc = get_pack(h, c, prec, G, st, &pk);
  if (rc) return rc;
  int* err = h->err_flag;

  UmmaParams P;
  memset(&P, 0, sizeof(P));
  P.s = s; P.a = a; P.smin = c->smin; P.smax = c->smax; P.q = q_out;
  P.R = (long long)B * N; P.N = N; P.S = c->S; P.A = c->A;
  P.act_per_state = act_mode == RLC_ACT_PER_STATE;
  P.K1P = G.K1P; P.KC1 = G.K1P / 8; P.H1P = G.H1P; P.KC2 = G.H1P / 8; P.H2P = G.H2P;
  P.NA = G.NA; P.NB = G.NB; P.nch = G.nch; P.CH = G.CH;
  for (int i = 0; i < UM_MAXCH; ++i) { P.ch0[i] = G.ch0[i]; P.chw[i] = G.chw[i]; }
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
  cfg.blockDim = dim3(UM_THREADS);
  cfg.dynamicSmemBytes = (size_t)sp.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (prec == RLC_PREC_BF16) {
    RLC_CUDA(cudaFuncSetAttribute(k_critic_umma<RLC_PREC_BF16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
    RLC_CUDA(cudaLaunchKernelEx(&cfg, k_critic_umma<RLC_PREC_BF16>, P));
  } else {
    RLC_CUDA(cudaFuncSetAttribute(k_critic_umma<RLC_PREC_FP16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
    RLC_CUDA(cudaLaunchKernelEx(&cfg, k_critic_umma<RLC_PREC_FP16>, P));
  }
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
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
