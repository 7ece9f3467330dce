// Tensor-core GEMM of the B-row training path (sm_100a, tcgen05): the dense contractions of
//   critic regression      forwardkl_network.py:133-140,199-201 (q_net forward/backward on the minibatch rows)
//   value / policy nets    forwardkl_network.py:143-158,196-209
//   dQ/da on stacked rows  ae_network.py:117,352-358, sql_network.py:101-107
// at minibatch sizes where they ARE dense GEMMs (B >= ~1k rows x 400 x 300).
//
// Arithmetic: the reference computes these in fp32, and the post-update weights are held to 5e-6 of the reference's
// (tests/test_gpu_parity.py::test_critic_step_matches_reference_update), so the operands are NOT rounded to 11 bits once:
// every fp32 operand x is split into tf32 parts  hi = rna_tf32(x), lo = rna_tf32(x - hi)  and the product is
//   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo          (A_lo.B_lo ~ 2^-22 relative: dropped)
// three tcgen05.mma kind::tf32 per K step, fp32 accumulation in tensor memory: fp32-class results (<= ~1e-6 relative)
// with the full fp32 exponent range (gradients of 1e-7 and activations of 1e+3 need no scaling, which an fp16 split would).
//
// Structure (one CTA = one 128 x 128 output tile of one K slice; cta_group::1):
//   * the operands live in global memory as fp32 in whatever orientation the caller has (row-major A or A^T, B or B^T), so
//     they cannot be bulk-copied: two groups of 8 producer warps load them (coalesced along the contiguous dimension),
//     split them and store both parts in the UMMA no-swizzle K-major core-matrix layout ([16-byte K chunk][row][4 x
//     tf32]) -- the transpose of a row-contiguous operand costs nothing, it is only a different register -> shared-memory
//     mapping; the groups take the ring steps in turn, so one group's loads fly while the other splits and stores;
//   * 3-stage shared-memory ring (64 KB per stage: A_hi, A_lo, B_hi, B_lo of a 32-deep K step), full/empty mbarriers;
//   * one elected thread issues the 12 MMAs of a stage and commits the stage back to the producers;
//   * the producer warps then drain the 128 x 128 fp32 accumulator (tcgen05.ld), turn it through shared memory so that
//     lanes walk the columns (256-byte row segments per store instead of 32 scattered sectors) and apply the fused
//     epilogue (alpha, bias, ReLU-derivative mask) on the way to global memory;
//   * split-K over gridDim.z for the weight gradients (K = batch rows): every slice writes its own slab, the slabs are
//     summed in a fixed order by the caller (deterministic).
// Bounded mbarrier waits raise the handle's error flag instead of hanging.  Non-finite operands: inf splits into
// (inf, inf - inf = NaN), so an infinite input yields NaN where the fp32 FMA kernel would yield inf.
#include "rows_gemm_tc.cuh"

#include <stdlib.h>

#define GT_BM 128
#define GT_BN 128
#define GT_BK 32
#define GT_STAGES 3
#define GT_GROUPS 2            // producer groups of 8 warps; group g fills the ring steps it = g, g + GT_GROUPS, ...
#define GT_PROD_WARPS (8 * GT_GROUPS)
#define GT_THREADS (32 * (GT_PROD_WARPS + 1))
#define GT_EPI_WARPS 16                             // warps 1..16 drain the accumulator
#define GT_EPI_COLS (GT_BN / (GT_EPI_WARPS / 4))    // accumulator columns per epilogue warp
#define GT_EPI_LD (GT_EPI_COLS + 4)                 // staging row pitch in words: conflict-free both ways
#define GT_OP_BYTES (128 * GT_BK * 4)
#define GT_STAGE_BYTES (4 * GT_OP_BYTES)
#define GT_LBO (128 * 16)
#define GT_WAIT_LIMIT (1u << 24)
#define GT_SMEM_BYTES (GT_STAGES * GT_STAGE_BYTES + 1024 + 256)

struct GemmTcParams {
  const float *A, *B;
  float* C;
  const float *bias, *maskZ;
  int M, N, K, lda, ldb, ldc, ldz, reluA;
  float alpha;
  int klen;
  long long cz_stride;
  int* err;
  int a_vec, b_vec, c_vec, z_vec;  // 16-byte accesses allowed
  int micro;                       // RLC_GEMM_TC_MICRO (timing decomposition, results invalid): 1 = producers only signal,
                                   // 2 = no MMAs issued, 4 = no global loads (registers stay zero)
  int a_ones;                      // row of op(A) that is all ones instead of memory (-1: none); it is row M-1
};

namespace gt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > GT_WAIT_LIMIT) {
      atomicCAS(err, 0, code);
      return false;
    }
    if ((spins & 0xffff) == 0 && *(volatile int*)err != 0) return false;
  }
  return true;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 128 x N fp32] (+)= A[smem] * B[smem]^T, tf32 operands
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// c_format F32 (1) @4 ; a_format, b_format TF32 (2) @7, @10 ; both K-major ; N>>3 @17 ; M>>4 @24
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// no-swizzle K-major descriptor: start address, LBO (bytes between 16-byte K chunks), SBO = 128 B between 8-row groups,
// descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((uint32_t)(GT_LBO >> 4) & 0x3FFFu) << 16);
  constexpr uint32_t hi = (128u >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}

// Global loads as volatile asm: __ldg is an invariant load the compiler may (and did) sink to its first use, across the
// barrier waits -- which turns the two-register-set prefetch below back into load -> wait -> use.  Volatile asm
// statements keep their order relative to the mbarrier asm, so the loads of step it+1 are in flight during step it.
__device__ __forceinline__ float ldg1(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// fp32 -> (hi, lo) tf32 parts, both rounded to nearest (ties away, what cvt.rna.tf32.f32 does -- that instruction is a
// five-instruction emulation in SASS, the integer form below is two): x = hi + lo up to 2^-22 |x|
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  const float l = x - __uint_as_float(hi);
  lo = (__float_as_uint(l) + 0x1000u) & 0xffffe000u;
}

// One operand tile of a stage: 128 rows x 32 k = 1024 (row, 4-k chunk) items of 16 bytes, four per producer thread.
//   row-contiguous global layout (RC): lanes walk the rows, a thread owns row pt & 127 and chunks (pt >> 7) + 2i --
//     every LDG.32 of a warp is one 128-byte line.
//   k-contiguous global layout (!RC): a row's 32 k are ONE line, so a warp must not put 32 rows into one load (ncu: the
//     first mapping, a row per lane, spent 2 100 cycles per stage in the L1 wavefront queue at 32 lines per LDG.128).
//     Warp w owns rows 16w..16w+15; a quarter-warp reads 8 rows along a DIAGONAL of chunks -- lane l: row (l & 7) [+8],
//     chunk ((l & 7) + (l >> 3) [+4]) & 7 -- so an LDG.128 touches 8 lines, and its 8 lanes of a quarter-warp still hit 8
//     different rows = 8 different 16-byte bank groups of the core-matrix layout (conflict-free STS.128).
template <bool RC>
struct OpMap {
  const float* rp[2];   // !RC: &G[row_h][kbeg]   RC: rp[0] = &G[kbeg][row]
  long long ld;
  int koff[4];          // k offset of item i inside a stage (4 * chunk)
  int smoff[4];         // byte offset of item i inside an operand part
  bool rok[2], ones[2], vec;

  __device__ __forceinline__ void init(const float* G, int ld_, int row0, int rows, int kbeg, int vec_, int ones_row,
                                       int pt) {
    ld = ld_;
    vec = vec_ != 0;
    if (RC) {
      const int rl = pt & 127;
      rok[0] = rok[1] = row0 + rl < rows;
      ones[0] = ones[1] = row0 + rl == ones_row;
      rp[0] = rp[1] = G + (long long)kbeg * ld + row0 + rl;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (pt >> 7) + 2 * i;
        koff[i] = 4 * c;
        smoff[i] = c * GT_LBO + rl * 16;
      }
    } else {
      const int w = pt >> 5, l = pt & 31;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rl = 16 * w + 8 * h + (l & 7);
        rok[h] = row0 + rl < rows;
        ones[h] = row0 + rl == ones_row;
        rp[h] = G + (long long)(row0 + rl) * ld + kbeg;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int c = ((l & 7) + (l >> 3) + 4 * d) & 7;
          koff[2 * h + d] = 4 * c;
          smoff[2 * h + d] = c * GT_LBO + rl * 16;
        }
      }
    }
  }
  // stage `it` (k = kbeg + 32 it ..) into registers; kleft = kend - (kbeg + 32 it)
  __device__ __forceinline__ void load(int it, int kleft, float (&r)[16]) const {
    const bool full = kleft >= GT_BK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = RC ? 0 : (i >> 1);
      const int k = it * GT_BK + koff[i];
      if (RC) {
        const float* p = rp[0] + (long long)k * ld;
        if (rok[0] && !ones[0] && full) {
#pragma unroll
          for (int e = 0; e < 4; ++e) r[4 * i + e] = ldg1(p + e * ld);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            r[4 * i + e] = (rok[0] && koff[i] + e < kleft) ? (ones[0] ? 1.f : ldg1(p + e * ld)) : 0.f;
        }
      } else {
        const float* p = rp[h] + k;
        if (rok[h] && !ones[h] && full && vec) {
          const float4 v = ldg4(p);
          r[4 * i + 0] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            r[4 * i + e] = (rok[h] && koff[i] + e < kleft) ? (ones[h] ? 1.f : ldg1(p + e)) : 0.f;
        }
      }
    }
  }
  __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo, const float (&r)[16], int relu) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 h4, l4;
      float x0 = r[4 * i + 0], x1 = r[4 * i + 1], x2 = r[4 * i + 2], x3 = r[4 * i + 3];
      if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
      split(x0, h4.x, l4.x);
      split(x1, h4.y, l4.y);
      split(x2, h4.z, l4.z);
      split(x3, h4.w, l4.w);
      *reinterpret_cast<uint4*>(hi + smoff[i]) = h4;
      *reinterpret_cast<uint4*>(lo + smoff[i]) = l4;
    }
  }
};

}  // namespace gt

template <bool A_RC, bool B_RC>
__global__ void __launch_bounds__(GT_THREADS, 1) k_gemm_tc(const GemmTcParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = gt::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* bp = smem_raw + (base - raw_addr);
  const uint32_t sBar = base + GT_STAGES * GT_STAGE_BYTES;
  auto full_bar = [&](uint32_t s) { return sBar + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return sBar + 8u * (GT_STAGES + s); };
  const uint32_t acc_bar = sBar + 8u * (2 * GT_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + GT_STAGES * GT_STAGE_BYTES + 8 * (2 * GT_STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * GT_BN;   // row tiles on x: no 65 535 limit on the batch dimension
  const int kbeg = blockIdx.z * P.klen;
  const int kend = (kbeg + P.klen < P.K) ? kbeg + P.klen : P.K;
  const int nk = (kend > kbeg) ? (kend - kbeg + GT_BK - 1) / GT_BK : 0;
  float* __restrict__ C = P.C + (long long)blockIdx.z * P.cz_stride;

  if (tid == 0) {
    for (uint32_t s = 0; s < GT_STAGES; ++s) {
      gt::mbar_init(full_bar(s), 8);
      gt::mbar_init(empty_bar(s), 1);
    }
    gt::mbar_init(acc_bar, 1);
    gt::fence_mbar_init();
  }
  if (warp == 0) gt::tmem_alloc(gt::smem_u32(tmem_slot), 2 * GT_BN);   // main accumulator + correction accumulator
  gt::tc_fence_before();
  __syncthreads();
  gt::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================================= MMA issuer =======================================
    if (gt::elect_one()) {
      const uint32_t idesc = gt::idesc_tf32(GT_BM, GT_BN);
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (int it = 0; it < nk; ++it) {
        if (!gt::mbar_wait(full_bar(s), ph, P.err, 201)) { ok = false; break; }
        // The generic-proxy stores of the producers reach this thread through the barrier (release arrive / acquire wait);
        // the proxy fence towards the tensor core's async-proxy reads sits HERE, on that causality path, and not in the
        // producers: there it also waits for the producer's own prefetch loads of the next step (ncu: long-scoreboard
        // stall on FENCE.VIEW.ASYNC) and serialises load latency with the split arithmetic.
        gt::fence_proxy_async();
        gt::tc_fence_after();
        const uint32_t a_hi = base + s * GT_STAGE_BYTES, a_lo = a_hi + GT_OP_BYTES, b_hi = a_hi + 2 * GT_OP_BYTES,
                       b_lo = a_hi + 3 * GT_OP_BYTES;
        const int rem = kend - (kbeg + it * GT_BK);
        const int ks = rem >= GT_BK ? GT_BK / 8 : (rem + 7) / 8;
        for (int k = 0; k < ((P.micro & 2) ? 0 : ks); ++k) {
          const uint32_t off = (uint32_t)k * 2u * GT_LBO;  // 8 tf32 = two 16-byte K chunks per MMA
          const uint64_t dah = gt::desc(a_hi + off), dal = gt::desc(a_lo + off), dbh = gt::desc(b_hi + off),
                         dbl = gt::desc(b_lo + off);
          // The two correction products go to their OWN accumulator (columns GT_BN..): the tensor core's fp32 accumulate
          // step truncates, which biases a long K sum by ~0.5 ulp of the running sum per MMA (measured: 5x the rms error
          // of an fp32 FMA chain at K = 400 with all three products in one accumulator); the corrections are 2^-11 of
          // the main sum, so their own truncation is invisible and the main accumulator sees a third of the steps.
          gt::mma_tf32(tmem_base, dah, dbh, idesc, (uint32_t)((it | k) != 0));
          gt::mma_tf32(tmem_base + GT_BN, dal, dbh, idesc, (uint32_t)((it | k) != 0));
          gt::mma_tf32(tmem_base + GT_BN, dah, dbl, idesc, 1u);
        }
        gt::commit(empty_bar(s));
        if (++s == GT_STAGES) { s = 0; ph ^= 1u; }
      }
      if (ok && nk > 0) gt::commit(acc_bar);
    }
    __syncwarp();
  } else {
    // ======================================= producers =======================================
    // Two groups of 8 warps take the ring steps in turn: while one group splits and stores its step, the global loads
    // of the other group's step are in flight.  (Measured decomposition, RLC_GEMM_TC_MICRO: per 32-deep step the L2
    // loads, the split + shared-memory stores and the 12 MMAs cost 0.6 / 0.46 / 0.53 us each; one group doing
    // load -> split -> store in sequence ran at their SUM.)
    const int pt = (tid - 32) & 255, grp = (tid - 32) >> 8;
    bool ok = true;
    gt::OpMap<A_RC> ma;
    gt::OpMap<B_RC> mb;
    ma.init(P.A, P.lda, m0, P.M, kbeg, P.a_vec, P.a_ones, pt);
    mb.init(P.B, P.ldb, n0, P.N, kbeg, P.b_vec, -1, pt);
    float ra[16] = {}, rb[16] = {};
    for (int it = grp; it < nk; it += GT_GROUPS) {
      const int kleft = kend - (kbeg + it * GT_BK);
      if (!(P.micro & 4)) {
        ma.load(it, kleft, ra);
        mb.load(it, kleft, rb);
      }
      const uint32_t s = (uint32_t)(it % GT_STAGES), n = (uint32_t)(it / GT_STAGES);
      if (n > 0) {
        if (!gt::mbar_wait(empty_bar(s), (n - 1u) & 1u, P.err, 202)) { ok = false; break; }
      }
      unsigned char* st = bp + s * GT_STAGE_BYTES;
      if (!(P.micro & 1)) {
        ma.store(st, st + GT_OP_BYTES, ra, P.reluA);
        mb.store(st + 2 * GT_OP_BYTES, st + 3 * GT_OP_BYTES, rb, 0);
      }
      __syncwarp();
      if (lane == 0) gt::mbar_arrive(full_bar(s));
    }
    // ======================================= epilogue =======================================
    // accumulator quarter -> registers -> this warp's staging block in the (now idle) operand ring -> coalesced rows
    if (ok && nk > 0) ok = gt::mbar_wait(acc_bar, 0u, P.err, 203);
    if (ok && warp <= GT_EPI_WARPS) {
      gt::tc_fence_after();
      const int q = warp & 3;              // TMEM lane quarter this warp may read
      const int cbase = ((warp - 1) >> 2) * GT_EPI_COLS;   // this warp's column block
      if (n0 + cbase < P.N) {              // warp-uniform
        float* tb = reinterpret_cast<float*>(bp + (warp - 1) * (32 * GT_EPI_LD * 4));
        {
          uint32_t v[GT_EPI_COLS];
          if (nk > 0) {
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase;
#pragma unroll
            for (int g = 0; g < GT_EPI_COLS / 16; ++g) gt::tmem_ld16(ta + 16u * g, v + 16 * g);
            uint32_t c[GT_EPI_COLS];
#pragma unroll
            for (int g = 0; g < GT_EPI_COLS / 16; ++g) gt::tmem_ld16(ta + GT_BN + 16u * g, c + 16 * g);
            gt::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < GT_EPI_COLS; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(c[j]));
          } else {
#pragma unroll
            for (int j = 0; j < GT_EPI_COLS; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int g = 0; g < GT_EPI_COLS / 4; ++g)
            *reinterpret_cast<uint4*>(tb + lane * GT_EPI_LD + 4 * g) =
                make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        __syncwarp();
        constexpr int LPR = GT_EPI_COLS / 4;   // lanes per row in the read phase
        const int c4 = lane % LPR;
        const int n = n0 + cbase + 4 * c4;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (P.bias) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < P.N) bv[e] = __ldg(P.bias + n + e);
        }
        if (n < P.N) {
#pragma unroll 4
          for (int rr = 0; rr < LPR; ++rr) {
            const int r = (32 / LPR) * rr + lane / LPR;
            const int row = m0 + q * 32 + r;
            if (row >= P.M) continue;
            const float4 t = *reinterpret_cast<const float4*>(tb + r * GT_EPI_LD + 4 * c4);
            float o[4] = {fmaf(t.x, P.alpha, bv[0]), fmaf(t.y, P.alpha, bv[1]), fmaf(t.z, P.alpha, bv[2]),
                          fmaf(t.w, P.alpha, bv[3])};
            float* crow = C + (long long)row * P.ldc + n;
            const float* zrow = P.maskZ ? P.maskZ + (long long)row * P.ldz + n : nullptr;
            if (n + 4 <= P.N) {
              if (zrow) {
                if (P.z_vec) {
                  const float4 z = __ldg(reinterpret_cast<const float4*>(zrow));
                  if (!(z.x > 0.f)) o[0] = 0.f;
                  if (!(z.y > 0.f)) o[1] = 0.f;
                  if (!(z.z > 0.f)) o[2] = 0.f;
                  if (!(z.w > 0.f)) o[3] = 0.f;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (!(__ldg(zrow + e) > 0.f)) o[e] = 0.f;
                }
              }
              if (P.c_vec) {
                *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) crow[e] = o[e];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (n + e < P.N) {
                  float x = o[e];
                  if (zrow && !(__ldg(zrow + e) > 0.f)) x = 0.f;
                  crow[e] = x;
                }
              }
            }
          }
        }
      }
    }
  }
  gt::tc_fence_before();
  __syncthreads();
  if (warp == 0) gt::tmem_dealloc(tmem_base, 2 * GT_BN);
}

// ---------------------------------------------------------------------------------------------------------------
static int gemm_tc_mode() {
  static int mode = -1;  // 0 = off, 1 = auto (default), 2 = every shape (tests)
  if (mode < 0) {
    const char* e = getenv("RLC_GEMM_TC");
    mode = e ? atoi(e) : 1;
    if (mode < 0 || mode > 2) mode = 1;
  }
  return mode;
}

static thread_local int g_force = -1;
void rlc_gemm_tc_force(int mode) { g_force = mode; }
int rlc_gemm_tc_forced() { return g_force; }
extern "C" int rlc_rows_gemm_force(int mode) {
  const int prev = g_force;
  g_force = (mode == 0 || mode == 2) ? mode : -1;
  return prev;
}

bool rlc_gemm_tc_ok(const rlc_handle* h, int M, int N, int K) {
  const int mode = g_force >= 0 ? g_force : gemm_tc_mode();
  if (mode == 0 || !h || h->sm_major != 10) return false;
  if (M < 1 || N < 1 || K < 1) return false;
  if (mode == 2) return true;
  // Both kernels are latency-bound per CTA at these sizes (a launch of this one costs ~6 us + 0.83 us per 32-deep K step
  // whatever M is; the CUDA-core tile kernel 23 us at 128..1024 x 300 x 400): the tensor path wins from ~128 rows of a
  // 400-300 layer on (scripts/time_rows_gemm.py B); below that the small launches stay on the CUDA cores
  return N >= 32 && K >= 8 && (double)M * N * K >= (double)(1 << 23);
}

void rlc_gemm_tc_splitk_plan(const rlc_handle* h, int M, int N, int K, int max_slabs, int* nz_out, int* klen_out) {
  const int tiles = ((M + GT_BM - 1) / GT_BM) * ((N + GT_BN - 1) / GT_BN);
  int nz = (h->num_sms > 0 ? h->num_sms : 148) / (tiles > 0 ? tiles : 1);
  if (nz > max_slabs) nz = max_slabs;
  if (nz > K / 64) nz = K / 64;   // at least two 32-deep steps per slice
  if (nz < 1) nz = 1;
  int klen = (K + nz - 1) / nz;
  klen = (klen + GT_BK - 1) / GT_BK * GT_BK;
  nz = (K + klen - 1) / klen;
  *nz_out = nz;
  *klen_out = klen;
}

int rlc_gemm_tc(rlc_handle* h, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                float* C, int ldc, GemmEpi epi, int nz, int klen, long long cz_stride, cudaStream_t st) {
  if (M == 0 || N == 0) return RLC_OK;
  RLC_REQUIRE(nz >= 1 && (nz == 1 || klen % GT_BK == 0));
  GemmTcParams P;
  P.A = A; P.B = Bm; P.C = C; P.bias = epi.bias; P.maskZ = epi.maskZ;
  P.M = M; P.N = N; P.K = K; P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.ldz = epi.ldz; P.reluA = epi.reluA;
  P.alpha = epi.alpha;
  P.a_ones = epi.onesA ? M - 1 : -1;
  {
    static int micro = -1;
    if (micro < 0) {
      const char* e = getenv("RLC_GEMM_TC_MICRO");
      micro = e ? atoi(e) : 0;
    }
    P.micro = micro;
  }
  P.klen = nz == 1 ? (K > 0 ? K : 1) : klen;
  P.cz_stride = cz_stride;
  P.err = h->err_flag;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  P.a_vec = (!ta && lda % 4 == 0 && al16(A)) ? 1 : 0;
  P.b_vec = (tb && ldb % 4 == 0 && al16(Bm)) ? 1 : 0;
  P.c_vec = (ldc % 4 == 0 && al16(C) && cz_stride % 4 == 0) ? 1 : 0;
  P.z_vec = (epi.maskZ && epi.ldz % 4 == 0 && al16(epi.maskZ)) ? 1 : 0;
  dim3 grid((M + GT_BM - 1) / GT_BM, (N + GT_BN - 1) / GT_BN, nz);
  RLC_REQUIRE(grid.y <= 65535u && grid.z <= 65535u);
  // operand orientation in global memory: A is row-contiguous (element (m,k) at A[k*lda+m]) when ta;
  // B (rows = n) is row-contiguous (element (n,k) at B[k*ldb+n]) when NOT tb
  const bool a_rc = ta, b_rc = !tb;
#define GT_LAUNCH(ARC, BRC)                                                                                  \
  do {                                                                                                       \
    RLC_CUDA(cudaFuncSetAttribute(k_gemm_tc<ARC, BRC>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  GT_SMEM_BYTES));                                                           \
    k_gemm_tc<ARC, BRC><<<grid, GT_THREADS, GT_SMEM_BYTES, st>>>(P);                                         \
  } while (0)
  if (!a_rc && !b_rc) GT_LAUNCH(false, false);
  else if (a_rc && !b_rc) GT_LAUNCH(true, false);
  else if (!a_rc && b_rc) GT_LAUNCH(false, true);
  else GT_LAUNCH(true, true);
#undef GT_LAUNCH
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
