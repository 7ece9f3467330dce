// Tiled fp32 SGEMM, split-K weight-gradient GEMM and deterministic column reductions shared by the
// B-row training paths (critic_fp32.cu: critic regression / dQ/da; mlp_rows.cu: value and policy nets).
#pragma once
#include "common.cuh"
#include "rows_gemm_tc.cuh"

// =============================================================================================
// Tiled SGEMM used by the B-row training path and the T-in dQ/da:
//   C[M,N] = opA(A)[M,K] * opB(B)[K,N]  (+ bias[N]) ; optional relu on A at load (A := relu(A)),
//   optional mask multiply C *= (Z > 0).  Row-major with leading dimensions (GemmEpi: rows_gemm_tc.cuh).
// Launches that are dense contractions (from ~128 rows of a 400-300 layer) go to the tcgen05 kernel of rows_gemm_tc.cu
// (3 x TF32 split, fp32-class results); this CUDA-core kernel keeps the latency-bound small ones.
// =============================================================================================
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
k_gemm(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ Bm,
       int ldb, float* __restrict__ C, int ldc, GemmEpi epi, int klen, long long cz_stride) {
  // 64 x 64 output tile, 16-deep k steps, double-buffered shared memory: the global loads of step
  // k+1 are issued before the FMAs of step k and parked in registers, so their latency overlaps the
  // math (the B-row training GEMMs are latency-bound: a few CTAs, a dozen dependent k steps).
  __shared__ float As[2][16][64 + 4];
  __shared__ float Bs[2][16][64 + 4];
  // split-K: slice blockIdx.z covers K range [z*klen, min(K,(z+1)*klen)) and writes its own C slab
  const int kbeg = blockIdx.z * klen;
  K = (kbeg + klen < K) ? kbeg + klen : K;
  C += (long long)blockIdx.z * cz_stride;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tr = tid >> 4, tc = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = tid + l * 256;  // 0..1023
      {
        int m, k;
        if (TA) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.f;
        if (gm < M && gk < K) v = TA ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
        if (epi.reluA) v = fmaxf(v, 0.f);
        ra[l] = v;
      }
      {
        int n, k;
        if (TB) { k = e & 15; n = e >> 4; } else { n = e & 63; k = e >> 6; }
        const int gn = n0 + n, gk = k0 + k;
        float v = 0.f;
        if (gn < N && gk < K) v = TB ? Bm[(long long)gn * ldb + gk] : Bm[(long long)gk * ldb + gn];
        rb[l] = v;
      }
    }
  };
  auto park = [&](int buf) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = tid + l * 256;
      if (TA) As[buf][e >> 6][e & 63] = ra[l]; else As[buf][e & 15][e >> 4] = ra[l];
      if (TB) Bs[buf][e & 15][e >> 4] = rb[l]; else Bs[buf][e >> 6][e & 63] = rb[l];
    }
  };
  if (kbeg < K) {
    fetch(kbeg);
    park(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < K; k0 += 16) {
    const bool more = k0 + 16 < K;
    if (more) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[buf][k][tr * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[buf][k][tc * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) park(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tr * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tc * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] * epi.alpha;
      if (epi.bias) v += epi.bias[gn];
      if (epi.maskZ && !(epi.maskZ[(long long)gm * epi.ldz + gn] > 0.f)) v = 0.f;
      C[(long long)gm * ldc + gn] = v;
    }
  }
}

static int gemm_z(rlc_handle* h, bool ta, bool tb, int M, int N, int K, const float* A, int lda,
                  const float* Bm, int ldb, float* C, int ldc, GemmEpi epi, int nz, int klen,
                  long long cz_stride, cudaStream_t st) {
  if (M == 0 || N == 0) return RLC_OK;
  if ((nz == 1 || klen % 32 == 0) && rlc_gemm_tc_ok(h, M, N, K))
    return rlc_gemm_tc(h, ta, tb, M, N, K, A, lda, Bm, ldb, C, ldc, epi, nz, klen, cz_stride, st);
  dim3 grid((N + 63) / 64, (M + 63) / 64, nz);
  if (!ta && !tb) k_gemm<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, Bm, ldb, C, ldc, epi, klen, cz_stride);
  else if (ta && !tb) k_gemm<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, Bm, ldb, C, ldc, epi, klen, cz_stride);
  else if (!ta && tb) k_gemm<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, Bm, ldb, C, ldc, epi, klen, cz_stride);
  else k_gemm<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, Bm, ldb, C, ldc, epi, klen, cz_stride);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

static int gemm(rlc_handle* h, bool ta, bool tb, int M, int N, int K, const float* A, int lda,
                const float* Bm, int ldb, float* C, int ldc, GemmEpi epi, cudaStream_t st) {
  return gemm_z(h, ta, tb, M, N, K, A, lda, Bm, ldb, C, ldc, epi, 1, K > 0 ? K : 1, 0, st);
}

// out[i] = sum_z part[z][i]  (fixed order: deterministic)
static __global__ void k_sum_slabs(const float* __restrict__ part, long long n, int nz, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int z = 0; z < nz; ++z) acc += part[(long long)z * n + i];
  out[i] = acc;
}

// Weight-gradient GEMM C[M,N] = op(A)^T B with a long reduction dimension K = batch rows and a small
// M x N: split K over gridDim.z so the grid fills the GPU, then sum the slabs in a fixed order.
#define SPLITK_MAX 32
static int gemm_splitk(rlc_handle* h, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                       float* C, GemmEpi epi, float* slabs, cudaStream_t st) {
  int nz, klen = K > 0 ? K : 1;
  if (slabs && rlc_gemm_tc_ok(h, M, N, K)) {
    rlc_gemm_tc_splitk_plan(h, M, N, K, SPLITK_MAX, &nz, &klen);   // slices sized to fill the SMs with 128 x 128 tiles
  } else {
    nz = (K + 127) / 128;
    if (nz > SPLITK_MAX) nz = SPLITK_MAX;
    if (nz > 1) {
      klen = (K + nz - 1) / nz;
      klen = (klen + 15) & ~15;
      nz = (K + klen - 1) / klen;
    }
  }
  if (nz <= 1 || !slabs) return gemm(h, true, false, M, N, K, A, lda, Bm, ldb, C, N, epi, st);
  int rc = gemm_z(h, true, false, M, N, K, A, lda, Bm, ldb, slabs, N, epi, nz, klen, (long long)M * N, st);
  if (rc) return rc;
  const long long n = (long long)M * N;
  k_sum_slabs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slabs, n, nz, C);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// forward declaration (defined below)
template <int MODE>
static int colred(rlc_handle* h, const float* Mx, const float* dq, long long R, int N, int ld, float* part,
                  float* out, cudaStream_t st);

// Weight AND bias gradient of one layer, contiguous in theta's layout: C[(M+1),N] = [op(A) | 1]^T G, i.e. rows 0..M-1 =
// A^T G (dW, [in,out]-major) and row M = column sums of G (db).  Dense launches: ONE split-K tensor-core GEMM whose
// producer supplies the ones row; small ones: the CUDA-core split-K GEMM + the two-stage column reduction.
static int gemm_splitk_bias(rlc_handle* h, int M, int N, int K, const float* A, int lda, const float* G, int ldg,
                            float* C, GemmEpi epi, float* slabs, float* part, cudaStream_t st) {
  if (slabs && rlc_gemm_tc_ok(h, M + 1, N, K)) {
    int nz, klen;
    rlc_gemm_tc_splitk_plan(h, M + 1, N, K, SPLITK_MAX, &nz, &klen);
    epi.onesA = 1;
    const long long n = (long long)(M + 1) * N;
    if (nz <= 1) return rlc_gemm_tc(h, true, false, M + 1, N, K, A, lda, G, ldg, C, N, epi, 1, K, 0, st);
    int rc = rlc_gemm_tc(h, true, false, M + 1, N, K, A, lda, G, ldg, slabs, N, epi, nz, klen, n, st);
    if (rc) return rc;
    k_sum_slabs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slabs, n, nz, C);
    RLC_LAUNCH_CHECK(h);
    return RLC_OK;
  }
  int rc = gemm_splitk(h, M, N, K, A, lda, G, ldg, C, epi, slabs, st);
  if (rc) return rc;
  return colred<0>(h, G, nullptr, K, N, ldg, part, C + (long long)M * N, st);
}

// Column reductions over the batch rows, two deterministic stages: stage 1 = one block per
// (128 columns x one row chunk) writing part[chunk][n]; stage 2 = k_sum_slabs over the chunks.
//   MODE 0: out[n] = sum_r M[r,n]
//   MODE 1: out[j] = sum_r relu(Z2[r,j]) dq[r] (j < N) ; out[N] = sum_r dq[r]     (head grads)
template <int MODE>
__global__ void __launch_bounds__(128)
k_colred_part(const float* __restrict__ Mx, const float* __restrict__ dq, long long R, int N, int ld,
              int rows_per_chunk, float* __restrict__ part) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int ncols = N + (MODE == 1 ? 1 : 0);
  if (n >= ncols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = (r0 + rows_per_chunk < R) ? r0 + rows_per_chunk : R;
  float acc = 0.f;
  if (MODE == 1 && n == N) {
    for (long long r = r0; r < r1; ++r) acc += dq[r];
  } else if (MODE == 1) {
    for (long long r = r0; r < r1; ++r) acc = fmaf(fmaxf(Mx[r * ld + n], 0.f), dq[r], acc);
  } else {
    for (long long r = r0; r < r1; ++r) acc += Mx[r * ld + n];
  }
  part[(long long)blockIdx.y * ncols + n] = acc;
}

#define COLRED_MAX_CHUNKS 256
template <int MODE>
static int colred(rlc_handle* h, const float* Mx, const float* dq, long long R, int N, int ld, float* part,
                  float* out, cudaStream_t st) {
  const int ncols = N + (MODE == 1 ? 1 : 0);
  int nchunks = (int)((R + 15) / 16);
  if (R <= 64) nchunks = 1;  // small batches are launch-latency-bound: one stage, straight into `out`
  if (nchunks > COLRED_MAX_CHUNKS) nchunks = COLRED_MAX_CHUNKS;
  if (nchunks < 1) nchunks = 1;
  const int rpc = (int)((R + nchunks - 1) / nchunks);
  nchunks = (int)((R + rpc - 1) / rpc);
  dim3 grid((ncols + 127) / 128, nchunks);
  k_colred_part<MODE><<<grid, 128, 0, st>>>(Mx, dq, R, N, ld, rpc, nchunks == 1 ? out : part);
  RLC_LAUNCH_CHECK(h);
  if (nchunks == 1) return RLC_OK;
  k_sum_slabs<<<(ncols + 255) / 256, 256, 0, st>>>(part, ncols, nchunks, out);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

