// Interface of the tensor-core GEMM of the B-row training path (rows_gemm_tc.cu) towards rows_gemm.cuh.
#pragma once
#include "common.cuh"

// C[M,N] = alpha * opA(A)[M,K] * opB(B)[K,N] (+ bias[N]) ; optional relu on A at load (A := relu(A)),
// optional mask C *= (Z > 0).  Row-major with leading dimensions.
struct GemmEpi {
  const float* bias;   // per output column or nullptr
  const float* maskZ;  // same shape/ld as C or nullptr: C *= (Z>0)
  int ldz;
  int reluA;
  float alpha;
  int onesA = 0;       // tensor path only: row M-1 of op(A) is all ones, not memory ([X | 1]^T G = [dW ; db])
};

// true when this shape is worth the tcgen05 path on this device (sm_100, enough work per launch, RLC_GEMM_TC != 0)
bool rlc_gemm_tc_ok(const rlc_handle* h, int M, int N, int K);
// Same contract as gemm_z() in rows_gemm.cuh; klen must be a multiple of 32 when nz > 1.
int rlc_gemm_tc(rlc_handle* h, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                float* C, int ldc, GemmEpi epi, int nz, int klen, long long cz_stride, cudaStream_t st);
// split-K plan of the tensor-core path for a weight-gradient GEMM: slices and slice length (multiple of 32)
void rlc_gemm_tc_splitk_plan(const rlc_handle* h, int M, int N, int K, int max_slabs, int* nz, int* klen);
// per-thread override of the dispatcher for the next calls: -1 = none, 0 = CUDA cores, 2 = tensor cores whatever the shape
void rlc_gemm_tc_force(int mode);
int rlc_gemm_tc_forced();
