// K6: replay minibatch gather / scatter over a device-resident struct-of-arrays ring
// (utils/replaybuffer.py:25-37).  HBM-bound: algorithmic bytes = B*(2S+A+2)*4 read + the same
// written (+8 B per index).  Gather: one thread per output float (k_replay_gather_flat); scatter and the
// fallback gather: one warp per transition, every field row copied with coalesced (16-byte when aligned) loads.
#include "common.cuh"

__device__ __forceinline__ void copy_row(const float* __restrict__ src, float* __restrict__ dst,
                                         int n, int lane) {
  if ((n & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < (n >> 2); i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = __ldg(src + i);
  }
}

__global__ void __launch_bounds__(256)
k_replay_gather(const float* __restrict__ state, const float* __restrict__ action,
                const float* __restrict__ reward, const float* __restrict__ next_state,
                const float* __restrict__ gamma, long long cap, int S, int A,
                const long long* __restrict__ idx, int B, float* __restrict__ s_out,
                float* __restrict__ a_out, float* __restrict__ r_out, float* __restrict__ s2_out,
                float* __restrict__ g_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const long long j = idx[row];
  if (j < 0 || j >= cap) return;  // host wrapper validates indices; never read out of bounds
  copy_row(state + j * S, s_out + (long long)row * S, S, lane);
  copy_row(next_state + j * S, s2_out + (long long)row * S, S, lane);
  copy_row(action + j * A, a_out + (long long)row * A, A, lane);
  if (lane == 0) {
    r_out[row] = reward[j];
    g_out[row] = gamma[j];
  }
}

// Element-per-thread gather (default): a CTA owns `gr` consecutive sampled rows and walks their gr*(2S+A+2) output
// floats with one thread per float, so every lane of a load instruction is live (the warp-per-row kernel below keeps
// 17 + 17 + 6 + 2 of 4 x 32 lanes busy at S=17, A=6), the loads of a batch of UNROLL elements are all issued before the
// first store, and both the ring reads (within a field row) and the output writes are contiguous across lanes.
// Element -> (row, field offset) by a multiply-high with a host-computed reciprocal (exact for every i < 32 E when E <= 8192; checked exhaustively in tests/test_abi.py).
#define GATHER_THREADS 128
#define GATHER_MAX_ROWS 32
#define GATHER_UNROLL 4

__global__ void __launch_bounds__(GATHER_THREADS)
k_replay_gather_flat(const float* __restrict__ state, const float* __restrict__ action,
                     const float* __restrict__ reward, const float* __restrict__ next_state,
                     const float* __restrict__ gamma, long long cap, int S, int A,
                     const long long* __restrict__ idx, int B, int gr, unsigned magic,
                     float* __restrict__ s_out, float* __restrict__ a_out, float* __restrict__ r_out,
                     float* __restrict__ s2_out, float* __restrict__ g_out) {
  __shared__ long long sidx[GATHER_MAX_ROWS];
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * gr;
  const int nrow = (int)min((long long)gr, (long long)B - row0);
  if (tid < nrow) {
    const long long j = idx[row0 + tid];
    sidx[tid] = (j < 0 || j >= cap) ? -1ll : j;   // host wrapper validates indices; never read out of bounds
  }
  __syncthreads();
  const unsigned E = (unsigned)(2 * S + A + 2);
  const unsigned total = (unsigned)nrow * E;
  for (unsigned base = tid; base < total; base += GATHER_THREADS * GATHER_UNROLL) {
    float v[GATHER_UNROLL];
    float* dst[GATHER_UNROLL];
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const unsigned i = base + u * GATHER_THREADS;
      dst[u] = nullptr;
      if (i < total) {
        const unsigned r = __umulhi(i, magic);
        int e = (int)(i - r * E);
        const long long j = sidx[r];
        const long long row = row0 + r;
        if (j >= 0) {
          const float* src;
          if (e < S) { src = state + j * S + e; dst[u] = s_out + row * S + e; }
          else if ((e -= S) < S) { src = next_state + j * S + e; dst[u] = s2_out + row * S + e; }
          else if ((e -= S) < A) { src = action + j * A + e; dst[u] = a_out + row * A + e; }
          else if (e == A) { src = reward + j; dst[u] = r_out + row; }
          else { src = gamma + j; dst[u] = g_out + row; }
          v[u] = __ldg(src);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u)
      if (dst[u]) *dst[u] = v[u];
  }
}

__global__ void __launch_bounds__(256)
k_replay_scatter(float* __restrict__ state, float* __restrict__ action, float* __restrict__ reward,
                 float* __restrict__ next_state, float* __restrict__ gamma, long long cap, int S,
                 int A, const long long* __restrict__ slot, int n, const float* __restrict__ s_in,
                 const float* __restrict__ a_in, const float* __restrict__ r_in,
                 const float* __restrict__ s2_in, const float* __restrict__ g_in) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const long long j = slot[row];
  if (j < 0 || j >= cap) return;
  copy_row(s_in + (long long)row * S, state + j * S, S, lane);
  copy_row(s2_in + (long long)row * S, next_state + j * S, S, lane);
  copy_row(a_in + (long long)row * A, action + j * A, A, lane);
  if (lane == 0) {
    reward[j] = r_in[row];
    gamma[j] = g_in[row];
  }
}

extern "C" int rlc_replay_gather(rlc_handle* h, const float* state, const float* action,
                                 const float* reward, const float* next_state, const float* gamma,
                                 int64_t cap, int S, int A, const int64_t* idx, int B,
                                 float* s_out, float* a_out, float* r_out, float* s2_out,
                                 float* g_out, void* stream) {
  RLC_REQUIRE(h && state && action && reward && next_state && gamma && idx);
  RLC_REQUIRE(s_out && a_out && r_out && s2_out && g_out && cap >= 1 && S >= 1 && A >= 1 && B >= 0);
  if (B == 0) return RLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long E = 2LL * S + A + 2;
  if (E > 8192) {   // reciprocal checked exhaustively for E <= 8192 (tests/test_abi.py)
    k_replay_gather<<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, st>>>(
        state, action, reward, next_state, gamma, cap, S, A, (const long long*)idx, B, s_out, a_out,
        r_out, s2_out, g_out);
  } else {
    // small minibatches: few rows per CTA so the launch still spreads over the SMs
    const int gr = B >= h->num_sms * 16 * GATHER_MAX_ROWS ? GATHER_MAX_ROWS : 8;
    const unsigned magic = (unsigned)((1ULL << 32) / (unsigned long long)E) + 1u;
    k_replay_gather_flat<<<(unsigned)((B + gr - 1) / gr), GATHER_THREADS, 0, st>>>(
        state, action, reward, next_state, gamma, cap, S, A, (const long long*)idx, B, gr, magic, s_out,
        a_out, r_out, s2_out, g_out);
  }
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_replay_scatter(rlc_handle* h, float* state, float* action, float* reward,
                                  float* next_state, float* gamma, int64_t cap, int S, int A,
                                  const int64_t* slot, int n, const float* s_in, const float* a_in,
                                  const float* r_in, const float* s2_in, const float* g_in,
                                  void* stream) {
  RLC_REQUIRE(h && state && action && reward && next_state && gamma && slot);
  RLC_REQUIRE(s_in && a_in && r_in && s2_in && g_in && cap >= 1 && S >= 1 && A >= 1 && n >= 0);
  if (n == 0) return RLC_OK;
  k_replay_scatter<<<(unsigned)(((long long)n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      state, action, reward, next_state, gamma, cap, S, A, (const long long*)slot, n, s_in, a_in,
      r_in, s2_in, g_in);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Record layout (the HBM-first layout; ReplayBuffer(layout="record")): one ring of fixed-stride records
// rec[cap, stride] = [state S | action A | reward | next_state S | gamma | pad], stride a multiple of 16 floats so a
// record starts on a 64-byte boundary -- the DRAM access granularity.  A random transition then costs
// ceil(4(2S+A+2)/64) 64-byte reads of ONE DRAM page instead of five partially used reads in five arrays (ncu on the
// struct-of-arrays gather at S=17, A=6: 646 MB read per 1M transitions against 176 MB algorithmic; a 192-byte record
// reads 192 MB).  Same CTA structure as k_replay_gather_flat, but field by field: the dense side of a field is
// contiguous over the CTA's rows, so only the record address needs the (row, column) split.
// ---------------------------------------------------------------------------------------------
template <bool GATHER>
__device__ __forceinline__ void rec_field(float* __restrict__ rec, int stride, const long long* sidx,
                                          float* __restrict__ dense, int off, int W, unsigned magic, int nrow,
                                          int tid) {
  const unsigned total = (unsigned)(nrow * W);
  for (unsigned base = tid; base < total; base += GATHER_THREADS * GATHER_UNROLL) {
    float v[GATHER_UNROLL];
    long long ra[GATHER_UNROLL];
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      const unsigned i = base + u * GATHER_THREADS;
      ra[u] = -1;
      if (i < total) {
        const unsigned r = W == 1 ? i : __umulhi(i, magic);
        const long long j = sidx[r];
        if (j >= 0) {
          ra[u] = j * stride + off + (int)(i - r * (unsigned)W);
          v[u] = GATHER ? __ldg(rec + ra[u]) : __ldg(dense + i);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GATHER_UNROLL; ++u) {
      if (ra[u] >= 0) {
        if (GATHER) dense[base + u * GATHER_THREADS] = v[u];
        else rec[ra[u]] = v[u];
      }
    }
  }
}

template <bool GATHER>
__global__ void __launch_bounds__(GATHER_THREADS)
k_replay_rec(float* __restrict__ rec, long long cap, int stride, int S, int A, const long long* __restrict__ idx,
             int B, int gr, unsigned magicS, unsigned magicA, float* __restrict__ s, float* __restrict__ a,
             float* __restrict__ r, float* __restrict__ s2, float* __restrict__ g) {
  __shared__ long long sidx[GATHER_MAX_ROWS];
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * gr;
  const int nrow = (int)min((long long)gr, (long long)B - row0);
  if (tid < nrow) {
    const long long j = idx[row0 + tid];
    sidx[tid] = (j < 0 || j >= cap) ? -1ll : j;   // out-of-range slots are skipped, never dereferenced
  }
  __syncthreads();
  rec_field<GATHER>(rec, stride, sidx, s + row0 * S, 0, S, magicS, nrow, tid);
  rec_field<GATHER>(rec, stride, sidx, a + row0 * A, S, A, magicA, nrow, tid);
  rec_field<GATHER>(rec, stride, sidx, s2 + row0 * S, S + A + 1, S, magicS, nrow, tid);
  if (tid < nrow && sidx[tid] >= 0) {             // the two scalars of a row: one thread per row
    float* p = rec + sidx[tid] * stride;
    if (GATHER) { r[row0 + tid] = __ldg(p + S + A); g[row0 + tid] = __ldg(p + 2 * S + A + 1); }
    else { p[S + A] = r[row0 + tid]; p[2 * S + A + 1] = g[row0 + tid]; }
  }
}

extern "C" int rlc_replay_rec_stride(int S, int A) {
  if (S < 1 || A < 1 || S > 4000 || A > 4000) return RLC_ERR_INVALID;
  return (2 * S + A + 2 + 15) / 16 * 16;
}

template <bool GATHER>
static int replay_rec_launch(rlc_handle* h, float* rec, int64_t cap, int stride, int S, int A, const int64_t* idx,
                             int B, float* s, float* a, float* r, float* s2, float* g, void* stream) {
  RLC_REQUIRE(h && rec && idx && s && a && r && s2 && g && cap >= 1 && B >= 0);
  RLC_REQUIRE(S >= 1 && A >= 1 && S <= 4000 && A <= 4000 && stride >= 2 * S + A + 2);
  if (B == 0) return RLC_OK;
  const int gr = B >= h->num_sms * 16 * GATHER_MAX_ROWS ? GATHER_MAX_ROWS : 8;
  const unsigned magicS = S == 1 ? 0u : (unsigned)((1ULL << 32) / (unsigned)S) + 1u;
  const unsigned magicA = A == 1 ? 0u : (unsigned)((1ULL << 32) / (unsigned)A) + 1u;
  k_replay_rec<GATHER><<<(unsigned)((B + gr - 1) / gr), GATHER_THREADS, 0, (cudaStream_t)stream>>>(
      rec, cap, stride, S, A, (const long long*)idx, B, gr, magicS, magicA, s, a, r, s2, g);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}

extern "C" int rlc_replay_gather_rec(rlc_handle* h, const float* rec, int64_t cap, int stride, int S, int A,
                                     const int64_t* idx, int B, float* s_out, float* a_out, float* r_out,
                                     float* s2_out, float* g_out, void* stream) {
  return replay_rec_launch<true>(h, const_cast<float*>(rec), cap, stride, S, A, idx, B, s_out, a_out, r_out, s2_out,
                                 g_out, stream);
}

extern "C" int rlc_replay_scatter_rec(rlc_handle* h, float* rec, int64_t cap, int stride, int S, int A,
                                      const int64_t* slot, int n, const float* s_in, const float* a_in,
                                      const float* r_in, const float* s2_in, const float* g_in, void* stream) {
  return replay_rec_launch<false>(h, rec, cap, stride, S, A, slot, n, const_cast<float*>(s_in),
                                  const_cast<float*>(a_in), const_cast<float*>(r_in), const_cast<float*>(s2_in),
                                  const_cast<float*>(g_in), stream);
}

// ---------------------------------------------------------------------------------------------
// N4: device-side minibatch index sampling -- k DISTINCT uniform indices in [0, n), written as ring
// slots (head + i) % cap, with no host round trip (the host sampler RandomAccessQueue.sample_n_k,
// custom_collections.py:107-131, stays the default because it reproduces the reference's stream).
// Counter-based Philox4x32-10 keyed by (seed, call counter): draw (slot i, round r) is a pure function
// of (seed, counter, i, r).  Duplicates are resolved deterministically: every unresolved slot inserts
// (value, priority = round<<16 | slot) into an open-addressing hash set in shared memory with
// atomicMin on the packed word, the lowest priority keeps the value, the others redraw next round.
// One CTA; k <= 4096, 3k < n (the reference's own condition for its rejection scheme).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define SAMPLE_THREADS 256
#define SAMPLE_MAX_K 4096

__global__ void __launch_bounds__(SAMPLE_THREADS)
k_sample_distinct(long long n, int k, unsigned long long seed, unsigned long long counter, long long head,
                  long long cap, long long* __restrict__ idx_out, long long* __restrict__ slot_out,
                  int table_size) {
  extern __shared__ unsigned long long tab[];   // entry: (value+1) << 32 | priority ; 0 = empty
  __shared__ int unresolved;
  const int tid = threadIdx.x;
  for (int i = tid; i < table_size; i += SAMPLE_THREADS) tab[i] = 0ull;
  // per-slot state lives in registers: slots tid, tid+256, ...
  constexpr int PER = SAMPLE_MAX_K / SAMPLE_THREADS;
  unsigned int val[PER];
  bool done[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) { done[j] = (tid + j * SAMPLE_THREADS) >= k; val[j] = 0u; }
  __syncthreads();
  const unsigned mask = (unsigned)table_size - 1u;
  for (int round = 0; round < 64; ++round) {
    if (tid == 0) unresolved = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (done[j]) continue;
      const int slot = tid + j * SAMPLE_THREADS;
      uint32_t r[4];
      philox4x32_10((uint32_t)slot, (uint32_t)round, (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)seed,
                    (uint32_t)(seed >> 32), r);
      // 64 random bits -> [0, n) by multiply-high (bias < n / 2^64)
      const unsigned long long r64 = ((unsigned long long)r[0] << 32) | r[1];
      const unsigned int x = (unsigned int)__umul64hi(r64, (unsigned long long)n);
      val[j] = x;
      const unsigned long long word = ((unsigned long long)(x + 1u) << 32) | (unsigned)((round << 16) | slot);
      unsigned h = (x * 2654435761u) & mask;
      while (true) {
        const unsigned long long cur = *(volatile unsigned long long*)&tab[h];
        if (cur == 0ull) {
          if (atomicCAS(&tab[h], 0ull, word) == 0ull) break;
          continue;                                   // lost the race for the empty entry: re-read it
        }
        if ((unsigned int)(cur >> 32) == x + 1u) { atomicMin(&tab[h], word); break; }
        h = (h + 1u) & mask;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (done[j]) continue;
      const int slot = tid + j * SAMPLE_THREADS;
      const unsigned int x = val[j];
      unsigned h = (x * 2654435761u) & mask;
      while ((unsigned int)(*(volatile unsigned long long*)&tab[h] >> 32) != x + 1u) h = (h + 1u) & mask;
      if ((unsigned int)(tab[h] & 0xffffffffu) == (unsigned)((round << 16) | slot)) {
        done[j] = true;
        if (idx_out) idx_out[slot] = (long long)x;
        if (slot_out) slot_out[slot] = (head + (long long)x) % cap;
      } else {
        atomicAdd(&unresolved, 1);
      }
    }
    __syncthreads();
    if (unresolved == 0) break;
    __syncthreads();
  }
}

extern "C" int rlc_replay_sample(rlc_handle* h, int64_t n, int k, uint64_t seed, uint64_t counter, int64_t head,
                                 int64_t cap, int64_t* idx_out, int64_t* slot_out, void* stream) {
  RLC_REQUIRE(h && (idx_out || slot_out) && n >= 1 && k >= 0 && cap >= n && head >= 0 && head < cap);
  RLC_REQUIRE(n < (1LL << 31));
  if (k > SAMPLE_MAX_K || 3LL * k >= n) return RLC_ERR_UNSUPPORTED;   // the host sampler covers these
  if (k == 0) return RLC_OK;
  int table = 1;
  while (table < 4 * k) table <<= 1;                                   // load factor <= 1/4... 1/2
  const size_t smem = (size_t)table * sizeof(unsigned long long);
  if (smem > h->smem_optin) return RLC_ERR_UNSUPPORTED;
  RLC_CUDA(cudaFuncSetAttribute(k_sample_distinct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sample_distinct<<<1, SAMPLE_THREADS, smem, (cudaStream_t)stream>>>(
      (long long)n, k, (unsigned long long)seed, (unsigned long long)counter, (long long)head, (long long)cap,
      (long long*)idx_out, (long long*)slot_out, table);
  RLC_LAUNCH_CHECK(h);
  return RLC_OK;
}
