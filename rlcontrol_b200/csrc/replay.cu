// K6: replay minibatch gather / scatter over a device-resident struct-of-arrays ring
// (utils/replaybuffer.py:25-37).  HBM-bound: algorithmic bytes = B*(2S+A+2)*4 read + the same
// written (+8 B per index).  One warp per sampled transition; every field row is copied with
// coalesced (16-byte when aligned) loads.
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
  k_replay_gather<<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, st>>>(
      state, action, reward, next_state, gamma, cap, S, A, (const long long*)idx, B, s_out, a_out,
      r_out, s2_out, g_out);
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
