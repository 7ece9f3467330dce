/* Plain-C restatement of the reference critic forward passes (TEST INFRASTRUCTURE ONLY -- see
 * oracle/oracle_np.py; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this).  Scalar loops in the order the reference's ops define:
 *   T-in : SoftQNetwork.forward, forwardkl_network.py:263-268 (torch [out,in] weights)
 *   T-mid: critic network(), critic_network.py:77-99 / qt_opt_network.py:83-105 (TF [in,out])
 * The stacked [B*N, .] inputs of forwardkl_network.py:160-164 are materialised row by row, as
 * the reference does, rather than hoisted. */
#include <math.h>
#include <stdlib.h>

static inline float relu(float x) { return x > 0.f ? x : 0.f; }

/* q[B*N]; s[B,S]; a[N,A] (per_state=0) or [B,N,A] (per_state=1); rows [r_begin, r_end) only, so the
 * caller can spread row ranges over host threads (ctypes releases the GIL; no OpenMP in this image). */
int oracle_tin_eval(const float* s, const float* a, int per_state, int B, int N, int S, int A, int H1,
                    int H2, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* W3, const float* b3, float* q, long r_begin, long r_end) {
  const int K1 = S + A;
  {
    float* x = (float*)malloc(sizeof(float) * (K1 + H1 + H2));
    float* h1 = x + K1;
    float* h2 = h1 + H1;
    for (long r = r_begin; r < r_end && r < (long)B * N; ++r) {
      const long b = r / N, n = r % N;
      for (int k = 0; k < S; ++k) x[k] = s[b * S + k];
      const float* ar = per_state ? a + r * A : a + n * A;
      for (int k = 0; k < A; ++k) x[S + k] = ar[k];
      for (int j = 0; j < H1; ++j) {
        float acc = b1[j];
        for (int k = 0; k < K1; ++k) acc += x[k] * W1[(long)j * K1 + k];
        h1[j] = relu(acc);
      }
      for (int j = 0; j < H2; ++j) {
        float acc = b2[j];
        for (int k = 0; k < H1; ++k) acc += h1[k] * W2[(long)j * H1 + k];
        h2[j] = relu(acc);
      }
      float acc = b3[0];
      for (int j = 0; j < H2; ++j) acc += h2[j] * W3[j];
      q[r] = acc;
    }
    free(x);
  }
  return 0;
}

int oracle_tmid_eval(const float* s, const float* a, int per_state, int B, int N, int S, int A, int H1,
                     int H2, const float* W1, const float* b1, const float* W2, const float* b2,
                     const float* W3, const float* b3, const float* smin, const float* smax, float* q,
                     long r_begin, long r_end) {
  {
    float* x = (float*)malloc(sizeof(float) * (S + H1 + A + H2));
    float* z = x + S; /* [h1 ; a] */
    float* h2 = z + H1 + A;
    for (long r = r_begin; r < r_end && r < (long)B * N; ++r) {
      const long b = r / N, n = r % N;
      for (int k = 0; k < S; ++k) {
        float v = s[b * S + k];
        if (smin) v = fminf(fmaxf(v, smin[k]), smax[k]);
        x[k] = v;
      }
      for (int j = 0; j < H1; ++j) {
        float acc = b1[j];
        for (int k = 0; k < S; ++k) acc += x[k] * W1[(long)k * H1 + j];
        z[j] = relu(acc);
      }
      const float* ar = per_state ? a + r * A : a + n * A;
      for (int k = 0; k < A; ++k) z[H1 + k] = ar[k];
      for (int j = 0; j < H2; ++j) h2[j] = b2[j];
      for (int k = 0; k < H1 + A; ++k) {
        const float zk = z[k];
        const float* w = W2 + (long)k * H2;
        for (int j = 0; j < H2; ++j) h2[j] += zk * w[j];
      }
      float acc = b3[0];
      for (int j = 0; j < H2; ++j) acc += relu(h2[j]) * W3[j];
      q[r] = acc;
    }
    free(x);
  }
  return 0;
}
