/* Host helper of the device-resident loop (rlcontrol_b200/device_loop.py): the minibatch indices of a whole chunk of
 * steps, drawn from numpy's legacy RandomState stream exactly as the reference draws them one call at a time.
 *
 * Reference: RandomAccessQueue.sample_n_k (utils/custom_collections.py:107-131) on np.random.RandomState(seed)
 * (utils/replaybuffer.py:18): when 3k < n, result = rng.choice(n, 2k); duplicates among the first k are replaced from
 * the second half, which is refilled with rng.choice(n, k) when exhausted.  RandomState.choice(int n, size) without p is
 * randint(0, n, size): 32-bit MT19937 outputs, masked to the smallest 2^b - 1 >= n - 1, rejected while > n - 1 (numpy
 * random/_bounded_integers + legacy-distributions; restated from the published algorithm, and pinned bit for bit on
 * numpy itself by tests/test_oracle_env.py).  The Python loop costs ~10 us per step per run; with 8 runs sharing a GPU
 * that made the single host thread the bottleneck.
 *
 * Plain C, no CUDA: built by the same Makefile into librlc_host.so and loaded with ctypes; device_loop.py falls back to
 * its Python implementation (same stream) when the library is absent. */
#include <stdint.h>
#include <string.h>

#define MT_N 624
#define MT_M 397

typedef struct {
  uint32_t key[MT_N];
  int pos;
} mt_state;

static void mt_gen(mt_state* s) {
  uint32_t y;
  int i;
  for (i = 0; i < MT_N - MT_M; i++) {
    y = (s->key[i] & 0x80000000u) | (s->key[i + 1] & 0x7fffffffu);
    s->key[i] = s->key[i + MT_M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  for (; i < MT_N - 1; i++) {
    y = (s->key[i] & 0x80000000u) | (s->key[i + 1] & 0x7fffffffu);
    s->key[i] = s->key[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  }
  y = (s->key[MT_N - 1] & 0x80000000u) | (s->key[0] & 0x7fffffffu);
  s->key[MT_N - 1] = s->key[MT_M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
  s->pos = 0;
}

static inline uint32_t mt_next(mt_state* s) {
  uint32_t y;
  if (s->pos == MT_N) mt_gen(s);
  y = s->key[s->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* randint(0, n, size=cnt) of the legacy stream, n <= 2^32 */
static void bounded_fill(mt_state* s, uint32_t n, int cnt, int64_t* out) {
  const uint32_t rng = n - 1;
  uint32_t mask = rng, val;
  int i;
  if (rng == 0) {
    for (i = 0; i < cnt; i++) out[i] = 0;
    return;
  }
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  for (i = 0; i < cnt; i++) {
    if (rng == 0xffffffffu) val = mt_next(s);
    else while ((val = (mt_next(s) & mask)) > rng) {}
    out[i] = (int64_t)val;
  }
}

/* sample_n_k for the rejection branch (3k < n): out[k]. scratch: 2k int64. */
static void sample_n_k(mt_state* s, int64_t n, int k, int64_t* result, int32_t* out) {
  int i, j = k, t;
  bounded_fill(s, (uint32_t)n, 2 * k, result);
  for (i = 0; i < k; i++) {
    int64_t x = result[i];
    for (;;) {                                   /* while x in selected (= result[0..i)) */
      int dup = 0;
      for (t = 0; t < i; t++)
        if (result[t] == x) { dup = 1; break; }
      if (!dup) break;
      x = result[i] = result[j];
      j += 1;
      if (j == 2 * k) {
        bounded_fill(s, (uint32_t)n, k, result + k);
        j = k;
      }
    }
  }
  for (i = 0; i < k; i++) out[i] = (int32_t)result[i];
}

/* For each of `steps` steps: n[i] = population (replay size) or 0 = no minibatch that step; rows with n > 0 must satisfy
 * 3k < n <= 2^32 (the caller handles the permutation branch in Python).  out[steps][k] int32.  key/pos: the MT19937
 * state of RandomState.get_state(), updated in place.  Returns 0, or -1 on invalid arguments. */
int rlc_host_sample_chunk(uint32_t* key, int* pos, const int64_t* n, int steps, int k, int32_t* out) {
  mt_state s;
  int64_t scratch[2 * 4096];
  int i;
  if (!key || !pos || !n || !out || k < 1 || k > 4096 || *pos < 0 || *pos > MT_N) return -1;
  memcpy(s.key, key, sizeof(s.key));
  s.pos = *pos;
  for (i = 0; i < steps; i++) {
    if (n[i] == 0) continue;
    if (3 * (int64_t)k >= n[i] || n[i] > 4294967296LL) return -1;
    sample_n_k(&s, n[i], k, scratch, out + (size_t)i * k);
  }
  memcpy(key, s.key, sizeof(s.key));
  *pos = s.pos;
  return 0;
}
