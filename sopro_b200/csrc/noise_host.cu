// Host-side noise tapes: the Exp(1) draws torch.multinomial consumes on the CPU, reproduced bit for bit WITHOUT computing the
// draws nobody reads.  torch.multinomial(p, 1) on CPU is argmax(p / q), q = torch.empty_like(p).exponential_(1) drawn from
// the CPU generator (reference sampling.py:83-93 via ATen): per element one 64-bit random = two mt19937 outputs
// (CPUGeneratorImpl::random64: hi word first), u = (r & (2^53 - 1)) * 2^-53, q = float(-log1p(-u))
// (ATen/core/DistributionsHelper.h: exponential_distribution<double>, TransformationHelper.h: uniform_real, exponential;
// ATen/native/cpu/DistributionTemplates.h: serial kernel, row-major element order).  The sampler reads only the first
// top_k columns of each [vocab] row (rank-aligned draw, sopro_b200/sampling.py), so the other 2 x (vocab - top_k)
// generator outputs of a row are skipped: the state advances (one twist per output) but no tempering, no logarithm.
// 64 utterances x 401 steps x 2049 draws: ~6 ms on 16 threads instead of the 40-70 ms torch takes to materialise them.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>

#include "../../include/sopro_b200.h"

struct sopro_noise {
  uint32_t st[624];
  int next;  // index of the next output in st (624 = refill first)
};

namespace {
inline void mt_seed(sopro_noise* g, uint64_t seed) {
  g->st[0] = (uint32_t)(seed & 0xffffffffu);
  for (int j = 1; j < 624; ++j) g->st[j] = 1812433253u * (g->st[j - 1] ^ (g->st[j - 1] >> 30)) + (uint32_t)j;
  g->next = 624;
}
inline uint32_t twist(uint32_t u, uint32_t v) { return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u); }
inline void mt_refill(sopro_noise* g) {
  uint32_t* p = g->st;
  int j = 0;
  for (; j < 624 - 397; ++j) p[j] = p[j + 397] ^ twist(p[j], p[j + 1]);
  for (; j < 623; ++j) p[j] = p[j + 397 - 624] ^ twist(p[j], p[j + 1]);
  p[623] = p[396] ^ twist(p[623], p[0]);
  g->next = 0;
}
inline uint32_t mt_next(sopro_noise* g) {
  if (g->next >= 624) mt_refill(g);
  uint32_t y = g->st[g->next++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
inline void mt_discard(sopro_noise* g, long long n) {
  while (n > 0) {
    if (g->next >= 624) mt_refill(g);
    const long long avail = 624 - g->next;
    const long long take = n < avail ? n : avail;
    g->next += (int)take;
    n -= take;
  }
}
}  // namespace

extern "C" {

int sopro_noise_create(uint64_t seed, sopro_noise_t** out) {
  if (!out) return SOPRO_ERR_INVALID;
  sopro_noise* g = new (std::nothrow) sopro_noise();
  if (!g) return SOPRO_ERR_CUDA;
  mt_seed(g, seed);
  *out = g;
  return SOPRO_OK;
}

int sopro_noise_rows(sopro_noise_t* g, int n_rows, int vocab, int keep, float* out) {
  if (!g || !out || n_rows < 0 || vocab < 1 || keep < 1 || keep > vocab) return SOPRO_ERR_INVALID;
  const double scale = 1.0 / 9007199254740992.0;  // 2^-53
  for (int r = 0; r < n_rows; ++r) {
    float* o = out + (size_t)r * keep;
    for (int c = 0; c < keep; ++c) {
      const uint64_t hi = mt_next(g), lo = mt_next(g);
      const uint64_t v = (hi << 32) | lo;
      const double u = (double)(v & ((1ull << 53) - 1)) * scale;
      o[c] = (float)(-1.0 * std::log1p(-u));
    }
    mt_discard(g, 2ll * (vocab - keep));
  }
  return SOPRO_OK;
}

int sopro_noise_destroy(sopro_noise_t* g) {
  delete g;
  return SOPRO_OK;
}

}  // extern "C"
