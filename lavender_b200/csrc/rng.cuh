// Counter-based dropout masks (HF BERT hidden / attention-probability dropout, model.py:242 in train() mode).
// A mask bit is a pure function of (seed, step, site, row, column[, head]): the forward kernel and the backward
// kernel regenerate the same bits, nothing mask-shaped is stored, and a CUDA-graph replay sees a new step counter
// through device memory.  Generator: Philox-4x32 with 7 rounds (Salmon et al., SC'11: passes BigCrush), one call
// per 8 elements, 16 random bits per element (drop iff bits < round(p * 65536)).
#pragma once
#include <stdint.h>

#include "../../include/lavender_b200.h"

namespace lav {

struct DropParams {  // resolved on the host from LavDropout; `on == 0` disables
  const uint64_t* rng;  // device: [0] seed, [1] step
  uint32_t site;
  uint32_t thresh;      // 16-bit threshold: element dropped iff r16 < thresh
  float inv_keep;       // 1 / (1 - thresh / 65536)
  int on;
};

static inline DropParams make_drop(const LavDropout* d) {
  DropParams p{};
  if (d && d->rng && d->p > 0.f) {
    p.rng = d->rng, p.site = d->site, p.on = 1;
    double t = (double)d->p * 65536.0 + 0.5;
    if (t > 65535.0) t = 65535.0;
    p.thresh = (uint32_t)t;
    p.inv_keep = (float)(1.0 / (1.0 - (double)p.thresh / 65536.0));
  }
  return p;
}

struct DropKey {
  uint32_t k0, k1, site;
};
__device__ __forceinline__ DropKey drop_key(const DropParams& p) {
  const uint64_t seed = p.rng[0], step = p.rng[1];
  DropKey k;
  k.k0 = (uint32_t)seed ^ (uint32_t)step;
  k.k1 = (uint32_t)(seed >> 32) + (uint32_t)(step >> 32) * 0x85EBCA6Bu;
  k.site = p.site;
  return k;
}

// keep bits of the 8 elements (a, 8*b .. 8*b+7, c): bit j set = element j is KEPT
__device__ __forceinline__ uint32_t drop_keep8(const DropKey& key, uint32_t thresh, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t c0 = a, c1 = b, c2 = c, c3 = key.site;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0, c1 = lo1, c2 = hi0 ^ c3 ^ k1, c3 = lo0;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  uint32_t m = 0;
  m |= ((c0 & 0xFFFFu) >= thresh) ? 1u : 0u;
  m |= ((c0 >> 16) >= thresh) ? 2u : 0u;
  m |= ((c1 & 0xFFFFu) >= thresh) ? 4u : 0u;
  m |= ((c1 >> 16) >= thresh) ? 8u : 0u;
  m |= ((c2 & 0xFFFFu) >= thresh) ? 16u : 0u;
  m |= ((c2 >> 16) >= thresh) ? 32u : 0u;
  m |= ((c3 & 0xFFFFu) >= thresh) ? 64u : 0u;
  m |= ((c3 >> 16) >= thresh) ? 128u : 0u;
  return m;
}

}  // namespace lav
