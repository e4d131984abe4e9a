// Counter-based Philox4x32-10 RNG (host+device), used for QSGD-style stochastic
// rounding. Replaces the reference's per-thread xorshift128+ state arrays
// (/root/reference/src/common/compression/gpu_rand.h:22-58), which are seeded
// from wall-clock time and cannot be reproduced or checkpointed. A counter RNG
// needs no state buffer, is independent of the thread mapping, and makes the
// CPU path an exact oracle for the GPU path.
#pragma once
#include <cstdint>
#include "quant_math.h"

namespace cgx {

struct Philox4 {
  uint32_t v[4];
};

CGX_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

CGX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c0, hi0, lo0);
    philox_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

// uniform in [0,1) with 16 bits of resolution (exactly representable in fp32)
CGX_HD float u16_to_unit_float(uint32_t x) { return (float)(x & 0xFFFFu) * (1.0f / 65536.0f); }

// Parameters that select a random stream: (seed, call sequence number, a
// stream id mixing rank and SRA phase). The per-group counter is the index of
// the group's first element.
struct RngKey {
  uint32_t seed_lo;
  uint32_t seed_hi;
  uint32_t seq;     // allreduce call counter (same on all ranks)
  uint32_t stream;  // rank * 2 + phase  (decorrelates ranks and requantization)
  uint32_t enabled; // 0 => deterministic r = 0.5
};

// Rounding offsets of one pack group (8 values): ONE Philox call keyed by the index (from the
// tensor base) of the group's first element, 16 random bits per value: value j uses the low
// (j even) or high (j odd) half of v[j/2]. Independent of how the buffer is cut into blocks,
// chunks and lanes, so every code path (CPU, fused kernel, standalone kernels) agrees.
CGX_HD Philox4 rounding_bits(const RngKey& k, uint32_t first_elem) {
  return philox4x32_10(first_elem, 0u, 0u, k.seq, k.seed_lo ^ (k.stream * 0x9E3779B9u), k.seed_hi);
}
CGX_HD float rounding_from_bits(const Philox4& a, int j) {
  return u16_to_unit_float(a.v[j >> 1] >> ((j & 1) * 16));
}
CGX_HD void rounding_offsets8(const RngKey& k, uint32_t first_elem, float r[8]) {
  if (!k.enabled) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = 0.5f;
    return;
  }
  const Philox4 a = rounding_bits(k, first_elem);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = rounding_from_bits(a, j);
}

}  // namespace cgx
