// Single-process, W-rank simulation of the compressed Scatter-Reduce-AllGather
// on CPU memory. Performs exactly the arithmetic of the fused CUDA kernel
// (same plan, same block primitives, same summation order, same RNG streams),
// so tests can demand bit-equality between the GPU result and this oracle.
// Numerics spec: SURVEY.md §2.7 ("SRA numerics"),
// /root/reference/src/common/scatter_reduce_allgather.cc:143-160.
#pragma once
#include <cstdint>
#include <vector>
#include "philox.h"
#include "plan.h"

namespace cgx {

struct RngParams {
  uint64_t seed = 0;
  uint32_t seq = 0;
  bool stochastic = false;
};

inline RngKey make_rng_key(const RngParams& r, int rank, int phase) {
  RngKey k;
  k.seed_lo = (uint32_t)r.seed;
  k.seed_hi = (uint32_t)(r.seed >> 32);
  k.seq = r.seq;
  k.stream = (uint32_t)rank * 2u + (uint32_t)phase;
  k.enabled = r.stochastic ? 1u : 0u;
  return k;
}

// bufs[r] = base pointer of rank r's tensor (dtype plan.dtype), reduced in place.
void sra_simulate(const Plan& plan, const std::vector<void*>& bufs, float prescale, const RngParams& rng);

// One-shot allreduce oracle: result = T( sum over ranks r (in order) of decode(Q_r(x_r * prescale)) ),
// every contribution quantized exactly once; `plan` must be a single-chunk plan (world == 1).
void oneshot_simulate(const Plan& plan, const std::vector<void*>& bufs, float prescale, const RngParams& rng);

// Quantize->dequantize round trip of one rank's buffer through the plan's
// blocks (what a single compression step does to the data); used by the Python
// ops and the tests.
void roundtrip_simulate(const Plan& plan, void* buf, float prescale, const RngParams& rng, int rank, int phase);

}  // namespace cgx
