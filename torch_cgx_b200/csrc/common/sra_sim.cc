#include "sra_sim.h"

#include <stdexcept>

#include "block_ops.h"

namespace cgx {

void sra_simulate(const Plan& plan, const std::vector<void*>& bufs, float prescale, const RngParams& rng) {
  const int W = plan.world;
  if ((int)bufs.size() != W) throw std::invalid_argument("sra_simulate: need one buffer per rank");
  std::vector<float> acc(kMaxBlockElems), tmp(kMaxBlockElems);
  std::vector<uint8_t> rec(kMaxBlockElems * 4 + 64 + kMaxBlockBuckets * 8 + 64);
  for (int c = 0; c < W; ++c) {
    for (uint32_t b = plan.chunk_begin(c); b < plan.chunk_end(c); ++b) {
      const BlockDesc& d = plan.blocks[b];
      cpu::load_block(bufs[c], plan.dtype, d, prescale, acc.data());
      for (int q = 0; q < W; ++q) {
        if (q == c) continue;
        cpu::load_block(bufs[q], plan.dtype, d, prescale, tmp.data());
        cpu::quantize_block(tmp.data(), plan.dtype, d, rec.data(), make_rng_key(rng, q, 0), b);
        cpu::decode_block_add(rec.data(), plan.dtype, d, acc.data());
      }
      cpu::quantize_block(acc.data(), plan.dtype, d, rec.data(), make_rng_key(rng, c, 1), b);
      for (int p = 0; p < W; ++p) cpu::decode_block_store(rec.data(), plan.dtype, d, bufs[p]);
    }
  }
}

void oneshot_simulate(const Plan& plan, const std::vector<void*>& bufs, float prescale, const RngParams& rng) {
  if (plan.world != 1) throw std::invalid_argument("oneshot_simulate: plan must have a single chunk");
  const int W = (int)bufs.size();
  std::vector<float> acc(kMaxBlockElems), tmp(kMaxBlockElems);
  std::vector<uint8_t> rec(kMaxBlockElems * 4 + 64 + kMaxBlockBuckets * 8 + 64);
  for (uint32_t b = 0; b < plan.blocks.size(); ++b) {
    const BlockDesc& d = plan.blocks[b];
    const uint32_t n = block_n(d);
    for (uint32_t i = 0; i < n; ++i) acc[i] = 0.f;
    for (int q = 0; q < W; ++q) {
      cpu::load_block(bufs[q], plan.dtype, d, prescale, tmp.data());
      cpu::quantize_block(tmp.data(), plan.dtype, d, rec.data(), make_rng_key(rng, q, 0), b);
      cpu::decode_block_add(rec.data(), plan.dtype, d, acc.data());
    }
    for (int p = 0; p < W; ++p)
      for (uint32_t i = 0; i < n; ++i) cpu::store_elem(bufs[p], plan.dtype, (uint64_t)d.elem_off + i, acc[i]);
  }
}

void roundtrip_simulate(const Plan& plan, void* buf, float prescale, const RngParams& rng, int rank, int phase) {
  std::vector<float> acc(kMaxBlockElems);
  std::vector<uint8_t> rec(kMaxBlockElems * 4 + 64 + kMaxBlockBuckets * 8 + 64);
  for (uint32_t b = 0; b < plan.blocks.size(); ++b) {
    const BlockDesc& d = plan.blocks[b];
    cpu::load_block(buf, plan.dtype, d, prescale, acc.data());
    cpu::quantize_block(acc.data(), plan.dtype, d, rec.data(), make_rng_key(rng, rank, phase), b);
    cpu::decode_block_store(rec.data(), plan.dtype, d, buf);
  }
}

}  // namespace cgx
