#include "engine.h"

#include <algorithm>
#include <stdexcept>

namespace cgx {

std::vector<std::vector<LayerSpec>> split_for_fusion(const std::vector<LayerSpec>& layers, int elsize,
                                                     int64_t fusion_bytes) {
  std::vector<std::vector<LayerSpec>> groups;
  const uint64_t cap = std::max<uint64_t>(1, (uint64_t)fusion_bytes / (uint64_t)elsize);
  std::vector<LayerSpec> cur;
  uint64_t cur_elems = 0;
  auto flush = [&]() {
    if (!cur.empty()) groups.push_back(std::move(cur));
    cur.clear();
    cur_elems = 0;
  };
  for (const LayerSpec& l : layers) {
    if (l.numel == 0) continue;
    if (l.numel > cap) {
      // slice a huge layer at bucket-aligned boundaries, each slice its own call
      flush();
      const uint64_t gran = (l.bits >= kRawBits) ? 512 : std::max<uint32_t>(1u, l.bucket);
      uint64_t step = std::max<uint64_t>(gran, cap / gran * gran);
      for (uint64_t off = 0; off < l.numel; off += step) {
        LayerSpec s = l;
        s.elem_off = l.elem_off + off;
        s.numel = std::min<uint64_t>(step, l.numel - off);
        groups.push_back({s});
      }
      continue;
    }
    if (cur_elems + l.numel > cap) flush();
    cur.push_back(l);
    cur_elems += l.numel;
  }
  flush();
  return groups;
}

AllreduceEngine::AllreduceEngine(int rank, int world, const EngineConfig& cfg)
    : rank_(rank), world_(world), cfg_(cfg) {}

AllreduceEngine::~AllreduceEngine() {
  fused_.reset();
  heap_.reset();
}

size_t AllreduceEngine::required_slot_bytes(const EngineConfig& cfg, int world) {
  // a raw chunk of a full fusion buffer + slack for chunk imbalance (one block)
  // and per-block 16 B padding
  size_t per = (size_t)cfg.fusion_bytes / (size_t)world;
  return per + per / 8 + (size_t)kMaxBlockElems * 4 + 4096;
}

void AllreduceEngine::attach_heap(std::unique_ptr<SymmetricHeap> heap, int lanes) {
  heap_ = std::move(heap);
  fused_ = std::make_unique<FusedSra>(heap_.get(), lanes, cfg_.timeout_ms, cfg_.min_lane_elems);
}

void AllreduceEngine::check_health() {
  if (fused_) fused_->check_status();
}

void AllreduceEngine::allreduce_cuda(void* data, int dtype, int64_t numel, bool average, int explicit_bucket,
                                     cudaStream_t stream) {
  if (numel <= 0) return;
  CompressionEnv env = CompressionEnv::read();
  int64_t n_eff = numel;
  if (cfg_.fake_ratio < 1.0) n_eff = std::max<int64_t>(1, (int64_t)((double)numel * cfg_.fake_ratio));
  int resolved = -1;
  std::vector<LayerSpec> layers =
      LayerRegistry::instance().extract(n_eff, env, cfg_.min_compress_elems, explicit_bucket, &resolved);
  allreduce_cuda_layers(data, dtype, layers, average, env, stream);
}

void AllreduceEngine::allreduce_cuda_layers(void* data, int dtype, const std::vector<LayerSpec>& layers_in,
                                            bool average, const CompressionEnv& env, cudaStream_t stream) {
  if (!fused_) throw std::runtime_error("cgx: P2P path is not initialised");
  fused_->check_status();
  std::vector<LayerSpec> layers = layers_in;
  if (cfg_.dummy_compression || !cfg_.intra_compress)
    for (LayerSpec& l : layers) l.bits = kRawBits;
  const int elsize = dtype_size(dtype);
  const float prescale = average ? 1.0f / (float)world_ : 1.0f;
  ++call_seq_;
  ++stats_.calls;
  RngParams rng;
  rng.seed = env.seed;
  rng.seq = call_seq_;
  rng.stochastic = env.stochastic;

  // work list of layer groups; a group whose plan does not fit the heap slots is halved
  std::vector<std::vector<LayerSpec>> todo = split_for_fusion(layers, elsize, cfg_.fusion_bytes);
  std::reverse(todo.begin(), todo.end());
  uint32_t sub = 0;
  while (!todo.empty()) {
    std::vector<LayerSpec> g = std::move(todo.back());
    todo.pop_back();
    const DevicePlan* dp = fused_->prepare(g, dtype, env.skip_incomplete, stream);
    if (dp == nullptr) {
      if (g.size() > 1) {
        size_t half = g.size() / 2;
        std::vector<LayerSpec> a(g.begin(), g.begin() + half), b(g.begin() + half, g.end());
        todo.push_back(std::move(b));
        todo.push_back(std::move(a));
      } else {
        const LayerSpec& l = g[0];
        const uint64_t gran = (l.bits >= kRawBits) ? 512 : std::max<uint32_t>(1u, l.bucket);
        if (l.numel <= gran) throw std::runtime_error("cgx: fusion buffer too small for a single quantization bucket");
        uint64_t h = std::max<uint64_t>(gran, (l.numel / 2) / gran * gran);
        LayerSpec a = l, b = l;
        a.numel = h;
        b.elem_off = l.elem_off + h;
        b.numel = l.numel - h;
        todo.push_back({b});
        todo.push_back({a});
      }
      continue;
    }
    rng.seq = (call_seq_ << 6) | (sub++ & 63u);  // distinct random stream per sub-call
    fused_->run(*dp, data, prescale, rng, stream);
    ++stats_.kernel_launches;
    stats_.elements += dp->plan.numel;
    // bytes pushed by this rank: phase A = every chunk but mine, phase B = mine to W-1 peers
    const uint64_t mine = dp->plan.chunk_wire_bytes[rank_];
    stats_.wire_bytes += (dp->plan.total_wire - mine) + mine * (uint64_t)(world_ - 1);
    const uint64_t my_elems = dp->plan.chunk_elems[rank_];
    stats_.raw_bytes += ((dp->plan.numel - my_elems) + my_elems * (uint64_t)(world_ - 1)) * (uint64_t)elsize;
  }
}

}  // namespace cgx
