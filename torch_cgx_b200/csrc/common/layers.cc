#include "layers.h"

#include <numeric>
#include <stdexcept>

namespace cgx {

LayerRegistry& LayerRegistry::instance() {
  static LayerRegistry inst;
  return inst;
}

void LayerRegistry::register_layer(unsigned bucket_idx, unsigned layer_idx, int64_t numel, int bits,
                                   int bucket_size) {
  if (numel < 0) throw std::invalid_argument("cgx: register_layer with negative numel");
  std::lock_guard<std::mutex> g(mu_);
  if (sizes_.size() <= bucket_idx) sizes_.resize(bucket_idx + 1);
  auto& b = sizes_[bucket_idx];
  if (b.size() <= layer_idx) b.resize(layer_idx + 1, 0);
  b[layer_idx] = numel;
  LayerConfig c;
  c.bits = (bits >= 1 && bits <= 8) ? bits : kRawBits;
  c.bucket_size = bucket_size > 0 ? bucket_size : kDefaultBucketSize;
  configs_[{bucket_idx, layer_idx}] = c;
  ++version_;
}

void LayerRegistry::set_bits(unsigned bucket_idx, unsigned layer_idx, int bits) {
  std::lock_guard<std::mutex> g(mu_);
  auto& c = configs_[{bucket_idx, layer_idx}];
  c.bits = (bits >= 1 && bits <= 8) ? bits : kRawBits;
  ++version_;
}

void LayerRegistry::set_bucket_size(unsigned bucket_idx, unsigned layer_idx, int bucket_size) {
  if (bucket_size <= 0) throw std::invalid_argument("cgx: bucket size must be positive");
  std::lock_guard<std::mutex> g(mu_);
  configs_[{bucket_idx, layer_idx}].bucket_size = bucket_size;
  ++version_;
}

void LayerRegistry::reset() {
  std::lock_guard<std::mutex> g(mu_);
  sizes_.clear();
  configs_.clear();
  cursor_ = 0;
  ++version_;
}

size_t LayerRegistry::num_buckets() const {
  std::lock_guard<std::mutex> g(mu_);
  return sizes_.size();
}

std::vector<int64_t> LayerRegistry::bucket_layer_sizes(unsigned bucket_idx) const {
  std::lock_guard<std::mutex> g(mu_);
  if (bucket_idx >= sizes_.size()) return {};
  return sizes_[bucket_idx];
}

LayerConfig LayerRegistry::layer_config(unsigned bucket_idx, unsigned layer_idx) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = configs_.find({bucket_idx, layer_idx});
  return it == configs_.end() ? LayerConfig{} : it->second;
}

uint64_t LayerRegistry::version() const {
  std::lock_guard<std::mutex> g(mu_);
  return version_;
}

std::vector<LayerSpec> LayerRegistry::extract(int64_t numel, const CompressionEnv& env, int min_compress_elems,
                                              int explicit_bucket, int* resolved_bucket) {
  std::lock_guard<std::mutex> g(mu_);
  if (resolved_bucket) *resolved_bucket = -1;
  std::vector<LayerSpec> out;
  auto single = [&](int bits, int bucket) {
    LayerSpec l;
    l.elem_off = 0;
    l.numel = (uint64_t)numel;
    l.bits = compression_enabled(numel, bits, min_compress_elems) ? bits : kRawBits;
    l.bucket = (uint32_t)bucket;
    out.push_back(l);
  };
  // tiny tensors are never compressed (reference: numel < 16 -> uncompressed)
  if (numel < kMinCompressElems) {
    single(kRawBits, env.bucket_size);
    return out;
  }
  auto total = [&](size_t b) { return std::accumulate(sizes_[b].begin(), sizes_[b].end(), (int64_t)0); };
  int chosen = -1;
  if (!sizes_.empty()) {
    if (explicit_bucket >= 0) {
      if ((size_t)explicit_bucket < sizes_.size() && total((size_t)explicit_bucket) == numel) chosen = explicit_bucket;
    } else {
      if (cursor_ < sizes_.size() && total(cursor_) == numel) {
        chosen = (int)cursor_;
      } else {
        int found = -1, count = 0;
        for (size_t b = 0; b < sizes_.size(); ++b)
          if (total(b) == numel) {
            found = (int)b;
            ++count;
          }
        if (count == 1) chosen = found;
      }
      if (chosen >= 0) cursor_ = ((size_t)chosen + 1) % sizes_.size();
    }
  }
  if (chosen < 0) {
    // no registered layout: the whole buffer is one layer with the env config
    single(env.bits, env.bucket_size);
    return out;
  }
  if (resolved_bucket) *resolved_bucket = chosen;
  uint64_t off = 0;
  const auto& ls = sizes_[(size_t)chosen];
  out.reserve(ls.size());
  for (size_t i = 0; i < ls.size(); ++i) {
    if (ls[i] == 0) continue;
    LayerSpec l;
    l.elem_off = off;
    l.numel = (uint64_t)ls[i];
    auto it = configs_.find({(unsigned)chosen, (unsigned)i});
    LayerConfig c = it == configs_.end() ? LayerConfig{env.bits, env.bucket_size} : it->second;
    l.bits = compression_enabled(ls[i], c.bits, min_compress_elems) ? c.bits : kRawBits;
    l.bucket = (uint32_t)c.bucket_size;
    out.push_back(l);
    off += (uint64_t)ls[i];
  }
  return out;
}

}  // namespace cgx
