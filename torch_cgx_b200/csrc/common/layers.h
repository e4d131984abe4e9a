// Process-global registry of the layers inside each DDP gradient bucket and
// their per-layer compression config. Filled by the Python comm hook through
// `register_layer` (third backward pass, once DDP's bucket layout is final).
//
// Reference: MPIAllReduce_Operation::RegisterLayer / SetQBits / SetQBucketSize /
// extractLayers (/root/reference/src/mpi_allreduce_operations.h:37-59,
// .cc:257-285), Compressor::layers_configs (/root/reference/src/common/
// compressor.h:93-125), Layer (/root/reference/src/common/layer.h:26-45).
// Fixed vs. the reference: set_quantization_bucket_size really sets the bucket
// size; a bucket whose size does not match the cursor is
// looked up by size instead of throwing.
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include "config.h"
#include "plan.h"

namespace cgx {

struct LayerConfig {
  int bits = kDefaultBits;
  int bucket_size = kDefaultBucketSize;
};

class LayerRegistry {
 public:
  static LayerRegistry& instance();

  void register_layer(unsigned bucket_idx, unsigned layer_idx, int64_t numel, int bits, int bucket_size);
  void set_bits(unsigned bucket_idx, unsigned layer_idx, int bits);
  void set_bucket_size(unsigned bucket_idx, unsigned layer_idx, int bucket_size);
  void reset();
  size_t num_buckets() const;
  std::vector<int64_t> bucket_layer_sizes(unsigned bucket_idx) const;
  LayerConfig layer_config(unsigned bucket_idx, unsigned layer_idx) const;

  // Resolve the layer list of a flat buffer with `numel` elements.
  //  explicit_bucket >= 0: the caller knows which DDP bucket this is.
  //  explicit_bucket <  0: reference behaviour -- a cursor that cycles through
  //                        the registered buckets in call order.
  // Layers that must not be compressed come back with bits == 32.
  // `resolved_bucket` receives the bucket index used, or -1 (unregistered).
  std::vector<LayerSpec> extract(int64_t numel, const CompressionEnv& env, int min_compress_elems,
                                 int explicit_bucket, int* resolved_bucket);

  // bumps whenever anything changes (plan cache invalidation)
  uint64_t version() const;

 private:
  mutable std::mutex mu_;
  std::vector<std::vector<int64_t>> sizes_;               // [bucket][layer] -> numel
  std::map<std::pair<unsigned, unsigned>, LayerConfig> configs_;
  size_t cursor_ = 0;
  uint64_t version_ = 1;
};

inline bool compression_enabled(int64_t numel, int bits, int min_compress_elems) {
  return numel > (int64_t)min_compress_elems && bits >= 1 && bits <= 8;
}

}  // namespace cgx
