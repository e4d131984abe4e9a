// Allreduce plan: how a flat gradient buffer made of registered layers is cut
// into blocks, W rank-chunks and G lanes.
//
// Covers the roles of Quantizer::GetSizesAndOffsets (layer-aware, aligned
// chunk split, /root/reference/src/common/compressor.cc:265-299), of the
// per-slice walk in Compressor::Compress/Decompress (:62-179) and of
// MPIAllReduce_Operation::extractLayers
// (/root/reference/src/mpi_allreduce_operations.cc:257-285) -- but computed
// ONCE per bucket layout on the host and cached as a device-resident table, so
// the hot path is a single kernel launch instead of O(layers) launches.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "wire.h"

namespace cgx {

// One registered layer inside the flat buffer.
struct LayerSpec {
  uint64_t elem_off;  // offset in elements from the tensor base
  uint64_t numel;
  int bits;           // 1..8 compress, >= 32 raw
  uint32_t bucket;    // quantization bucket size
};

struct PlanOptions {
  int world = 1;
  int lanes = 1;                 // max lanes (CTAs) per rank
  int dtype = kF32;
  bool skip_incomplete = false;  // send the < bucket tail of a layer raw
  uint32_t min_lane_elems = 2048;  // do not spread fewer elements than this per lane
  uint32_t max_block_elems = kMaxBlockElems;
};

struct Plan {
  int world = 1;
  int lanes = 1;  // lanes actually used (<= options.lanes)
  int dtype = kF32;
  uint64_t numel = 0;                     // elements covered by the plan
  std::vector<BlockDesc> blocks;          // ordered by (chunk, lane)
  std::vector<uint32_t> lane_first;       // [world * lanes + 1] first block of slot (chunk*lanes + lane)
  std::vector<uint32_t> chunk_wire_bytes; // [world]
  std::vector<uint64_t> chunk_elems;      // [world]
  uint32_t max_chunk_wire = 0;
  uint64_t total_wire = 0;
  // flattened warp work list (see wire.h), same (chunk, lane) slot order as `blocks`
  std::vector<WarpItem> items;
  std::vector<uint32_t> item_first;       // [world * lanes + 1]
  std::vector<uint32_t> block_item_first; // [blocks + 1]: items of block b are [block_item_first[b], [b+1])
  uint32_t slice_elems = 512;             // elements of a kItemFull item (512, or 1024 when every
                                          // compressed layer uses 1024-element buckets)
  int uniform_bits = 0;                   // bits shared by every compressed block, else 0

  uint32_t slot_begin(int chunk, int lane) const { return lane_first[(size_t)chunk * lanes + lane]; }
  uint32_t slot_end(int chunk, int lane) const { return lane_first[(size_t)chunk * lanes + lane + 1]; }
  uint32_t chunk_begin(int chunk) const { return lane_first[(size_t)chunk * lanes]; }
  uint32_t chunk_end(int chunk) const { return lane_first[(size_t)(chunk + 1) * lanes]; }
};

// Throws std::invalid_argument on malformed input (overlapping layers, bits
// out of range, buffer too large for 32-bit offsets).
Plan build_plan(const std::vector<LayerSpec>& layers, const PlanOptions& opt);

// Hash of everything that determines a plan (cache key).
uint64_t plan_key(const std::vector<LayerSpec>& layers, const PlanOptions& opt);

std::string describe_plan(const Plan& p);

}  // namespace cgx
