// Wire format of the compressed allreduce and the "block" work unit.
//
// The unit of work of every kernel (fused P2P kernel, standalone
// quantize/dequantize kernels) and of the CPU path is a *block*: a run of at
// most kMaxBlockElems consecutive elements that lies inside ONE registered
// layer, starts on a quantization-bucket boundary of that layer and therefore
// has a single (bits, bucket_size) configuration.
//
// Block record on the wire (all sections padded to 16 B so every record can be
// moved with 128-bit / TMA bulk accesses):
//   compressed: [ nb x {float unit, float min} ][ ceil(n*bits/8) payload bytes ]
//   raw       : [ n x T ]
// Payload packing is the reference's: groups of 8 consecutive values -> `bits`
// little-endian bytes (/root/reference/src/common/compression/
// cuda_compression_operations.cu:287-371); the meta is fp32
// here (the reference stores it in T).
#pragma once
#include <cstdint>
#include "quant_math.h"

namespace cgx {

enum DType : int32_t { kF32 = 0, kF16 = 1, kBF16 = 2 };

CGX_HD int dtype_size(int dt) { return dt == kF32 ? 4 : 2; }

constexpr uint32_t kMaxBlockElems = 8192;   // also the largest quantization bucket
constexpr uint32_t kMaxBlockBuckets = 512;
constexpr uint32_t kWireAlign = 16;
constexpr int kMaxPeers = 16;

struct alignas(16) BlockDesc {
  uint32_t elem_off;  // element offset from the tensor base
  uint32_t wire_off;  // byte offset of the record inside the chunk's wire slot
  uint32_t n_bits;    // n in the low 24 bits, bits in the high 8 (32 => raw)
  uint32_t bucket;    // quantization bucket size in elements
};

CGX_HD uint32_t block_n(const BlockDesc& b) { return b.n_bits & 0xFFFFFFu; }
CGX_HD int block_bits(const BlockDesc& b) { return (int)(b.n_bits >> 24); }
CGX_HD bool block_is_raw(const BlockDesc& b) { return block_bits(b) >= kRawBits; }

// ---- warp work items --------------------------------------------------------
// The fused kernels do not walk blocks: the plan is flattened ONCE on the host into
// *warp items*, the unit one warp processes end to end in registers. The wire layout
// is still the block record above; an item only carries pre-resolved offsets into it.
//   kItemFull    one slice of `slice_elems` (512 or 1024) elements made of whole buckets whose
//                size is a power of two in [8, slice_elems] -> no per-element predicates; 16 B
//                vector access is decided at run time from the actual pointers
//   kItemBucket  ONE bucket (whole or partial, any size <= kMaxBlockElems): generic two-pass path
//   kItemRaw     512 uncompressed elements, all present
//   kItemRawTail < 512 uncompressed elements
enum ItemKind : uint32_t { kItemFull = 0, kItemBucket = 1, kItemRaw = 2, kItemRawTail = 3 };
constexpr uint32_t kRawItemElems = 512;

struct alignas(16) WarpItem {
  uint32_t elem_off;  // first element, from the tensor base
  uint32_t meta_off;  // byte offset inside the chunk's wire slot of the item's first {unit,min}
                      // (raw items: of its first element)
  uint32_t pay_off;   // byte offset of the item's first packed group
  uint32_t info;      // kind [0:1] | log2(bucket/8) [3:5] | bits [8:15] | n [16:31]
};
CGX_HD uint32_t item_kind(const WarpItem& it) { return it.info & 3u; }
CGX_HD uint32_t item_lpb_log2(const WarpItem& it) { return (it.info >> 3) & 7u; }
CGX_HD int item_bits(const WarpItem& it) { return (int)((it.info >> 8) & 0xFFu); }
CGX_HD uint32_t item_n(const WarpItem& it) { return it.info >> 16; }
CGX_HD uint32_t make_item_info(uint32_t kind, uint32_t lpb_log2, int bits, uint32_t n) {
  return kind | (lpb_log2 << 3) | ((uint32_t)bits << 8) | (n << 16);
}

CGX_HD uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
CGX_HD uint32_t div_up(uint32_t x, uint32_t a) { return (x + a - 1) / a; }

CGX_HD uint32_t block_num_buckets(uint32_t n, uint32_t bucket) { return div_up(n, bucket); }
CGX_HD uint32_t block_meta_bytes(uint32_t n, uint32_t bucket) {
  return align_up(block_num_buckets(n, bucket) * 8u, kWireAlign);
}
// Pack groups (8 values -> `bits` bytes) never straddle a bucket: every bucket of a block starts
// a fresh group, so a bucket is a self-contained unit of work (one warp item). For bucket sizes
// that are multiples of 8 -- all the usual ones -- this is exactly the reference's global packing.
CGX_HD uint32_t bucket_groups(uint32_t bucket) { return div_up(bucket, 8u); }
CGX_HD uint32_t block_num_groups(uint32_t n, uint32_t bucket) {
  const uint32_t full = n / bucket;
  return full * bucket_groups(bucket) + div_up(n - full * bucket, 8u);
}
CGX_HD uint32_t block_payload_bytes(uint32_t n, int bits, uint32_t bucket) {
  return align_up(block_num_groups(n, bucket) * (uint32_t)bits, kWireAlign);
}
CGX_HD uint32_t block_wire_bytes(uint32_t n, int bits, uint32_t bucket, int elsize) {
  if (bits >= kRawBits) return align_up(n * (uint32_t)elsize, kWireAlign);
  return block_meta_bytes(n, bucket) + block_payload_bytes(n, bits, bucket);
}

}  // namespace cgx
