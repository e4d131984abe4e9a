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

constexpr uint32_t kMaxBlockElems = 8192;   // fp32 accumulator tile = 32 KB of smem
constexpr uint32_t kMaxBlockBuckets = 512;  // meta tile = 4 KB of smem
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

CGX_HD uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
CGX_HD uint32_t div_up(uint32_t x, uint32_t a) { return (x + a - 1) / a; }

CGX_HD uint32_t block_num_buckets(uint32_t n, uint32_t bucket) { return div_up(n, bucket); }
CGX_HD uint32_t block_meta_bytes(uint32_t n, uint32_t bucket) {
  return align_up(block_num_buckets(n, bucket) * 8u, kWireAlign);
}
CGX_HD uint32_t block_payload_bytes(uint32_t n, int bits) {
  // whole groups of 8 values -> `bits` bytes each (a partial last group still
  // occupies a full `bits` bytes), then 16 B padding
  return align_up(div_up(n, 8u) * (uint32_t)bits, kWireAlign);
}
CGX_HD uint32_t block_wire_bytes(uint32_t n, int bits, uint32_t bucket, int elsize) {
  if (bits >= kRawBits) return align_up(n * (uint32_t)elsize, kWireAlign);
  return block_meta_bytes(n, bucket) + block_payload_bytes(n, bits);
}

}  // namespace cgx
