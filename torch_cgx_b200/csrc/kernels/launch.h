// Host-callable launchers of every CUDA kernel in the framework (no torch
// dependency: these TUs compile in seconds and are what `cuobjdump -sass`
// inspects).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../common/philox.h"
#include "../common/wire.h"

namespace cgx {

// ---- fused Scatter-Reduce-AllGather over peer memory ----------------------
struct SraParams {
  void* data;                    // this rank's gradient buffer (T*), reduced in place
  const BlockDesc* blocks;       // device plan table
  const uint32_t* lane_first;    // [world*lanes + 1]
  int rank;
  int world;
  int lanes;                     // == gridDim.x
  int dtype;
  uint32_t epoch;                // call sequence number (>= 1), identical on all ranks
  float prescale;                // gradients are multiplied by this before anything else (1/W for AVG)
  RngKey rng;                    // .stream is filled in by the kernel
  uint32_t slot_bytes;           // stride between per-source slots inside recv1/recv2
  uint32_t flag_stride;          // lanes capacity of the flag arrays
  uint8_t* recv1[kMaxPeers];     // recv1[p]: base of rank p's phase-1 receive region (peer-mapped)
  uint8_t* recv2[kMaxPeers];     // recv2[p]: base of rank p's phase-2 receive region
  uint32_t* flags1[kMaxPeers];   // flags1[p][src*flag_stride + lane]
  uint32_t* flags2[kMaxPeers];
  uint32_t* status;              // local error word (0 = ok)
  uint64_t timeout_ns;
  int variant;                   // 0 = warp-centric kernel (default), 1 = CTA/shared-memory kernel (v1)
  unsigned long long* trace;     // optional [lanes][8] device timestamps (ns, globaltimer), nullptr = off
  int uniform_bits;              // common bits of all compressed blocks (2/4/8 select a specialised kernel), else 0
};

constexpr int kSraThreads = 512;
enum SraStatus : uint32_t { kSraOk = 0, kSraTimeoutPhase1 = 1, kSraTimeoutPhase2 = 2 };

// max CTAs of the fused kernel that can be co-resident on the current device
int sra_max_resident_ctas(int dtype);
cudaError_t launch_sra_fused(const SraParams& p, cudaStream_t stream);

// ---- standalone block kernels (generic reducers, tests, cross-node path) ---
// All operate on blocks [first, first+count) of a device plan table; `wire`
// is the base of the chunk's wire slot (records live at wire + desc.wire_off).
cudaError_t launch_quantize_blocks(const void* src, int dtype, const BlockDesc* blocks, uint32_t first,
                                   uint32_t count, uint8_t* wire, float prescale, const RngKey& rng,
                                   cudaStream_t stream);
// src is an fp32 scratch laid out like the tensor (index = desc.elem_off + i - base_elem);
// also writes the self-decoded values into `out` (T) when out != nullptr.
cudaError_t launch_quantize_blocks_f32(const float* src_f32, uint32_t base_elem, int dtype,
                                       const BlockDesc* blocks, uint32_t first, uint32_t count,
                                       uint8_t* wire, const RngKey& rng, void* out, cudaStream_t stream);
// dst[T] = decode(wire)
cudaError_t launch_dequantize_blocks(const uint8_t* wire, int dtype, const BlockDesc* blocks, uint32_t first,
                                     uint32_t count, void* dst, cudaStream_t stream);
// acc_f32[elem_off - base_elem + i] (+)= decode(wire) ; if init_src != nullptr first
// acc = float(init_src[elem_off + i]) * prescale
cudaError_t launch_accumulate_blocks_f32(const uint8_t* wire, int dtype, const BlockDesc* blocks,
                                         uint32_t first, uint32_t count, float* acc_f32, uint32_t base_elem,
                                         const void* init_src, float prescale, cudaStream_t stream);

// ---- elementwise helpers (K4/K7 of the reference) ---------------------------
cudaError_t launch_scale_inplace(void* data, int dtype, uint64_t n, float scale, cudaStream_t stream);
cudaError_t launch_add(const void* x, const void* y, void* sum, int dtype, uint64_t n, cudaStream_t stream);
cudaError_t launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, uint64_t n,
                           cudaStream_t stream);

}  // namespace cgx
