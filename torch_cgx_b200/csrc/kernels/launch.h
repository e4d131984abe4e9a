// Host-callable launchers of every CUDA kernel in the framework (no torch
// dependency: these TUs compile in seconds and are what `cuobjdump -sass`
// inspects).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../common/philox.h"
#include "../common/wire.h"

namespace cgx {

// Device-resident call counter of one heap: the epoch of a launch is read from here by every
// CTA and bumped by the last CTA to finish, so a captured CUDA graph replays with fresh epochs
// (kernel arguments are frozen at capture time).
struct DeviceSync {
  uint32_t epoch;     // number of completed fused launches on this heap
  uint32_t finished;  // CTAs of the running launch that are done
  uint32_t pad[2];
};

// ---- fused Scatter-Reduce-AllGather / one-shot over peer memory ------------
struct SraParams {
  void* data;                    // this rank's gradient buffer (T*), reduced in place
  const WarpItem* items;         // device plan: flattened warp items ...
  const uint32_t* item_first;    // ... and the first item of every (chunk, lane) slot [world*lanes + 1]
  int rank;
  int world;
  int lanes;                     // == gridDim.x
  int dtype;
  uint32_t epoch_hint;           // host mirror of the epoch this launch will get (== device epoch
                                 // outside graph replay; only used to advance the RNG sequence)
  float prescale;                // gradients are multiplied by this before anything else (1/W for AVG)
  RngKey rng;                    // .stream is filled in by the kernel
  uint32_t slot_bytes;           // stride between per-source slots inside recv1/recv2
  uint32_t flag_stride;          // lanes capacity of the flag arrays
  uint32_t os_parity_stride;     // one-shot: byte distance between the two alternating regions
  uint8_t* recv1[kMaxPeers];     // recv1[p]: base of rank p's phase-1 receive region (peer-mapped)
  uint8_t* recv2[kMaxPeers];     // recv2[p]: base of rank p's phase-2 receive region
  uint32_t* flags1[kMaxPeers];   // flags1[p][src*flag_stride + lane]
  uint32_t* flags2[kMaxPeers];
  uint8_t* mc_recv1;             // NVLS multicast aliases of the same regions (nullptr: unicast only)
  uint8_t* mc_recv2;
  int mc_reduce;                 // raw items are reduced inside the switch (multimem.ld_reduce)
  uint32_t* status;              // local error word (0 = ok), host-mapped
  const uint32_t* abort_word;    // host-mapped: non-zero = stop waiting for peers (ProcessGroup::abort)
  uint64_t timeout_ns;
  DeviceSync* sync;
  unsigned long long* trace;     // optional [lanes][8] device timestamps (ns, globaltimer), nullptr = off
  int uniform_bits;              // common bits of all compressed items (2/4/8 select a specialised kernel), else 0
  int slice_elems;               // 512 or 1024 (plan.slice_elems)
  int oneshot;                   // 0: three-phase SRA, 1: one-shot
  int stages;                    // SRA: pieces every chunk is pipelined in (1..4, stages * world <= 32)
};

constexpr int kSraThreads = 256;
constexpr int kSraWarps = kSraThreads / 32;
constexpr int kSraCtasPerSm = 2;
enum SraStatus : uint32_t { kSraOk = 0, kSraTimeoutPhase1 = 1, kSraTimeoutPhase2 = 2, kSraAborted = 3 };

// max CTAs of the fused kernel that can be co-resident on the current device
int sra_max_resident_ctas(int dtype);
cudaError_t launch_sra_fused(const SraParams& p, cudaStream_t stream);

// ---- standalone item kernels (generic reducers, Python ops, tests, cross-node path) ----
// All operate on items [first, first+count) of a device item table; `wire` is the base of the
// chunk's wire slot (records live at wire + item offsets). `slice_elems` / `uniform_bits` are
// the plan's.
struct ItemKernelArgs {
  const WarpItem* items;
  uint32_t first, count;
  int dtype;
  int slice_elems;
  int uniform_bits;
};
// wire = quantize(src[T] * prescale)
cudaError_t launch_quantize_items(const ItemKernelArgs& a, const void* src, uint8_t* wire, float prescale,
                                  const RngKey& rng, cudaStream_t stream);
// src is an fp32 scratch laid out like the tensor (index = elem_off - base_elem);
// also writes the self-decoded values into `out` (T, index = elem_off) when out != nullptr.
cudaError_t launch_quantize_items_f32(const ItemKernelArgs& a, const float* src_f32, uint32_t base_elem,
                                      uint8_t* wire, const RngKey& rng, void* out, cudaStream_t stream);
// dst[T] = decode(wire)
cudaError_t launch_dequantize_items(const ItemKernelArgs& a, const uint8_t* wire, void* dst, cudaStream_t stream);
// acc_f32[elem_off - base_elem + i] = float(init_src[elem_off + i]) * prescale   (init_src != nullptr)
// acc_f32[...] += decode(wire)                                                   (wire != nullptr)
cudaError_t launch_accumulate_items_f32(const ItemKernelArgs& a, const uint8_t* wire, float* acc_f32,
                                        uint32_t base_elem, const void* init_src, float prescale,
                                        cudaStream_t stream);

// ---- elementwise helpers (K4/K7 of the reference) ---------------------------
cudaError_t launch_scale_inplace(void* data, int dtype, uint64_t n, float scale, cudaStream_t stream);
cudaError_t launch_add(const void* x, const void* y, void* sum, int dtype, uint64_t n, cudaStream_t stream);
cudaError_t launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, uint64_t n,
                           cudaStream_t stream);

}  // namespace cgx
