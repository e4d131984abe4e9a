// The product: ONE kernel per fusion buffer that performs the whole compressed
// Scatter-Reduce-AllGather allreduce over NVLink/NVSwitch peer memory.
//
//   phase A  for every peer p: per-bucket min/max -> quantize -> pack my copy of
//            chunk p and store it straight into p's receive slot (P2P stores),
//            then release-signal p's lane flag
//   phase B  acquire-wait the W-1 incoming copies of MY chunk, dequantize +
//            accumulate them with my raw values in fp32 (smem), requantize,
//            write the packed result to every peer's gather slot and the
//            self-decoded values to my gradient buffer, release-signal
//   phase C  acquire-wait each peer's reduced chunk, dequantize into my buffer
//
// The grid is G persistent "lanes" (CTAs). Lane c of every rank only ever
// talks to lane c of the other ranks through single-writer epoch flags, so
// there is no grid-wide or host synchronisation anywhere, and lanes pipeline
// independently (lane 3 can be in phase C while lane 90 is still in phase A).
//
// Reference behaviour being replaced (4-5 launches per layer + host polling):
//   /root/reference/src/common/scatter_reduce_allgather.cc:94-202
//   /root/reference/src/common/nccl_reduce.cc:103-198
//   /root/reference/src/common/shm_communicator.cc:110-177
// Uncompressed layers (bits == 32) ride the same launch as "raw" blocks: the
// classic two-shot P2P allreduce with the 1/W prescale fused
// (scatter_reduce_allgather.cc:308-413, allreduce_hooks.py:48-59).
#include "block_device.cuh"
#include "launch.h"

namespace cgx {
using namespace dev;

namespace {

template <typename T>
__global__ void __launch_bounds__(kSraThreads, 2) sra_fused_kernel(const SraParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile& tile = *reinterpret_cast<Tile*>(smem_raw);
  __shared__ int s_abort;

  const int lane = blockIdx.x;
  const int r = p.rank, W = p.world, G = p.lanes;
  const uint32_t tid = threadIdx.x;
  T* data = reinterpret_cast<T*>(p.data);
  if (tid == 0) s_abort = 0;
  __syncthreads();

  // ------------------------------------------------------------------ phase A
  {
    RngKey rng = p.rng;
    rng.stream = (uint32_t)r * 2u;
    for (int s = 1; s < W; ++s) {
      const int dstp = (r + s) % W;
      const uint32_t b0 = p.lane_first[dstp * G + lane], b1 = p.lane_first[dstp * G + lane + 1];
      if (b0 == b1) continue;
      uint8_t* slot = p.recv1[dstp] + (size_t)r * p.slot_bytes;
      for (uint32_t b = b0; b < b1; ++b) {
        const BlockDesc d = p.blocks[b];
        uint8_t* rec = slot + d.wire_off;
        if (block_is_raw(d)) {
          send_raw<T>(data, d, p.prescale, rec);
          continue;
        }
        const uint32_t n = block_n(d);
        const int bits = block_bits(d);
        load_block<T>(data, d, p.prescale, tile.acc);
        __syncthreads();
        compute_meta(tile.acc, n, d.bucket, bits, tile.meta, tile.inv);
        __syncthreads();
        pack_block<T, false>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, nullptr);
        __syncthreads();
        store_record(tile.meta, tile.pay, block_meta_bytes(n, d.bucket), block_payload_bytes(n, bits), &rec, 1);
        __syncthreads();
      }
      __syncthreads();
      if (tid == 0) st_release_sys(p.flags1[dstp] + (size_t)r * p.flag_stride + lane, p.epoch);
    }
  }

  // ------------------------------------------------------------------ phase B
  {
    const uint32_t b0 = p.lane_first[r * G + lane], b1 = p.lane_first[r * G + lane + 1];
    if (b1 > b0) {
      if (tid < (uint32_t)W && (int)tid != r) {
        if (!wait_flag(p.flags1[r] + (size_t)tid * p.flag_stride + lane, p.epoch, p.timeout_ns)) {
          s_abort = 1;
          *p.status = kSraTimeoutPhase1 | ((uint32_t)tid << 8) | ((uint32_t)lane << 16);
        }
      }
      __syncthreads();
      if (s_abort) return;

      RngKey rng = p.rng;
      rng.stream = (uint32_t)r * 2u + 1u;
      const uint8_t* src_rec[kMaxPeers];
      uint8_t* dst_rec[kMaxPeers];
      for (uint32_t b = b0; b < b1; ++b) {
        const BlockDesc d = p.blocks[b];
        int np = 0;
        for (int q = 0; q < W; ++q) {
          if (q == r) continue;
          src_rec[np] = p.recv1[r] + (size_t)q * p.slot_bytes + d.wire_off;
          dst_rec[np] = p.recv2[q] + (size_t)r * p.slot_bytes + d.wire_off;
          ++np;
        }
        if (block_is_raw(d)) {
          reduce_raw<T>(data, d, p.prescale, src_rec, np, dst_rec, np);
          continue;
        }
        const uint32_t n = block_n(d);
        const int bits = block_bits(d);
        load_block<T>(data, d, p.prescale, tile.acc);
        __syncthreads();
        for (int k = 0; k < np; ++k) decode_add<T>(src_rec[k], d, tile.acc);
        __syncthreads();
        compute_meta(tile.acc, n, d.bucket, bits, tile.meta, tile.inv);
        __syncthreads();
        pack_block<T, true>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, data);
        __syncthreads();
        store_record(tile.meta, tile.pay, block_meta_bytes(n, d.bucket), block_payload_bytes(n, bits), dst_rec, np);
        __syncthreads();
      }
      __syncthreads();
      if (tid < (uint32_t)W && (int)tid != r)
        st_release_sys(p.flags2[tid] + (size_t)r * p.flag_stride + lane, p.epoch);
    }
  }

  // ------------------------------------------------------------------ phase C
  for (int s = 1; s < W; ++s) {
    const int q = (r + s) % W;
    const uint32_t b0 = p.lane_first[q * G + lane], b1 = p.lane_first[q * G + lane + 1];
    if (b0 == b1) continue;
    if (tid == 0) {
      if (!wait_flag(p.flags2[r] + (size_t)q * p.flag_stride + lane, p.epoch, p.timeout_ns)) {
        s_abort = 1;
        *p.status = kSraTimeoutPhase2 | ((uint32_t)q << 8) | ((uint32_t)lane << 16);
      }
    }
    __syncthreads();
    if (s_abort) return;
    const uint8_t* slot = p.recv2[r] + (size_t)q * p.slot_bytes;
    for (uint32_t b = b0; b < b1; ++b) {
      const BlockDesc d = p.blocks[b];
      decode_store<T>(slot + d.wire_off, d, data);
    }
  }
}

template <typename T>
cudaError_t launch_t(const SraParams& p, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(sra_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Tile));
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  sra_fused_kernel<T><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p);
  return cudaGetLastError();
}

template <typename T>
int max_resident_t() {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(sra_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sra_fused_kernel<T>, kSraThreads, sizeof(Tile)) !=
      cudaSuccess)
    return 0;
  return sms * per_sm;
}

}  // namespace

int sra_max_resident_ctas(int dtype) {
  switch (dtype) {
    case kF32: return max_resident_t<float>();
    case kF16: return max_resident_t<__half>();
    default: return max_resident_t<__nv_bfloat16>();
  }
}

cudaError_t launch_sra_fused(const SraParams& p, cudaStream_t stream) {
  switch (p.dtype) {
    case kF32: return launch_t<float>(p, stream);
    case kF16: return launch_t<__half>(p, stream);
    case kBF16: return launch_t<__nv_bfloat16>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cgx
