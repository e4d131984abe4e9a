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
#include "warp_path.cuh"

#include <type_traits>

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
    }
    // one release per lane for the whole phase (see the warp kernel)
    __syncthreads();
    if (tid < (uint32_t)W && (int)tid != r && p.lane_first[tid * G + lane] != p.lane_first[tid * G + lane + 1])
      st_release_sys(p.flags1[tid] + (size_t)r * p.flag_stride + lane, p.epoch);
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


// ===========================================================================
// v2: warp-centric kernel. Same protocol and bit-identical results, but the
// unit of scheduling is a warp "item" (one quantization bucket, or 2048 raw
// elements): items of a lane are dealt round-robin to the CTA's warps, which
// run without any block-wide barrier; chunk completion is tracked with
// shared-memory counters and published by whichever warp finishes last.
// ===========================================================================
constexpr uint32_t kNumWarps = kSraThreads / 32;

// Tracing: per-lane phase timestamps (slot 0: kernel start [min], 1: phase A done,
// 2: phase B inputs arrived, 3: phase B done, 4: last phase-C wait satisfied, 5: end [max]).
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int lane, int slot, bool is_min = false) {
  if (trace == nullptr || (threadIdx.x & 31u) != 0) return;
  const unsigned long long t = globaltimer_ns();
  if (is_min)
    atomicMin(&trace[(size_t)lane * 8 + slot], t);
  else
    atomicMax(&trace[(size_t)lane * 8 + slot], t);
}
#define CGX_INF_POS __int_as_float(0x7f800000)
#define CGX_INF_NEG __int_as_float(0xff800000)

// phase A: my copy of one bucket of a peer's chunk -> quantize -> peer's slot
template <typename T, bool FULL, int KB>
__device__ __forceinline__ void bucket_send(const T* __restrict__ blk, bool aligned, const BlockDesc& d, uint32_t bk,
                                            float prescale, const RngKey& rng, uint32_t b, uint8_t* const* recs,
                                            int nrec) {
  const int bits = KB ? KB : block_bits(d);
  const uint32_t meta_bytes = block_meta_bytes(block_n(d), d.bucket);
  const uint32_t ns = div_up(bucket_count(d, bk), kSliceElems);
  float x[kMaxGpl][8];
  float mn = CGX_INF_POS, mx = CGX_INF_NEG;
  BucketCtx c;
  for (uint32_t sl = 0; sl < ns; ++sl) {
    c = make_slice_ctx(d, bk, sl);
    warp_load_bucket<T, FULL>(blk, aligned, c, prescale, x);
    warp_minmax_update<FULL>(x, c, mn, mx);
  }
  const BucketMeta m = warp_minmax_finish(mn, mx, bits);
  warp_store_meta(m, bk, recs, nrec);
  if (ns == 1) {
    warp_quantize_store<T, false, FULL>(x, c, m, bits, meta_bytes, rng, b, recs, nrec, nullptr, false);
  } else {
    for (uint32_t sl = 0; sl < ns; ++sl) {
      c = make_slice_ctx(d, bk, sl);
      warp_load_bucket<T, FULL>(blk, aligned, c, prescale, x);
      warp_quantize_store<T, false, FULL>(x, c, m, bits, meta_bytes, rng, b, recs, nrec, nullptr, false);
    }
  }
}

// own slice + the W-1 decoded peer copies, summed in fixed rank order
template <typename T, bool FULL, int KB, int kPeerBatch>
__device__ __forceinline__ void slice_gather(T* __restrict__ blk, bool aligned, const BlockDesc& d, uint32_t bk,
                                             uint32_t sl, float prescale, uint32_t meta_bytes, int bits,
                                             const uint8_t* const* src_rec, int np, BucketCtx& c,
                                             float (&x)[kMaxGpl][8]) {
  c = make_slice_ctx(d, bk, sl);
  warp_load_bucket<T, FULL>(blk, aligned, c, prescale, x);
  // peers in batches: all loads of a batch are issued before any is consumed. With <= 4 bits a
  // group's word is 32 bits, so twice as many peers fit in the same register budget.
  constexpr bool kNarrow = (KB > 0 && KB <= 4);
  using word_t = typename std::conditional<kNarrow, uint32_t, uint64_t>::type;
  constexpr int kBatch = kNarrow ? kPeerBatch : kPeerBatch;
  for (int q0 = 0; q0 < np; q0 += kBatch) {
    word_t w[kBatch][kMaxGpl];
    BucketMeta pm[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      if (q0 + u < np) warp_fetch_peer<FULL, word_t>(src_rec[q0 + u], meta_bytes, bk, bits, c, w[u], pm[u]);
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      if (q0 + u < np) warp_accumulate<FULL, word_t>(w[u], pm[u], bits, c, x);
  }
}

// phase B: reduce one bucket of MY chunk, requantize, publish to every peer, self-decode
template <typename T, bool FULL, int KB, int kPeerBatch>
__device__ __forceinline__ void bucket_reduce(T* __restrict__ blk, bool aligned, const BlockDesc& d, uint32_t bk,
                                              float prescale, const RngKey& rng, uint32_t b,
                                              const uint8_t* const* src_rec, uint8_t* const* dst_rec, int np) {
  const int bits = KB ? KB : block_bits(d);
  const uint32_t meta_bytes = block_meta_bytes(block_n(d), d.bucket);
  const uint32_t ns = div_up(bucket_count(d, bk), kSliceElems);
  float x[kMaxGpl][8];
  float mn = CGX_INF_POS, mx = CGX_INF_NEG;
  BucketCtx c;
  for (uint32_t sl = 0; sl < ns; ++sl) {
    slice_gather<T, FULL, KB, kPeerBatch>(blk, aligned, d, bk, sl, prescale, meta_bytes, bits, src_rec, np, c, x);
    warp_minmax_update<FULL>(x, c, mn, mx);
  }
  const BucketMeta m = warp_minmax_finish(mn, mx, bits);
  warp_store_meta(m, bk, dst_rec, np);
  if (ns == 1) {
    warp_quantize_store<T, true, FULL>(x, c, m, bits, meta_bytes, rng, b, dst_rec, np, blk, aligned);
  } else {
    // the self-decode of slice sl overwrites only slice sl of my own gradient, which later
    // slices never re-read
    for (uint32_t sl = 0; sl < ns; ++sl) {
      slice_gather<T, FULL, KB, kPeerBatch>(blk, aligned, d, bk, sl, prescale, meta_bytes, bits, src_rec, np, c, x);
      warp_quantize_store<T, true, FULL>(x, c, m, bits, meta_bytes, rng, b, dst_rec, np, blk, aligned);
    }
  }
}

// phase C: a peer's reduced bucket -> my gradient buffer
template <typename T, bool FULL, int KB>
__device__ __forceinline__ void bucket_recv(const uint8_t* rec, T* __restrict__ blk, bool aligned, const BlockDesc& d,
                                            uint32_t bk) {
  const int bits = KB ? KB : block_bits(d);
  const uint32_t meta_bytes = block_meta_bytes(block_n(d), d.bucket);
  const uint32_t ns = div_up(bucket_count(d, bk), kSliceElems);
  for (uint32_t sl = 0; sl < ns; ++sl) {
    const BucketCtx c = make_slice_ctx(d, bk, sl);
    warp_decode_store<T, FULL>(rec, meta_bytes, bk, bits, c, blk, aligned);
  }
}

__device__ __forceinline__ bool bucket_is_full(const BlockDesc& d, uint32_t bk, bool aligned) {
  return aligned && (bucket_count(d, bk) % kSliceElems) == 0;
}

// KB: compile-time quantization bits when every compressed block of the plan uses
// the same width (the normal case) -- shifts/masks become immediates and the
// per-width switches fold away; KB == 0 keeps them as run-time values.
// kPeerBatch / kMinBlocks trade phase-B memory-level parallelism against occupancy:
//   <2,1>: 128 registers, 16 warps/SM, two peers' words in flight per bucket
//   <1,2>:  64 registers, 32 warps/SM (2 CTAs/SM), one peer at a time
template <typename T, int KB, int kPeerBatch, int kMinBlocks>
__global__ void __launch_bounds__(kSraThreads, kMinBlocks) sra_fused_warp_kernel(const SraParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile& tile = *reinterpret_cast<Tile*>(smem_raw);
  __shared__ uint32_t s_expect[kMaxPeers];
  __shared__ uint32_t s_done_a;   // items of phase A finished by this CTA (all destinations)
  __shared__ uint32_t s_total_a;  // items of phase A in this lane
  __shared__ uint32_t s_done_b;
  __shared__ int s_abort;

  const int lane = blockIdx.x;
  const int r = p.rank, W = p.world, G = p.lanes;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, wl = tid & 31u;
  T* data = reinterpret_cast<T*>(p.data);

  if (tid < (uint32_t)W) {
    uint32_t tot = 0;
    for (uint32_t b = p.lane_first[tid * G + lane]; b < p.lane_first[tid * G + lane + 1]; ++b)
      tot += block_items(p.blocks[b]);
    s_expect[tid] = tot;
  }
  if (tid == 0) {
    s_done_a = 0;
    s_done_b = 0;
    s_abort = 0;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t tot = 0;
    for (int q = 0; q < W; ++q)
      if (q != r) tot += s_expect[q];
    s_total_a = tot;
  }
  __syncthreads();
  trace_mark(p.trace, lane, 0, true);

  // ------------------------------------------------------------------ phase A
  {
    RngKey rng = p.rng;
    rng.stream = (uint32_t)r * 2u;
    uint32_t base = 0;  // items dealt so far: item i goes to warp (i % kNumWarps)
    const uint32_t total_a = s_total_a;
    // ONE system-scope release per lane for the whole phase: a fence.sys drains every
    // outstanding NVLink store of the SM (microseconds), and the consumer needs all W-1
    // sources before it can start phase B anyway, so per-destination signalling buys nothing.
    auto item_done_a = [&]() {
      __syncwarp();
      uint32_t last = 0;
      if (wl == 0) last = (atom_add_acq_rel_cta(&s_done_a, 1u) + 1u == total_a) ? 1u : 0u;
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last && wl < (uint32_t)W && (int)wl != r && s_expect[wl] > 0)
        st_release_sys(p.flags1[wl] + (size_t)r * p.flag_stride + lane, p.epoch);
    };
    for (int s = 1; s < W; ++s) {
      const int dstp = (r + s) % W;
      const uint32_t b0 = p.lane_first[dstp * G + lane], b1 = p.lane_first[dstp * G + lane + 1];
      if (b0 == b1) continue;
      uint8_t* slot = p.recv1[dstp] + (size_t)r * p.slot_bytes;
      for (uint32_t b = b0; b < b1; ++b) {
        const BlockDesc d = p.blocks[b];
        uint8_t* rec = slot + d.wire_off;
        const uint32_t n = block_n(d);
        const T* blk = data + d.elem_off;
        const bool aligned = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
        const uint32_t first = (warp - base) & (kNumWarps - 1);
        if (block_is_raw(d)) {
          const uint32_t ni = div_up(n, kRawItemElems);
          for (uint32_t i = first; i < ni; i += kNumWarps) {
            warp_send_raw<T>(blk, aligned, n, i, p.prescale, rec);
            item_done_a();
          }
          base += ni;
        } else if (block_is_fast(d)) {
          const uint32_t nb = block_num_buckets(n, d.bucket);
          for (uint32_t bk = first; bk < nb; bk += kNumWarps) {
            if (bucket_is_full(d, bk, aligned))
              bucket_send<T, true, KB>(blk, aligned, d, bk, p.prescale, rng, b, &rec, 1);
            else
              bucket_send<T, false, KB>(blk, aligned, d, bk, p.prescale, rng, b, &rec, 1);
            item_done_a();
          }
          base += nb;
        } else {
          const int bits = block_bits(d);
          load_block<T>(data, d, p.prescale, tile.acc);
          __syncthreads();
          compute_meta(tile.acc, n, d.bucket, bits, tile.meta, tile.inv);
          __syncthreads();
          pack_block<T, false>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, nullptr);
          __syncthreads();
          store_record(tile.meta, tile.pay, block_meta_bytes(n, d.bucket), block_payload_bytes(n, bits), &rec, 1);
          __syncthreads();
          if (warp == 0) item_done_a();
        }
      }
    }
  }

  trace_mark(p.trace, lane, 1);
  // ------------------------------------------------------------------ phase B
  {
    const uint32_t b0 = p.lane_first[r * G + lane], b1 = p.lane_first[r * G + lane + 1];
    if (b1 > b0) {
      bool ok = true;
      if (wl < (uint32_t)W && (int)wl != r) {
        ok = wait_flag(p.flags1[r] + (size_t)wl * p.flag_stride + lane, p.epoch, p.timeout_ns);
        if (!ok) {
          s_abort = 1;
          *p.status = kSraTimeoutPhase1 | (wl << 8) | ((uint32_t)lane << 16);
        }
      }
      if (!__all_sync(0xffffffffu, ok)) return;
      trace_mark(p.trace, lane, 2);

      RngKey rng = p.rng;
      rng.stream = (uint32_t)r * 2u + 1u;
      const uint32_t expect = s_expect[r];
      const uint8_t* src_rec[kMaxPeers];
      uint8_t* dst_rec[kMaxPeers];
      uint32_t base = 0;
      for (uint32_t b = b0; b < b1; ++b) {
        const BlockDesc d = p.blocks[b];
        int np = 0;
        for (int q = 0; q < W; ++q) {
          if (q == r) continue;
          src_rec[np] = p.recv1[r] + (size_t)q * p.slot_bytes + d.wire_off;
          dst_rec[np] = p.recv2[q] + (size_t)r * p.slot_bytes + d.wire_off;
          ++np;
        }
        const uint32_t n = block_n(d);
        T* blk = data + d.elem_off;
        const bool aligned = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
        const uint32_t first = (warp - base) & (kNumWarps - 1);
        bool finished = false;  // did this warp complete the chunk?
        if (block_is_raw(d)) {
          const uint32_t ni = div_up(n, kRawItemElems);
          for (uint32_t i = first; i < ni; i += kNumWarps) {
            warp_reduce_raw<T>(blk, aligned, n, i, p.prescale, src_rec, np, dst_rec, np);
            __syncwarp();
            uint32_t last = 0;
            if (wl == 0) last = (atom_add_acq_rel_cta(&s_done_b, 1u) + 1u == expect) ? 1u : 0u;
            finished |= __shfl_sync(0xffffffffu, last, 0) != 0;
          }
          base += ni;
        } else if (block_is_fast(d)) {
          const uint32_t nb = block_num_buckets(n, d.bucket);
          for (uint32_t bk = first; bk < nb; bk += kNumWarps) {
            if (bucket_is_full(d, bk, aligned))
              bucket_reduce<T, true, KB, kPeerBatch>(blk, aligned, d, bk, p.prescale, rng, b, src_rec, dst_rec, np);
            else
              bucket_reduce<T, false, KB, kPeerBatch>(blk, aligned, d, bk, p.prescale, rng, b, src_rec, dst_rec, np);
            __syncwarp();
            uint32_t last = 0;
            if (wl == 0) last = (atom_add_acq_rel_cta(&s_done_b, 1u) + 1u == expect) ? 1u : 0u;
            finished |= __shfl_sync(0xffffffffu, last, 0) != 0;
          }
          base += nb;
        } else {
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
          uint32_t last = 0;
          if (tid == 0) last = (atom_add_acq_rel_cta(&s_done_b, 1u) + 1u == expect) ? 1u : 0u;
          if (warp == 0) finished |= __shfl_sync(0xffffffffu, last, 0) != 0;
        }
        if (finished && wl < (uint32_t)W && (int)wl != r)
          st_release_sys(p.flags2[wl] + (size_t)r * p.flag_stride + lane, p.epoch);
      }
    }
  }

  trace_mark(p.trace, lane, 3);
  // ------------------------------------------------------------------ phase C
  {
    uint32_t base = 0;
    for (int s = 1; s < W; ++s) {
      const int q = (r + s) % W;
      const uint32_t b0 = p.lane_first[q * G + lane], b1 = p.lane_first[q * G + lane + 1];
      if (b0 == b1) continue;
      const uint32_t* flag = p.flags2[r] + (size_t)q * p.flag_stride + lane;
      const uint8_t* slot = p.recv2[r] + (size_t)q * p.slot_bytes;
      bool waited = false;
      auto ensure = [&]() -> bool {
        if (waited) return true;
        uint32_t ok = 1;
        if (wl == 0) {
          ok = wait_flag(flag, p.epoch, p.timeout_ns) ? 1u : 0u;
          if (!ok) {
            s_abort = 1;
            *p.status = kSraTimeoutPhase2 | ((uint32_t)q << 8) | ((uint32_t)lane << 16);
          }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        waited = true;
        trace_mark(p.trace, lane, 4);
        return ok != 0;
      };
      for (uint32_t b = b0; b < b1; ++b) {
        const BlockDesc d = p.blocks[b];
        const uint8_t* rec = slot + d.wire_off;
        const uint32_t n = block_n(d);
        T* blk = data + d.elem_off;
        const bool aligned = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
        const uint32_t first = (warp - base) & (kNumWarps - 1);
        if (block_is_raw(d)) {
          const uint32_t ni = div_up(n, kRawItemElems);
          for (uint32_t i = first; i < ni; i += kNumWarps) {
            if (!ensure()) return;
            warp_copy_raw<T>(rec, blk, aligned, n, i);
          }
          base += ni;
        } else if (block_is_fast(d)) {
          const uint32_t nb = block_num_buckets(n, d.bucket);
          for (uint32_t bk = first; bk < nb; bk += kNumWarps) {
            if (!ensure()) return;
            if (bucket_is_full(d, bk, aligned))
              bucket_recv<T, true, KB>(rec, blk, aligned, d, bk);
            else
              bucket_recv<T, false, KB>(rec, blk, aligned, d, bk);
          }
          base += nb;
        } else {
          if (!ensure()) return;
          decode_store<T>(rec, d, data);
        }
      }
    }
  }
  trace_mark(p.trace, lane, 5);
}


// ===========================================================================
// One-shot allreduce for small / latency-bound messages: ONE signalling hop.
// Every rank quantizes its WHOLE buffer once and stores the packed image into a
// dedicated slot of every rank (itself included); after one flag exchange every
// rank decodes all W images in rank order and sums them in fp32. Replicas are
// bit-identical (same bytes, same order) and each contribution is quantized
// exactly once (no requantization round). Costs W/2 x the wire bytes of SRA,
// which is irrelevant below ~1 MB where latency dominates. Raw layers travel
// as T. (SURVEY.md build plan step 5: "one-shot (small, latency-bound)";
// replaces the reference's tiny-tensor all-to-all, reducer.cc:35-94.)
// ===========================================================================
template <typename T, bool FULL, int KB>
__device__ __forceinline__ void bucket_sum_recv(const uint8_t* const* src_rec, int W, T* __restrict__ blk,
                                                bool aligned, const BlockDesc& d, uint32_t bk) {
  const int bits = KB ? KB : block_bits(d);
  const uint32_t meta_bytes = block_meta_bytes(block_n(d), d.bucket);
  const uint32_t ns = div_up(bucket_count(d, bk), kSliceElems);
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t sl = 0; sl < ns; ++sl) {
    const BucketCtx c = make_slice_ctx(d, bk, sl);
    float x[kMaxGpl][8];
#pragma unroll
    for (int k = 0; k < kMaxGpl; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[k][j] = 0.f;
    for (int q0 = 0; q0 < W; q0 += 2) {
      uint64_t w[2][kMaxGpl];
      BucketMeta pm[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (q0 + u < W) warp_fetch_peer<FULL, uint64_t>(src_rec[q0 + u], meta_bytes, bk, bits, c, w[u], pm[u]);
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (q0 + u < W) warp_accumulate<FULL, uint64_t>(w[u], pm[u], bits, c, x);
    }
#pragma unroll
    for (int k = 0; k < kMaxGpl; ++k) {
      if (!FULL && c.nv[k] == 0) continue;
      T* o = blk + c.e0 + ((uint32_t)k * 32u + lane) * 8u;
      store_group_values<T>(o, x[k], FULL || (c.nv[k] == 8 && aligned), c.nv[k]);
    }
  }
}

// raw item: out = T( sum_q float(rec_q[i]) ) in rank order
template <typename T>
__device__ __forceinline__ void warp_sum_raw(const uint8_t* const* src_rec, int W, T* __restrict__ blk, bool aligned,
                                             uint32_t n, uint32_t it) {
  constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lo = it * kRawItemElems;
  float f[U][V];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int k = 0; k < V; ++k) f[u][k] = 0.f;
  for (int q = 0; q < W; ++q) {
    uint4 pw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
      pw[u] = (i0 < n) ? ld_sys_v4(src_rec[q] + (size_t)i0 * sizeof(T)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float g[V];
      unpack16<T>(pw[u], g);
#pragma unroll
      for (int k = 0; k < V; ++k) f[u][k] += g[k];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
    if (i0 < n) raw_store_own<T>(blk, aligned, n, i0, pack16<T>(f[u]));
  }
}

template <typename T, int KB>
__global__ void __launch_bounds__(kSraThreads, 1) oneshot_kernel(const SraParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile& tile = *reinterpret_cast<Tile*>(smem_raw);
  __shared__ uint32_t s_total, s_done;
  __shared__ int s_abort;
  const int lane = blockIdx.x;
  const int r = p.rank, W = p.world;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, wl = tid & 31u;
  T* data = reinterpret_cast<T*>(p.data);
  // the plan has ONE chunk (the whole buffer): this lane's blocks
  const uint32_t b0 = p.lane_first[lane], b1 = p.lane_first[lane + 1];
  if (tid == 0) {
    uint32_t tot = 0;
    for (uint32_t b = b0; b < b1; ++b) tot += block_items(p.blocks[b]);
    s_total = tot;
    s_done = 0;
    s_abort = 0;
  }
  __syncthreads();
  if (b0 == b1) return;
  const uint32_t total = s_total;
  RngKey rng = p.rng;
  rng.stream = (uint32_t)r * 2u;

  // ---- phase 1: my quantized image -> slot r of every rank (including mine)
  {
    uint8_t* dst[kMaxPeers];
    auto item_done = [&]() {
      __syncwarp();
      uint32_t last = 0;
      if (wl == 0) last = (atom_add_acq_rel_cta(&s_done, 1u) + 1u == total) ? 1u : 0u;
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last && wl < (uint32_t)W && (int)wl != r)
        st_release_sys(p.flags1[wl] + (size_t)r * p.flag_stride + lane, p.epoch);
    };
    uint32_t base = 0;
    for (uint32_t b = b0; b < b1; ++b) {
      const BlockDesc d = p.blocks[b];
      for (int q = 0; q < W; ++q) dst[q] = p.recv1[q] + (size_t)r * p.slot_bytes + d.wire_off;
      const uint32_t n = block_n(d);
      const T* blk = data + d.elem_off;
      const bool aligned = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
      const uint32_t first = (warp - base) & (kNumWarps - 1);
      if (block_is_raw(d)) {
        const uint32_t ni = div_up(n, kRawItemElems);
        for (uint32_t i = first; i < ni; i += kNumWarps) {
          constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
          float f[U][V];
          raw_load_own<T>(blk, aligned, n, i * kRawItemElems, p.prescale, f);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t i0 = i * kRawItemElems + ((uint32_t)u * 32u + wl) * V;
            if (i0 >= n) continue;
            const uint4 packed = pack16<T>(f[u]);
            for (int q = 0; q < W; ++q) st_v4(dst[q] + (size_t)i0 * sizeof(T), packed);
          }
          item_done();
        }
        base += ni;
      } else if (block_is_fast(d)) {
        const uint32_t nb = block_num_buckets(n, d.bucket);
        for (uint32_t bk = first; bk < nb; bk += kNumWarps) {
          if (bucket_is_full(d, bk, aligned))
            bucket_send<T, true, KB>(blk, aligned, d, bk, p.prescale, rng, b, dst, W);
          else
            bucket_send<T, false, KB>(blk, aligned, d, bk, p.prescale, rng, b, dst, W);
          item_done();
        }
        base += nb;
      } else {
        const int bits = block_bits(d);
        load_block<T>(data, d, p.prescale, tile.acc);
        __syncthreads();
        compute_meta(tile.acc, n, d.bucket, bits, tile.meta, tile.inv);
        __syncthreads();
        pack_block<T, false>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, nullptr);
        __syncthreads();
        store_record(tile.meta, tile.pay, block_meta_bytes(n, d.bucket), block_payload_bytes(n, bits), dst, W);
        __syncthreads();
        if (warp == 0) item_done();
      }
    }
  }

  // ---- phase 2: wait for my own CTA's image and for the W-1 peers, then sum all W images
  {
    bool ok = true;
    if (wl < (uint32_t)W && (int)wl != r) {
      ok = wait_flag(p.flags1[r] + (size_t)wl * p.flag_stride + lane, p.epoch, p.timeout_ns);
      if (!ok) {
        s_abort = 1;
        *p.status = kSraTimeoutPhase1 | (wl << 8) | ((uint32_t)lane << 16);
      }
    }
    if (wl == 0) {  // other warps of this CTA wrote parts of my own slot
      uint32_t spins = 0;
      while (atom_add_acq_rel_cta(&s_done, 0u) < total)
        if (++spins > (1u << 28)) break;
    }
    if (!__all_sync(0xffffffffu, ok)) return;
    const uint8_t* src[kMaxPeers];
    uint32_t base = 0;
    for (uint32_t b = b0; b < b1; ++b) {
      const BlockDesc d = p.blocks[b];
      for (int q = 0; q < W; ++q) src[q] = p.recv1[r] + (size_t)q * p.slot_bytes + d.wire_off;
      const uint32_t n = block_n(d);
      T* blk = data + d.elem_off;
      const bool aligned = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
      const uint32_t first = (warp - base) & (kNumWarps - 1);
      if (block_is_raw(d)) {
        const uint32_t ni = div_up(n, kRawItemElems);
        for (uint32_t i = first; i < ni; i += kNumWarps) warp_sum_raw<T>(src, W, blk, aligned, n, i);
        base += ni;
      } else if (block_is_fast(d)) {
        const uint32_t nb = block_num_buckets(n, d.bucket);
        for (uint32_t bk = first; bk < nb; bk += kNumWarps) {
          if (bucket_is_full(d, bk, aligned))
            bucket_sum_recv<T, true, KB>(src, W, blk, aligned, d, bk);
          else
            bucket_sum_recv<T, false, KB>(src, W, blk, aligned, d, bk);
        }
        base += nb;
      } else {
        // slow block: fp32 accumulators in shared memory, all W images in rank order
        for (uint32_t i = tid; i < n; i += blockDim.x) tile.acc[i] = 0.f;
        __syncthreads();
        for (int q = 0; q < W; ++q) decode_add<T>(src[q], d, tile.acc);
        __syncthreads();
        T* out = data + d.elem_off;
        for (uint32_t i = tid; i < n; i += blockDim.x) out[i] = DT<T>::from_float(tile.acc[i]);
        __syncthreads();
      }
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
#define CGX_SET_SMEM(KB_)                                                                                   \
  e = cudaFuncSetAttribute(sra_fused_warp_kernel<T, KB_, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           (int)sizeof(Tile));                                                                \
  if (e != cudaSuccess) return e;                                                                             \
  e = cudaFuncSetAttribute(sra_fused_warp_kernel<T, KB_, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           (int)sizeof(Tile));                                                                \
  if (e != cudaSuccess) return e;
    CGX_SET_SMEM(0) CGX_SET_SMEM(2) CGX_SET_SMEM(4) CGX_SET_SMEM(8)
#define CGX_SET_SMEM_OS(KB_)                                                                                        \
  e = cudaFuncSetAttribute(oneshot_kernel<T, KB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile)); \
  if (e != cudaSuccess) return e;
    CGX_SET_SMEM_OS(0) CGX_SET_SMEM_OS(2) CGX_SET_SMEM_OS(4) CGX_SET_SMEM_OS(8)
    configured[dev & 63] = true;
  }
  if (p.variant == 3) {
    switch (p.uniform_bits) {
      case 2: oneshot_kernel<T, 2><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p); break;
      case 4: oneshot_kernel<T, 4><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p); break;
      case 8: oneshot_kernel<T, 8><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p); break;
      default: oneshot_kernel<T, 0><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p); break;
    }
  } else if (p.variant == 1) {
    sra_fused_kernel<T><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p);
  } else {
#define CGX_LAUNCH_WARP(KB_, PB_, MB_) \
  sra_fused_warp_kernel<T, KB_, PB_, MB_><<<p.lanes, kSraThreads, sizeof(Tile), stream>>>(p)
#define CGX_LAUNCH_WARP_KB(PB_, MB_)                          \
  switch (p.uniform_bits) {                                   \
    case 2: CGX_LAUNCH_WARP(2, PB_, MB_); break;              \
    case 4: CGX_LAUNCH_WARP(4, PB_, MB_); break;              \
    case 8: CGX_LAUNCH_WARP(8, PB_, MB_); break;              \
    default: CGX_LAUNCH_WARP(0, PB_, MB_); break;             \
  }
    if (p.variant == 2) {
      CGX_LAUNCH_WARP_KB(1, 2)
    } else {
      CGX_LAUNCH_WARP_KB(2, 1)
    }
  }
  return cudaGetLastError();
}

template <typename T>
int max_resident_t() {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(sra_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile));
  cudaFuncSetAttribute(sra_fused_warp_kernel<T, 0, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile));
  int per_sm2 = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sra_fused_kernel<T>, kSraThreads, sizeof(Tile)) !=
          cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, sra_fused_warp_kernel<T, 0, 2, 1>, kSraThreads,
                                                    sizeof(Tile)) != cudaSuccess)
    return 0;
  return sms * (per_sm < per_sm2 ? per_sm : per_sm2);
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
