// Warp-centric fast path of the fused kernel: ONE WARP owns ONE quantization
// bucket end to end, entirely in registers -- 128-bit loads of the gradients,
// shuffle min/max, quantize, pack, and stores straight to the peers' slots.
// No shared memory, no block-wide barriers: warps of a CTA run fully
// decoupled and the memory system always sees 16-32 independent warps per SM
// each with 4-8 outstanding 16 B loads.
//
// A warp streams its bucket in SLICES of 512 elements (2 pack groups of 8 per
// lane). Buckets of <= 512 elements are single-slice: one pass, the values
// never leave the registers. Larger buckets take two passes (min/max, then
// quantize) with the second pass re-reading the slice from L1/L2.
// Eligible blocks: bits 1..8 and bucket_size a multiple of 256 (the reference
// defaults 512 / 1024 both qualify). Everything else takes the CTA-cooperative
// shared-memory path of block_device.cuh.
#pragma once
#include "block_device.cuh"

namespace cgx {
namespace dev {

constexpr int kMaxGpl = 2;             // pack groups (of 8 values) per lane and slice
constexpr uint32_t kSliceElems = 256u * kMaxGpl;
constexpr uint32_t kRawItemElems = 512;  // raw (uncompressed) blocks are cut into warp items of this size

__device__ __forceinline__ bool block_is_fast(const BlockDesc& d) {
  return block_bits(d) <= 8 && (d.bucket & 255u) == 0;
}

// number of warp work items in a block
__device__ __forceinline__ uint32_t block_items(const BlockDesc& d) {
  if (block_is_raw(d)) return div_up(block_n(d), kRawItemElems);
  if (block_is_fast(d)) return block_num_buckets(block_n(d), d.bucket);
  return 1u;  // slow block: processed by the whole CTA as one item
}

struct BucketCtx {  // one SLICE of a bucket
  uint32_t e0;       // first element of the slice, relative to the block start
  uint32_t cnt;      // elements in the slice (<= kSliceElems)
  uint32_t grp0;     // index of the slice's first pack group inside the block
  int nv[kMaxGpl];   // valid elements of this lane's k-th group (0..8)
};

// number of slices of bucket `bk` and its element count
__device__ __forceinline__ uint32_t bucket_count(const BlockDesc& d, uint32_t bk) {
  return min(d.bucket, block_n(d) - bk * d.bucket);
}

__device__ __forceinline__ BucketCtx make_slice_ctx(const BlockDesc& d, uint32_t bk, uint32_t sl) {
  BucketCtx c;
  const uint32_t bcnt = bucket_count(d, bk);
  c.e0 = bk * d.bucket + sl * kSliceElems;
  c.cnt = min(kSliceElems, bcnt - sl * kSliceElems);
  c.grp0 = c.e0 >> 3;
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
    const uint32_t off = ((uint32_t)k * 32u + lane) * 8u;
    c.nv[k] = (c.cnt > off) ? (int)min(8u, c.cnt - off) : 0;
  }
  return c;
}

// x[k][j] = float(src[e0 + (k*32+lane)*8 + j]) * prescale
// FULL: the slice has all 512 elements and is 16 B aligned -> no per-element predicates.
template <typename T, bool FULL>
__device__ __forceinline__ void warp_load_bucket(const T* __restrict__ blk, bool aligned, const BucketCtx& c,
                                                 float prescale, float (&x)[kMaxGpl][8]) {
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
    const T* p = blk + c.e0 + ((uint32_t)k * 32u + lane) * 8u;
    if (FULL || (c.nv[k] == 8 && aligned)) {
      if (sizeof(T) == 4) {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        const uint4 b = *(reinterpret_cast<const uint4*>(p) + 1);
        unpack16<T>(a, x[k]);
        unpack16<T>(b, x[k] + 4);
      } else {
        unpack16<T>(*reinterpret_cast<const uint4*>(p), x[k]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[k][j] = (j < c.nv[k]) ? DT<T>::to_float(p[j]) : 0.f;
    }
    if (prescale != 1.0f) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[k][j] *= prescale;
    }
  }
}

// Issue the loads of one peer's packed words + meta for this slice (no use yet).
// WORD = uint32_t when a group of 8 levels fits in 32 bits (bits <= 4), else uint64_t
template <bool FULL, typename WORD>
__device__ __forceinline__ void warp_fetch_peer(const uint8_t* rec, uint32_t meta_bytes, uint32_t bk, int bits,
                                                const BucketCtx& c, WORD (&w)[kMaxGpl], BucketMeta& m) {
  const uint32_t lane = threadIdx.x & 31u;
  m = load_meta(rec, bk);
  const uint8_t* pay = rec + meta_bytes;
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k)
    w[k] = (FULL || c.nv[k] > 0) ? (WORD)load_group_word(pay, c.grp0 + (uint32_t)k * 32u + lane, bits) : (WORD)0;
}

// level j of a packed word; 32-bit arithmetic when the whole group fits in 32 bits
__device__ __forceinline__ uint32_t bfe32(uint32_t v, int pos, int len) {
  uint32_t r;
  asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
  return r;
}
__device__ __forceinline__ uint32_t bfi32(uint32_t field, uint32_t base, int pos, int len) {
  uint32_t r;
  asm("bfi.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(field), "r"(base), "r"(pos), "r"(len));
  return r;
}
// level j of a packed word: one bit-field extract on the 32-bit half that holds it
__device__ __forceinline__ uint32_t unpack_level(uint64_t w, int j, int bits) {
  if (bits <= 4) return bfe32((uint32_t)w, j * bits, bits);
  if (bits == 8) return bfe32(j < 4 ? (uint32_t)w : (uint32_t)(w >> 32), (j & 3) * 8, 8);
  return unpack1(w, j, bits);
}

template <bool FULL, typename WORD>
__device__ __forceinline__ void warp_accumulate(const WORD (&w)[kMaxGpl], const BucketMeta& m, int bits,
                                                const BucketCtx& c, float (&x)[kMaxGpl][8]) {
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (FULL || j < c.nv[k]) x[k][j] += decode_level(unpack_level(w[k], j, bits), m.unit, m.min);
  }
}

template <bool FULL>
__device__ __forceinline__ void warp_minmax_update(const float (&x)[kMaxGpl][8], const BucketCtx& c, float& mn,
                                                   float& mx) {
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (FULL || j < c.nv[k]) {
        mn = nan_min(mn, x[k][j]);
        mx = nan_max(mx, x[k][j]);
      }
  }
}

// Warp-wide min/max with ONE redux.sync each instead of 5 shuffle rounds: floats are mapped to
// order-preserving signed integers (-0 < +0, like min.f32/max.f32); NaN is handled by a vote so
// that it still poisons the bucket exactly like the NaN-propagating scalar path.
__device__ __forceinline__ int float_to_ordered(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7FFFFFFF)); }

__device__ __forceinline__ BucketMeta warp_minmax_finish(float mn, float mx, int bits) {
  const bool has_nan = __any_sync(0xffffffffu, (mn != mn) || (mx != mx));
  const int kmn = __reduce_min_sync(0xffffffffu, float_to_ordered(mn));
  const int kmx = __reduce_max_sync(0xffffffffu, float_to_ordered(mx));
  mn = ordered_to_float(kmn);
  mx = ordered_to_float(kmx);
  if (has_nan) mn = mx = __int_as_float(0x7fffffff);
  return make_meta(mn, mx, bits);
}

__device__ __forceinline__ void store_group_word(uint8_t* pay, uint32_t g, int bits, uint64_t w) {
  uint8_t* dst = pay + (size_t)g * bits;
  switch (bits) {
    case 8: asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"((uint32_t)w), "r"((uint32_t)(w >> 32)) : "memory"); break;
    case 4: asm volatile("st.global.u32 [%0], %1;" ::"l"(dst), "r"((uint32_t)w) : "memory"); break;
    case 2: asm volatile("st.global.u16 [%0], %1;" ::"l"(dst), "h"((uint16_t)w) : "memory"); break;
    default:
      for (int t = 0; t < bits; ++t)
        asm volatile("st.global.u8 [%0], %1;" ::"l"(dst + t), "r"((uint32_t)(w >> (8 * t)) & 0xFFu) : "memory");
  }
}

__device__ __forceinline__ void warp_store_meta(const BucketMeta& m, uint32_t bk, uint8_t* const* dst_rec, int ndst) {
  if ((threadIdx.x & 31u) == 0) {
    for (int d = 0; d < ndst; ++d)
      asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(dst_rec[d] + (size_t)bk * 8u), "f"(m.unit), "f"(m.min)
                   : "memory");
  }
}

__device__ __forceinline__ uint64_t pack_levels(const uint32_t (&q)[8], int bits) {
  if (bits <= 4) {
    uint32_t w = q[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) w = bfi32(q[j], w, j * bits, bits);
    return w;
  }
  if (bits == 8) {
    uint32_t lo = q[0], hi = q[4];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      lo = bfi32(q[j], lo, j * 8, 8);
      hi = bfi32(q[4 + j], hi, j * 8, 8);
    }
    return (uint64_t)lo | ((uint64_t)hi << 32);
  }
  return pack8(q, bits);
}

template <typename T>
__device__ __forceinline__ void store_group_values(T* __restrict__ o, const float (&dec)[8], bool vec, int nv) {
  if (vec) {
    if (sizeof(T) == 4) {
      *reinterpret_cast<uint4*>(o) = pack16<T>(dec);
      *(reinterpret_cast<uint4*>(o) + 1) = pack16<T>(dec + 4);
    } else {
      *reinterpret_cast<uint4*>(o) = pack16<T>(dec);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) o[j] = DT<T>::from_float(dec[j]);
  }
}

// Quantize the slice held in x, write its packed words to `ndst` records
// and (SELF) the decoded values to the owner's gradient buffer.
template <typename T, bool SELF, bool FULL>
__device__ __forceinline__ void warp_quantize_store(const float (&x)[kMaxGpl][8], const BucketCtx& c,
                                                    const BucketMeta& m, int bits, uint32_t meta_bytes,
                                                    const RngKey& rng, uint32_t block_id, uint8_t* const* dst_rec,
                                                    int ndst, T* __restrict__ own_blk, bool aligned) {
  const uint32_t lane = threadIdx.x & 31u;
  const float iv = inv_unit(m.unit);
  const float maxlvl = (float)max_level(bits);
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
    if (!FULL && c.nv[k] == 0) continue;
    const uint32_t g = c.grp0 + (uint32_t)k * 32u + lane;
    uint32_t q[8];
    if (rng.enabled) {
      float r[8];
      rounding_offsets8(rng, block_id, g, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = (FULL || j < c.nv[k]) ? encode_level(x[k][j], m.min, iv, r[j], maxlvl) : 0u;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = (FULL || j < c.nv[k]) ? encode_level(x[k][j], m.min, iv, 0.5f, maxlvl) : 0u;
    }
    const uint64_t w = pack_levels(q, bits);
    for (int d = 0; d < ndst; ++d) store_group_word(dst_rec[d] + meta_bytes, g, bits, w);
    if (SELF) {
      float dec[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dec[j] = decode_level(q[j], m.unit, m.min);
      T* o = own_blk + c.e0 + ((uint32_t)k * 32u + lane) * 8u;
      store_group_values<T>(o, dec, FULL || (c.nv[k] == 8 && aligned), c.nv[k]);
    }
  }
}

// Decode one slice of a (peer-written) record into the gradient buffer.
template <typename T, bool FULL>
__device__ __forceinline__ void warp_decode_store(const uint8_t* rec, uint32_t meta_bytes, uint32_t bk, int bits,
                                                  const BucketCtx& c, T* __restrict__ own_blk, bool aligned) {
  const uint32_t lane = threadIdx.x & 31u;
  uint64_t w[kMaxGpl];
  BucketMeta m;
  warp_fetch_peer<FULL, uint64_t>(rec, meta_bytes, bk, bits, c, w, m);
#pragma unroll
  for (int k = 0; k < kMaxGpl; ++k) {
    if (!FULL && c.nv[k] == 0) continue;
    float dec[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dec[j] = decode_level(unpack_level(w[k], j, bits), m.unit, m.min);
    T* o = own_blk + c.e0 + ((uint32_t)k * 32u + lane) * 8u;
    store_group_values<T>(o, dec, FULL || (c.nv[k] == 8 && aligned), c.nv[k]);
  }
}

// ---- raw (uncompressed) warp items -----------------------------------------
// item `it` of a raw block covers elements [it*kRawItemElems, ...) of the block:
// kRawU 16-byte vectors per lane, all loads of a source issued before any use.
template <typename T>
struct RawCfg {
  static constexpr int V = DT<T>::kVec;
  static constexpr int U = (int)(kRawItemElems / (32u * V));  // vectors per lane and item
};

template <typename T>
__device__ __forceinline__ void raw_load_own(const T* __restrict__ blk, bool aligned, uint32_t n, uint32_t lo,
                                             float prescale, float (&f)[RawCfg<T>::U][RawCfg<T>::V]) {
  constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
    if (aligned && i0 + V <= n) {
      unpack16<T>(*reinterpret_cast<const uint4*>(blk + i0), f[u]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) f[u][k] = (i0 + k < n) ? DT<T>::to_float(blk[i0 + k]) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) f[u][k] *= prescale;
  }
}

template <typename T>
__device__ __forceinline__ void raw_store_own(T* __restrict__ blk, bool aligned, uint32_t n, uint32_t i0,
                                              const uint4& packed) {
  constexpr int V = RawCfg<T>::V;
  if (aligned && i0 + V <= n) {
    *reinterpret_cast<uint4*>(blk + i0) = packed;
  } else {
    const T* pe = reinterpret_cast<const T*>(&packed);
#pragma unroll
    for (int k = 0; k < V; ++k)
      if (i0 + k < n) blk[i0 + k] = pe[k];
  }
}

template <typename T>
__device__ __forceinline__ void warp_send_raw(const T* __restrict__ blk, bool aligned, uint32_t n, uint32_t it,
                                              float prescale, uint8_t* rec) {
  constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lo = it * kRawItemElems;
  float f[U][V];
  raw_load_own<T>(blk, aligned, n, lo, prescale, f);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
    if (i0 < n) st_v4(rec + (size_t)i0 * sizeof(T), pack16<T>(f[u]));
  }
}

template <typename T>
__device__ __forceinline__ void warp_reduce_raw(T* __restrict__ blk, bool aligned, uint32_t n, uint32_t it,
                                                float prescale, const uint8_t* const* peer_rec, int npeer,
                                                uint8_t* const* dst_rec, int ndst) {
  constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lo = it * kRawItemElems;
  float f[U][V];
  raw_load_own<T>(blk, aligned, n, lo, prescale, f);
  for (int q = 0; q < npeer; ++q) {
    uint4 pw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
      pw[u] = (i0 < n) ? ld_sys_v4(peer_rec[q] + (size_t)i0 * sizeof(T)) : make_uint4(0, 0, 0, 0);
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
    if (i0 >= n) continue;
    const uint4 packed = pack16<T>(f[u]);
    raw_store_own<T>(blk, aligned, n, i0, packed);
    for (int d = 0; d < ndst; ++d) st_v4(dst_rec[d] + (size_t)i0 * sizeof(T), packed);
  }
}

template <typename T>
__device__ __forceinline__ void warp_copy_raw(const uint8_t* rec, T* __restrict__ blk, bool aligned, uint32_t n,
                                              uint32_t it) {
  constexpr int V = RawCfg<T>::V, U = RawCfg<T>::U;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lo = it * kRawItemElems;
  uint4 pw[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
    pw[u] = (i0 < n) ? ld_sys_v4(rec + (size_t)i0 * sizeof(T)) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i0 = lo + ((uint32_t)u * 32u + lane) * V;
    if (i0 < n) raw_store_own<T>(blk, aligned, n, i0, pw[u]);
  }
}

// ---- warp-level completion counting -----------------------------------------
// Each finished item bumps a CTA-shared counter; whoever completes the chunk
// publishes it system-wide. acq_rel at CTA scope chains the other warps'
// stores before the final system-scope release.
__device__ __forceinline__ uint32_t atom_add_acq_rel_cta(uint32_t* smem_ctr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;"
               : "=r"(old)
               : "r"((uint32_t)__cvta_generic_to_shared(smem_ctr)), "r"(v)
               : "memory");
  return old;
}

}  // namespace dev
}  // namespace cgx
