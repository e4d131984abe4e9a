// Item-level operations built from item_path.cuh. Each function is executed by ONE warp on ONE
// item. The hot versions work on a slice whose values are already in registers (the callers
// prefetch the next item's data while the current one is processed) and assume 16/32 B aligned
// gradients; everything rare -- unaligned slices, generic buckets, raw tails -- is deliberately
// out of line so that the hot loops stay a few hundred instructions long.
//
//   *_send     quantize my values of an item and store the packed words to a destination
//              (SRA phase A, one-shot phase 1, standalone quantize)
//   *_reduce   own values + the decoded copies of W-1 sources, requantize, publish, self-decode
//              (SRA phase B)
//   *_recv     decode one source -- or the sum of all sources -- into the gradient buffer
//              (SRA phase C, one-shot phase 2, standalone dequantize)
// Reference counterparts: Compressor::Compress / Decompress per layer slice
// (/root/reference/src/common/compressor.cc:98-179) and the SRA loop bodies
// (/root/reference/src/common/scatter_reduce_allgather.cc:116-199).
#pragma once
#include "item_path.cuh"

namespace cgx {
namespace dev {

template <int GPL>
struct SliceCfg {
  static constexpr uint32_t kElems = 256u * GPL;
  static constexpr uint32_t kLgAll = GPL == 4 ? 7u : 6u;  // log2(bucket/8) of a bucket spanning the slice
  static constexpr int kPeerBatch = 2;  // sources fetched before any is consumed
};

// one bucket == the whole slice (512 / 1024-element buckets): the common case
template <int GPL>
__device__ __forceinline__ bool slice_single_bucket(uint32_t lg) {
  return GPL == 4 || lg == SliceCfg<GPL>::kLgAll;
}

// ---- slice <-> registers -------------------------------------------------------------
// Lane l owns the GPL adjacent pack groups l*GPL .. l*GPL+GPL-1, i.e. GPL*8 consecutive elements
// (row k of the register tile = group l*GPL + k). Its packed words are then adjacent too.
template <int GPL>
__device__ __forceinline__ uint32_t lane_group(int k) {
  return lane_id() * GPL + (uint32_t)k;
}
// issue the (vector) loads of a slice; `src` must be group aligned
template <typename TS, int GPL>
__device__ __forceinline__ void slice_load_vec(const TS* __restrict__ src, float (&x)[GPL][8]) {
  const TS* p = src + lane_id() * (GPL * 8u);
#pragma unroll
  for (int k = 0; k < GPL; ++k) load8_vec<TS>(p + k * 8, x[k]);
}
template <typename TS, int GPL>
__device__ __forceinline__ void slice_load_scalar(const TS* __restrict__ src, float (&x)[GPL][8]) {
  const TS* p = src + lane_id() * (GPL * 8u);
#pragma unroll
  for (int k = 0; k < GPL; ++k) load8_scalar<TS>(p + k * 8, 8, x[k]);
}
// pull the slice towards L2 without occupying registers (phase B prefetches its next item this way:
// its registers are taken by the peers' words)
template <typename TS, int GPL>
__device__ __forceinline__ void slice_prefetch_l2(const TS* __restrict__ src) {
  const TS* p = src + lane_id() * (GPL * 8u);
#pragma unroll
  for (int k = 0; k < GPL; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + k * 8));
}
template <int GPL>
__device__ __forceinline__ void slice_scale(float (&x)[GPL][8], float prescale) {
  if (prescale != 1.0f) {
#pragma unroll
    for (int k = 0; k < GPL; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[k][j] = __fmul_rn(x[k][j], prescale);
  }
}
template <typename TO, int GPL, bool VEC>
__device__ __forceinline__ void slice_store(TO* __restrict__ dst, const float (&x)[GPL][8]) {
  TO* p = dst + lane_id() * (GPL * 8u);
#pragma unroll
  for (int k = 0; k < GPL; ++k) {
    if (VEC)
      store8_vec<TO>(p + k * 8, x[k]);
    else
      store8_scalar<TO>(p + k * 8, 8, x[k]);
  }
}

// per-bucket {unit, min} and 1/unit of every group row of the slice
template <int GPL>
__device__ __forceinline__ void slice_meta(const float (&x)[GPL][8], uint32_t lg, int bits, BucketMeta (&m)[GPL],
                                           float (&inv)[GPL]) {
  float mn[GPL], mx[GPL];
#pragma unroll
  for (int k = 0; k < GPL; ++k) {
    mn[k] = CGX_INF_POS;
    mx[k] = CGX_INF_NEG;
    minmax8(x[k], mn[k], mx[k]);
  }
  if (slice_single_bucket<GPL>(lg)) {
    float a = mn[0], b = mx[0];
#pragma unroll
    for (int k = 1; k < GPL; ++k) {
      a = nan_min(a, mn[k]);
      b = nan_max(b, mx[k]);
    }
    warp_minmax(a, b);
    const BucketMeta m0 = make_meta(a, b, bits);
    const float i0 = inv_unit(m0.unit);
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      m[k] = m0;
      inv[k] = i0;
    }
  } else if (lg == 0) {  // 8-element buckets: every row is its own bucket
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      m[k] = make_meta(mn[k], mx[k], bits);
      inv[k] = inv_unit(m[k].unit);
    }
  } else {
    // (512-slices only) a bucket of 2^lg groups = all rows of 2^(lg-1) adjacent lanes
    float a = mn[0], b = mx[0];
#pragma unroll
    for (int k = 1; k < GPL; ++k) {
      a = nan_min(a, mn[k]);
      b = nan_max(b, mx[k]);
    }
    subwarp_minmax(a, b, lg - 1u);
    const BucketMeta m0 = make_meta(a, b, bits);
    const float i0 = inv_unit(m0.unit);
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      m[k] = m0;
      inv[k] = i0;
    }
  }
}

template <int GPL, typename DST>
__device__ __forceinline__ void slice_store_meta(const BucketMeta (&m)[GPL], uint32_t lg, const DST& ds,
                                                 uint32_t meta_off) {
  if (slice_single_bucket<GPL>(lg)) {
    if (lane_id() == 0) dst_st_v2(ds, meta_off, __float_as_uint(m[0].unit), __float_as_uint(m[0].min));
  } else {
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      const uint32_t gi = lane_group<GPL>(k);
      if ((gi & ((1u << lg) - 1u)) == 0)
        dst_st_v2(ds, meta_off + (gi >> lg) * 8u, __float_as_uint(m[k].unit), __float_as_uint(m[k].min));
    }
  }
}

// quantize + pack + store the slice; SELF: also write the decoded values (what every receiver
// will decode from the same bytes) to `self_out` (VEC: it is group aligned).
// MODE 0: deterministic rounding, every bucket finite (no clamp needed)  1: deterministic  2: stochastic
template <typename TO, int KB, int GPL, bool SELF, bool VEC, int MODE, typename DST>
__device__ __forceinline__ void slice_encode_mode(const float (&x)[GPL][8], const BucketMeta (&m)[GPL],
                                                  const float (&inv)[GPL], int bits, const RngKey& rng,
                                                  uint32_t first_elem, const DST& ds, uint32_t pay_off,
                                                  TO* __restrict__ self_out) {
  const float maxlvl = (float)max_level(bits);
  uint32_t lo[GPL], hi[GPL];
#pragma unroll
  for (int k = 0; k < GPL; ++k) {
    const uint32_t gi = lane_group<GPL>(k);
    float u[8];
    if (MODE == 2) {
      float r[8];
      rounding8(rng, first_elem + gi * 8u, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = level_magic<true>(x[k][j], m[k].min, inv[k], r[j], maxlvl);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = level_magic<MODE == 1>(x[k][j], m[k].min, inv[k], 0.5f, maxlvl);
    }
    pack_magic<KB>(u, bits, lo[k], hi[k]);
    if (SELF) {
      float dec[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dec[j] = __fmaf_rn(m[k].unit, __fsub_rn(u[j], CGX_MAGIC), m[k].min);
      if (VEC)
        store8_vec<TO>(self_out + gi * 8u, dec);
      else
        store8_scalar<TO>(self_out + gi * 8u, 8, dec);
    }
  }
  store_words<KB, GPL>(ds, pay_off, bits, lo, hi);
}

template <typename TO, int KB, int GPL, bool SELF, bool VEC, typename DST>
__device__ __forceinline__ void slice_encode(const float (&x)[GPL][8], const BucketMeta (&m)[GPL],
                                             const float (&inv)[GPL], int bits, const RngKey& rng,
                                             uint32_t first_elem, const DST& ds, uint32_t pay_off,
                                             TO* __restrict__ self_out) {
  if (rng.enabled) {
    slice_encode_mode<TO, KB, GPL, SELF, VEC, 2>(x, m, inv, bits, rng, first_elem, ds, pay_off, self_out);
    return;
  }
  // every bucket of the slice has finite min/max (unit finite <=> max - min finite) -> no clamp
  bool finite = true;
#pragma unroll
  for (int k = 0; k < GPL; ++k) finite = finite && (fabsf(m[k].unit) < CGX_INF_POS);
  if (__all_sync(kAll, finite))
    slice_encode_mode<TO, KB, GPL, SELF, VEC, 0>(x, m, inv, bits, rng, first_elem, ds, pay_off, self_out);
  else
    slice_encode_mode<TO, KB, GPL, SELF, VEC, 1>(x, m, inv, bits, rng, first_elem, ds, pay_off, self_out);
}

// packed words + meta of one source for a slice, as loaded (not yet decoded)
template <int GPL>
struct SliceWords {
  uint32_t lo[GPL], hi[GPL];
  BucketMeta pm[GPL];
};

// issue the loads of one source's packed words + meta for the slice (no use yet)
template <int KB, int GPL>
__device__ __forceinline__ void slice_fetch(const uint8_t* rec, uint32_t meta_off, uint32_t pay_off, uint32_t lg,
                                            int bits, SliceWords<GPL>& w) {
  load_words<KB, GPL>(rec + pay_off, bits, w.lo, w.hi);
  if (slice_single_bucket<GPL>(lg)) {
    const uint2 v = ld_sys_v2(rec + meta_off);
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      w.pm[k].unit = __uint_as_float(v.x);
      w.pm[k].min = __uint_as_float(v.y);
    }
  } else {
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      const uint2 v = ld_sys_v2(rec + meta_off + ((lane_group<GPL>(k) >> lg) * 8u));
      w.pm[k].unit = __uint_as_float(v.x);
      w.pm[k].min = __uint_as_float(v.y);
    }
  }
}

// x (+)= decode(words)
template <int KB, int GPL, bool ADD>
__device__ __forceinline__ void slice_decode(const SliceWords<GPL>& w, int bits, float (&x)[GPL][8]) {
#pragma unroll
  for (int k = 0; k < GPL; ++k) {
    float qf[8];
    unpack_magic<KB>(w.lo[k], w.hi[k], bits, qf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = __fmaf_rn(w.pm[k].unit, qf[j], w.pm[k].min);
      x[k][j] = ADD ? __fadd_rn(x[k][j], d) : d;
    }
  }
}

// x += decode(source i) for every source of the set, in slot order.
// Common case (one bucket == the slice): ONE {unit, min} per source, so the words of up to 8
// sources are in flight at once -- at W = 8 the whole reduction of an item is a single round trip
// to L2/HBM instead of four dependent ones. Slices made of several small buckets keep a
// per-row meta and go two sources at a time.
template <int KB, int GPL>
__device__ __forceinline__ void slice_accumulate(const SrcSet& ss, uint32_t meta_off, uint32_t pay_off, uint32_t lg,
                                                 int bits, float (&x)[GPL][8]) {
  const int cnt = ss.n - (ss.skip >= 0 ? 1 : 0);
  if (slice_single_bucket<GPL>(lg)) {
    constexpr int kB = GPL == 4 ? 2 : 4;
    for (int i0 = 0; i0 < cnt; i0 += kB) {
      uint32_t lo[kB][GPL], hi[kB][GPL];
      float un[kB], mi[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        // past the end: re-read the last source (unconditional loads keep the arrays in
        // registers; the value is simply not added)
        const int i = min(i0 + u, cnt - 1);
        const int q = (ss.skip >= 0 && i >= ss.skip) ? i + 1 : i;
        const uint8_t* rec = ss.base + (size_t)q * ss.stride;
        load_words<KB, GPL>(rec + pay_off, bits, lo[u], hi[u]);
        const uint2 v = ld_sys_v2(rec + meta_off);
        un[u] = __uint_as_float(v.x);
        mi[u] = __uint_as_float(v.y);
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (i0 + u < cnt) {
#pragma unroll
          for (int k = 0; k < GPL; ++k) {
            float qf[8];
            unpack_magic<KB>(lo[u][k], hi[u][k], bits, qf);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[k][j] = __fadd_rn(x[k][j], __fmaf_rn(un[u], qf[j], mi[u]));
          }
        }
      }
    }
    return;
  }
  constexpr int kB = SliceCfg<GPL>::kPeerBatch;
  for (int i0 = 0; i0 < cnt; i0 += kB) {
    SliceWords<GPL> w[kB];
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      const int i = i0 + u;
      if (i < cnt) {
        const int q = (ss.skip >= 0 && i >= ss.skip) ? i + 1 : i;
        slice_fetch<KB, GPL>(ss.base + (size_t)q * ss.stride, meta_off, pay_off, lg, bits, w[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kB; ++u)
      if (i0 + u < cnt) slice_decode<KB, GPL, true>(w[u], bits, x);
  }
}

// ======================================================================================
// full items, values already in registers (x = raw loaded values, not yet prescaled)
// ======================================================================================
// TO: element type of the self-decoded output (SELF only)
template <typename TO, int KB, int GPL, bool SELF, bool VEC, typename DST>
__device__ __forceinline__ void full_send_x(float (&x)[GPL][8], const WarpItem& it, float prescale, const RngKey& rng,
                                            const DST& ds, TO* __restrict__ self_out) {
  const int bits = KB ? KB : item_bits(it);
  const uint32_t lg = item_lpb_log2(it);
  slice_scale<GPL>(x, prescale);
  BucketMeta m[GPL];
  float inv[GPL];
  slice_meta<GPL>(x, lg, bits, m, inv);
  slice_store_meta<GPL>(m, lg, ds, it.meta_off);
  slice_encode<TO, KB, GPL, SELF, VEC>(x, m, inv, bits, rng, it.elem_off, ds, it.pay_off, self_out);
}

template <typename T, int KB, int GPL, bool VEC, typename DST>
__device__ __forceinline__ void full_reduce_x(float (&x)[GPL][8], T* __restrict__ blk, const WarpItem& it,
                                              float prescale, const RngKey& rng, const SrcSet& ss, const DST& ds) {
  const int bits = KB ? KB : item_bits(it);
  const uint32_t lg = item_lpb_log2(it);
  slice_scale<GPL>(x, prescale);
  slice_accumulate<KB, GPL>(ss, it.meta_off, it.pay_off, lg, bits, x);
  BucketMeta m[GPL];
  float inv[GPL];
  slice_meta<GPL>(x, lg, bits, m, inv);
  slice_store_meta<GPL>(m, lg, ds, it.meta_off);
  slice_encode<T, KB, GPL, true, VEC>(x, m, inv, bits, rng, it.elem_off, ds, it.pay_off, blk);
}

// out = decode(already fetched words of ONE source)
template <typename TO, int KB, int GPL, bool VEC>
__device__ __forceinline__ void full_recv_w(const SliceWords<GPL>& w, const WarpItem& it, TO* __restrict__ out) {
  const int bits = KB ? KB : item_bits(it);
  float x[GPL][8];
  slice_decode<KB, GPL, false>(w, bits, x);
  slice_store<TO, GPL, VEC>(out, x);
}

// out = sum of all sources, slot order (one-shot phase 2)
template <typename TO, int KB, int GPL, bool VEC>
__device__ __forceinline__ void full_recv_sum(const SrcSet& ss, const WarpItem& it, TO* __restrict__ out) {
  const int bits = KB ? KB : item_bits(it);
  const uint32_t lg = item_lpb_log2(it);
  float x[GPL][8];
#pragma unroll
  for (int k = 0; k < GPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[k][j] = 0.f;
  slice_accumulate<KB, GPL>(ss, it.meta_off, it.pay_off, lg, bits, x);
  slice_store<TO, GPL, VEC>(out, x);
}

// ---- two items at once ------------------------------------------------------------------
// With 16 warps per SM the dependent chains of one item (min/max tree -> redux -> reciprocal ->
// encode -> pack) leave the schedulers idle; two independent items in ONE basic block give the
// compiler twice the instruction-level parallelism. Only the common case qualifies: both slices
// are a single bucket, deterministic rounding, finite values. Returns false (nothing done) if not.
template <int GPL>
__device__ __forceinline__ void slice_meta_single(const float (&x)[GPL][8], int bits, BucketMeta& m, float& inv) {
  float mn = CGX_INF_POS, mx = CGX_INF_NEG;
#pragma unroll
  for (int k = 0; k < GPL; ++k) minmax8(x[k], mn, mx);
  warp_minmax(mn, mx);
  m = make_meta(mn, mx, bits);
  inv = inv_unit(m.unit);
}

template <int KB, int GPL, typename DST>
__device__ __forceinline__ bool full_send_pair(float (&xa)[GPL][8], float (&xb)[GPL][8], const WarpItem& ita,
                                               const WarpItem& itb, float prescale, const RngKey& rng,
                                               const DST& dsa, const DST& dsb) {
  if (rng.enabled || !slice_single_bucket<GPL>(item_lpb_log2(ita)) || !slice_single_bucket<GPL>(item_lpb_log2(itb)))
    return false;
  const int bits_a = KB ? KB : item_bits(ita), bits_b = KB ? KB : item_bits(itb);
  slice_scale<GPL>(xa, prescale);
  slice_scale<GPL>(xb, prescale);
  BucketMeta ma, mb;
  float ia, ib;
  slice_meta_single<GPL>(xa, bits_a, ma, ia);
  slice_meta_single<GPL>(xb, bits_b, mb, ib);
  if (lane_id() == 0) {
    dst_st_v2(dsa, ita.meta_off, __float_as_uint(ma.unit), __float_as_uint(ma.min));
    dst_st_v2(dsb, itb.meta_off, __float_as_uint(mb.unit), __float_as_uint(mb.min));
  }
  const bool finite = __all_sync(kAll, fabsf(ma.unit) < CGX_INF_POS && fabsf(mb.unit) < CGX_INF_POS);
  BucketMeta mav[GPL], mbv[GPL];
  float iav[GPL], ibv[GPL];
#pragma unroll
  for (int k = 0; k < GPL; ++k) {
    mav[k] = ma;
    mbv[k] = mb;
    iav[k] = ia;
    ibv[k] = ib;
  }
  if (finite) {
    // interleave by rows so that both items' chains are in flight together
    const float maxa = (float)max_level(bits_a), maxb = (float)max_level(bits_b);
    uint32_t loa[GPL], hia[GPL], lob[GPL], hib[GPL];
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
      float ua[8], ub[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ua[j] = level_magic<false>(xa[k][j], ma.min, ia, 0.5f, maxa);
        ub[j] = level_magic<false>(xb[k][j], mb.min, ib, 0.5f, maxb);
      }
      pack_magic<KB>(ua, bits_a, loa[k], hia[k]);
      pack_magic<KB>(ub, bits_b, lob[k], hib[k]);
    }
    store_words<KB, GPL>(dsa, ita.pay_off, bits_a, loa, hia);
    store_words<KB, GPL>(dsb, itb.pay_off, bits_b, lob, hib);
  } else {
    slice_encode_mode<float, KB, GPL, false, true, 1>(xa, mav, iav, bits_a, rng, ita.elem_off, dsa, ita.pay_off,
                                                      (float*)nullptr);
    slice_encode_mode<float, KB, GPL, false, true, 1>(xb, mbv, ibv, bits_b, rng, itb.elem_off, dsb, itb.pay_off,
                                                      (float*)nullptr);
  }
  return true;
}

// decode two already fetched items (branch free: the two chains interleave)
template <typename TO, int KB, int GPL>
__device__ __forceinline__ void full_recv_pair(const SliceWords<GPL>& wa, const SliceWords<GPL>& wb,
                                               const WarpItem& ita, const WarpItem& itb, TO* __restrict__ oa,
                                               TO* __restrict__ ob) {
  const int bits_a = KB ? KB : item_bits(ita), bits_b = KB ? KB : item_bits(itb);
  float xa[GPL][8], xb[GPL][8];
  slice_decode<KB, GPL, false>(wa, bits_a, xa);
  slice_decode<KB, GPL, false>(wb, bits_b, xb);
  slice_store<TO, GPL, true>(oa, xa);
  slice_store<TO, GPL, true>(ob, xb);
}

// ---- the same for slices whose gradients are not vector aligned: out of line, scalar access ----
template <typename TS, typename TO, int KB, int GPL, bool SELF, typename DST>
__device__ __noinline__ void full_send_unaligned(const TS* __restrict__ src, const WarpItem it, float prescale,
                                                 const RngKey rng, const DST ds, TO* __restrict__ self_out) {
  float x[GPL][8];
  slice_load_scalar<TS, GPL>(src, x);
  full_send_x<TO, KB, GPL, SELF, false>(x, it, prescale, rng, ds, self_out);
}
template <typename T, int KB, int GPL, typename DST>
__device__ __noinline__ void full_reduce_unaligned(T* __restrict__ blk, const WarpItem it, float prescale,
                                                   const RngKey rng, const SrcSet ss, const DST ds) {
  float x[GPL][8];
  slice_load_scalar<T, GPL>(blk, x);
  full_reduce_x<T, KB, GPL, false>(x, blk, it, prescale, rng, ss, ds);
}
template <typename TO, int KB, int GPL>
__device__ __noinline__ void full_recv_unaligned(const SrcSet ss, const WarpItem it, TO* __restrict__ out) {
  if (ss.n == 1) {
    SliceWords<GPL> w;
    slice_fetch<KB, GPL>(ss.base, it.meta_off, it.pay_off, item_lpb_log2(it), KB ? KB : item_bits(it), w);
    full_recv_w<TO, KB, GPL, false>(w, it, out);
  } else {
    full_recv_sum<TO, KB, GPL, false>(ss, it, out);
  }
}

// ======================================================================================
// generic bucket items (cold): ONE bucket of n <= kMaxBlockElems elements, any alignment,
// any bit width. Two passes (min/max, then encode); lane l handles groups l, l+32, ...
// ======================================================================================
template <typename TS>
__device__ __forceinline__ void bucket_load_group(const TS* __restrict__ src, uint32_t g, uint32_t n, float prescale,
                                                  float (&x)[8], int& nv) {
  nv = (int)min(8u, n - g * 8u);
  load8_scalar<TS>(src + g * 8u, nv, x);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = __fmul_rn(x[j], prescale);
}

__device__ __forceinline__ void bucket_add_sources(const SrcSet& ss, const WarpItem& it, uint32_t g, int bits,
                                                   float (&x)[8]) {
  for (int q = 0; q < ss.n; ++q) {
    if (q == ss.skip) continue;
    const uint8_t* rec = ss.base + (size_t)q * ss.stride;
    uint32_t lo, hi;
    load_word<0>(rec + it.pay_off, g, bits, lo, hi);
    const uint2 mv = ld_sys_v2(rec + it.meta_off);
    float qf[8];
    unpack_magic<0>(lo, hi, bits, qf);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      x[j] = __fadd_rn(x[j], __fmaf_rn(__uint_as_float(mv.x), qf[j], __uint_as_float(mv.y)));
  }
}

// phase A / phase B / quantize of a generic bucket. `ss.n == 0`: nothing to add (send);
// self_out != nullptr: write the self-decoded values there
template <typename TS, typename TO, typename DST>
__device__ __noinline__ void bucket_quantize(const TS* __restrict__ src, const WarpItem it, float prescale,
                                             const RngKey rng, const SrcSet ss, const DST ds,
                                             TO* __restrict__ self_out) {
  const uint32_t n = item_n(it);
  const int bits = item_bits(it);
  const uint32_t ng = div_up(n, 8u);
  float mn = CGX_INF_POS, mx = CGX_INF_NEG;
  for (uint32_t g = lane_id(); g < ng; g += 32) {
    float x[8];
    int nv;
    bucket_load_group<TS>(src, g, n, prescale, x, nv);
    if (ss.n > 0) bucket_add_sources(ss, it, g, bits, x);
    minmax8_pred(x, nv, mn, mx);
  }
  warp_minmax(mn, mx);
  const BucketMeta m = make_meta(mn, mx, bits);
  const float inv = inv_unit(m.unit);
  const float maxlvl = (float)max_level(bits);
  // generic items always use the unicast mappings (sub-word stores have no multimem form)
  const DST du = unicast_of(ds);
  if (lane_id() == 0) dst_st_v2(du, it.meta_off, __float_as_uint(m.unit), __float_as_uint(m.min));
  for (uint32_t g = lane_id(); g < ng; g += 32) {
    float x[8];
    int nv;
    bucket_load_group<TS>(src, g, n, prescale, x, nv);
    if (ss.n > 0) bucket_add_sources(ss, it, g, bits, x);
    float r[8];
    if (rng.enabled) {
      rounding8(rng, it.elem_off + g * 8u, r);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.5f;
    }
    float u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = (j < nv) ? level_magic<true>(x[j], m.min, inv, r[j], maxlvl) : CGX_MAGIC;
    uint32_t lo, hi;
    pack_magic<0>(u, bits, lo, hi);
    // byte-granular store (a bucket's payload may start at any byte offset)
    {
      const uint64_t w = (uint64_t)lo | ((uint64_t)hi << 32);
      const uint32_t o = it.pay_off + g * (uint32_t)bits;
      for (int t = 0; t < bits; ++t) dst_st_u8(du, o + t, (uint32_t)(w >> (8 * t)) & 0xFFu);
    }
    if (self_out != nullptr) {
      float dec[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dec[j] = __fmaf_rn(m.unit, __fsub_rn(u[j], CGX_MAGIC), m.min);
      store8_scalar<TO>(self_out + g * 8u, nv, dec);
    }
  }
}

// decode one source (ss.n == 1) or the sum of all sources into `out`
template <typename TO>
__device__ __noinline__ void bucket_recv(const SrcSet ss, const WarpItem it, TO* __restrict__ out) {
  const uint32_t n = item_n(it);
  const int bits = item_bits(it);
  const uint32_t ng = div_up(n, 8u);
  SrcSet s2 = ss;
  s2.skip = -1;
  for (uint32_t g = lane_id(); g < ng; g += 32) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    bucket_add_sources(s2, it, g, bits, x);  // 0 + d == d exactly
    store8_scalar<TO>(out + g * 8u, (int)min(8u, n - g * 8u), x);
  }
}

// ======================================================================================
// raw (uncompressed) items: n <= 512 elements travelling as T
// ======================================================================================
// 8 values -> 8 T on the wire (16 B-aligned records): 2 x 16 B for fp32, 1 x 16 B otherwise
template <typename T, typename DST>
__device__ __forceinline__ void raw_wire_store(const DST& ds, uint32_t off, const float (&v)[8]) {
  if (sizeof(T) == 4) {
    dst_st_v4(ds, off, pack16<T>(v));
    dst_st_v4(ds, off + 16u, pack16<T>(v + 4));
  } else {
    dst_st_v4(ds, off, pack16<T>(v));
  }
}
template <typename T>
__device__ __forceinline__ void raw_wire_load(const uint8_t* p, float (&v)[8]) {
  if (sizeof(T) == 4) {
    unpack16<T>(ld_sys_v4(p), v);
    unpack16<T>(ld_sys_v4(p + 16), v + 4);
  } else {
    unpack16<T>(ld_sys_v4(p), v);
  }
}

// Full raw item (512 elements, vector-aligned gradients); x = the item's raw loaded values (MODE 0/1)
//   MODE 0: push (phase A / one-shot 1): dst <- T(x * prescale)
//   MODE 1: reduce (phase B): x * prescale + sum(sources) -> own, dst
//   MODE 2: receive: out <- source (ss.n == 1) or sum of all sources
// (x may have more than 2 rows -- the prefetch buffer of a 1024 slice -- only rows 0 and 1 are used)
template <typename T, int MODE, int ROWS, typename DST>
__device__ __forceinline__ void raw_full_x(float (&x)[ROWS][8], T* __restrict__ blk, const WarpItem& it,
                                           float prescale, const SrcSet& ss, const DST& ds) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[k][j] = MODE == 2 ? 0.f : __fmul_rn(x[k][j], prescale);
  if (MODE != 0) {
    for (int q = 0; q < ss.n; ++q) {
      if (q == ss.skip) continue;
      const uint8_t* rec = ss.base + (size_t)q * ss.stride + it.meta_off;
      float v[2][8];
#pragma unroll
      for (int k = 0; k < 2; ++k) raw_wire_load<T>(rec + (lane_id() * 2u + (uint32_t)k) * 8u * sizeof(T), v[k]);
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[k][j] = __fadd_rn(x[k][j], v[k][j]);
    }
  }
  if (MODE != 2) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
      raw_wire_store<T>(ds, it.meta_off + (lane_id() * 2u + (uint32_t)k) * 8u * (uint32_t)sizeof(T), x[k]);
  }
  if (MODE != 0) {
    T* o = blk + lane_id() * 16u;
#pragma unroll
    for (int k = 0; k < 2; ++k) store8_vec<T>(o + k * 8, x[k]);
  }
}
// rows 0..1 of a (possibly larger) register buffer <- the 512 elements of a raw item
template <typename T, int ROWS>
__device__ __forceinline__ void raw_load_vec(const T* __restrict__ src, float (&x)[ROWS][8]) {
  const T* p = src + lane_id() * 16u;
#pragma unroll
  for (int k = 0; k < 2; ++k) load8_vec<T>(p + k * 8, x[k]);
}

// In-switch reduction of a full raw item (NVLS): every rank staged T(src * prescale) at the same
// heap offset; the owner pulls the sum with multimem.ld_reduce and multicasts it back in place.
template <typename T>
__device__ __forceinline__ void raw_full_mc_reduce(T* __restrict__ blk, const WarpItem& it, uint8_t* mc_slot) {
  constexpr uint32_t kVecs = 512u * sizeof(T) / 16u / 32u;  // 16 B vectors per lane: 4 (fp32) or 2
  uint4 v[kVecs];
#pragma unroll
  for (uint32_t i = 0; i < kVecs; ++i) v[i] = mc_ld_reduce_v4<T>(mc_slot + it.meta_off + (i * 32u + lane_id()) * 16u);
  const bool vec = (reinterpret_cast<uintptr_t>(blk) & 15u) == 0;
#pragma unroll
  for (uint32_t i = 0; i < kVecs; ++i) {
    const uint32_t byte = (i * 32u + lane_id()) * 16u;
    mc_st_v4(mc_slot + it.meta_off + byte, v[i]);
    T* o = blk + byte / sizeof(T);
    if (vec) {
      *reinterpret_cast<uint4*>(o) = v[i];
    } else {
      const T* pe = reinterpret_cast<const T*>(&v[i]);
#pragma unroll
      for (uint32_t j = 0; j < 16u / sizeof(T); ++j) o[j] = pe[j];
    }
  }
}

// raw tail (< 512 elements) or a full raw item with unaligned gradients: element-wise, cold.
// mode as in raw_full_x.
template <typename T, typename DST>
__device__ __noinline__ void raw_generic(T* __restrict__ blk, const WarpItem it, float prescale, const SrcSet ss,
                                         const DST ds, int mode) {
  const uint32_t n = item_n(it);
  const DST du = unicast_of(ds);
  for (uint32_t i = lane_id(); i < n; i += 32) {
    float x = mode == 2 ? 0.f : __fmul_rn(DT<T>::to_float(blk[i]), prescale);
    if (mode != 0) {
      for (int q = 0; q < ss.n; ++q) {
        if (q == ss.skip) continue;
        const T* rec = reinterpret_cast<const T*>(ss.base + (size_t)q * ss.stride + it.meta_off);
        T v;
        if (sizeof(T) == 4) {
          const uint32_t b = ld_sys_u32(rec + i);
          v = *reinterpret_cast<const T*>(&b);
        } else {
          const uint16_t b = (uint16_t)ld_sys_u16(rec + i);
          v = *reinterpret_cast<const T*>(&b);
        }
        x = __fadd_rn(x, DT<T>::to_float(v));
      }
    }
    const T t = DT<T>::from_float(x);
    if (mode != 2) {
      const uint32_t o = it.meta_off + i * (uint32_t)sizeof(T);
      if (sizeof(T) == 4)
        dst_st_u32(du, o, *reinterpret_cast<const uint32_t*>(&t));
      else
        dst_st_u16(du, o, *reinterpret_cast<const uint16_t*>(&t));
    }
    if (mode != 0) blk[i] = t;
  }
}

}  // namespace dev
}  // namespace cgx
