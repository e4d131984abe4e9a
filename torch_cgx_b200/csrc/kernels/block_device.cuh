// CTA-cooperative block primitives (device side of common/block_ops.h).
// One CTA processes one block at a time out of a shared-memory tile:
//   load_block       gradients (T, global)  -> fp32 accumulators (smem), * prescale
//   compute_meta     per-bucket NaN-propagating min/max -> {unit,min} (smem)
//   pack_block       fp32 accumulators -> packed payload (smem) [+ self-decode to global]
//   store_record     meta + payload (smem) -> up to W wire slots (global / peer)
//   decode_add       wire record (global, peer-written) -> += fp32 accumulators
//   decode_store     wire record -> gradients (T, global)
// Replaces the reference's three kernels per layer slice (find_meta, pack_array,
// UnpackArray; /root/reference/src/common/compression/
// cuda_compression_operations.cu:98-153, :287-371, :474-544).
#pragma once
#include "../common/philox.h"
#include "../common/quant_math.h"
#include "../common/wire.h"
#include "device_utils.cuh"

namespace cgx {
namespace dev {

struct alignas(16) Tile {
  float acc[kMaxBlockElems];             // 32 KB
  uint8_t pay[kMaxBlockElems];           // 8 KB  (bits <= 8)
  BucketMeta meta[kMaxBlockBuckets + 2]; // 4 KB  (+pad entry so the section is 16 B padded)
  float inv[kMaxBlockBuckets];           // 2 KB
};

// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_block(const T* __restrict__ base, const BlockDesc& d, float prescale,
                                           float* acc) {
  constexpr int V = DT<T>::kVec;
  const uint32_t n = block_n(d);
  const T* src = base + d.elem_off;
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const uint32_t nvec = n / V;
    for (uint32_t i = tid; i < nvec; i += nt) {
      uint4 raw = ld_stream_v4(reinterpret_cast<const uint4*>(src) + i);
      float f[V];
      unpack16<T>(raw, f);
#pragma unroll
      for (int k = 0; k < V; k += 4) {
        float4 o = make_float4(f[k] * prescale, f[k + 1] * prescale, f[k + 2] * prescale, f[k + 3] * prescale);
        *reinterpret_cast<float4*>(acc + (size_t)i * V + k) = o;
      }
    }
    for (uint32_t i = nvec * V + tid; i < n; i += nt) acc[i] = DT<T>::to_float(src[i]) * prescale;
  } else {
    for (uint32_t i = tid; i < n; i += nt) acc[i] = DT<T>::to_float(src[i]) * prescale;
  }
}

// fp32 source (generic reducer accumulators), no prescale
__device__ __forceinline__ void load_block_f32(const float* __restrict__ src, uint32_t n, float* acc) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const uint32_t nvec = n / 4;
    for (uint32_t i = tid; i < nvec; i += nt)
      *reinterpret_cast<float4*>(acc + (size_t)i * 4) = *(reinterpret_cast<const float4*>(src) + i);
    for (uint32_t i = nvec * 4 + tid; i < n; i += nt) acc[i] = src[i];
  } else {
    for (uint32_t i = tid; i < n; i += nt) acc[i] = src[i];
  }
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ void compute_meta(const float* acc, uint32_t n, uint32_t B, int bits,
                                             BucketMeta* meta, float* inv) {
  const uint32_t nb = block_num_buckets(n, B);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarps = blockDim.x >> 5;
  for (uint32_t b = warp; b < nb; b += nwarps) {
    const uint32_t lo = b * B;
    const uint32_t hi = min(lo + B, n);
    float mn = acc[lo], mx = mn;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      float v = acc[i];
      mn = nan_min(mn, v);
      mx = nan_max(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      BucketMeta m = make_meta(mn, mx, bits);
      meta[b] = m;
      inv[b] = inv_unit(m.unit);
    }
  }
  // zero the padding entry so the 16 B-padded meta section is deterministic
  if (threadIdx.x == 0 && (nb & 1u)) {
    meta[nb].unit = 0.f;
    meta[nb].min = 0.f;
  }
}

// ---------------------------------------------------------------------------
// Quantize + pack the tile. If SELF_DECODE, also write T(decode(level)) to
// out[elem_off + i] (the owner's copy of the requantized sum, bit-identical to
// what every peer will decode from the same bytes).
template <typename T, bool SELF_DECODE>
__device__ __forceinline__ void pack_block(const float* acc, const BlockDesc& d, const BucketMeta* meta,
                                           const float* inv, uint8_t* pay, const RngKey& rng,
                                           uint32_t block_id, T* __restrict__ out_base) {
  const uint32_t n = block_n(d);
  const int bits = block_bits(d);
  const uint32_t B = d.bucket;
  const uint32_t groups = div_up(n, 8u);
  const float maxlvl = (float)max_level(bits);
  const bool uniform = (B & 7u) == 0;  // a group of 8 never straddles a bucket
  T* out = SELF_DECODE ? out_base + d.elem_off : nullptr;
  const bool out_vec = SELF_DECODE && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);

  for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
    const uint32_t i0 = g * 8u;
    float x[8];
    if (i0 + 8u <= n) {
      float4 a = *reinterpret_cast<const float4*>(acc + i0);
      float4 b = *reinterpret_cast<const float4*>(acc + i0 + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
      x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (i0 + j < n) ? acc[i0 + j] : 0.f;
    }
    float r[8];
    rounding_offsets8(rng, block_id, g, r);
    uint32_t q[8];
    float dec[8];
    if (uniform) {
      const uint32_t b = i0 / B;
      const BucketMeta m = meta[b];
      const float iv = inv[b];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        q[j] = (i0 + j < n) ? encode_level(x[j], m.min, iv, r[j], maxlvl) : 0u;
        if (SELF_DECODE) dec[j] = decode_level(q[j], m.unit, m.min);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t i = i0 + j;
        if (i < n) {
          const uint32_t b = i / B;
          const BucketMeta m = meta[b];
          q[j] = encode_level(x[j], m.min, inv[b], r[j], maxlvl);
          if (SELF_DECODE) dec[j] = decode_level(q[j], m.unit, m.min);
        } else {
          q[j] = 0u;
          if (SELF_DECODE) dec[j] = 0.f;
        }
      }
    }
    const uint64_t w = pack8(q, bits);
    uint8_t* dst = pay + (size_t)g * bits;
    switch (bits) {
      case 8: *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)w, (uint32_t)(w >> 32)); break;
      case 4: *reinterpret_cast<uint32_t*>(dst) = (uint32_t)w; break;
      case 2: *reinterpret_cast<uint16_t*>(dst) = (uint16_t)w; break;
      case 1: *dst = (uint8_t)w; break;
      default:
        for (int t = 0; t < bits; ++t) dst[t] = (uint8_t)(w >> (8 * t));
    }
    if (SELF_DECODE) {
      if (out_vec && i0 + 8u <= n) {
        if (sizeof(T) == 4) {
          st_v4(out + i0, pack16<T>(dec));
          st_v4(out + i0 + 4, pack16<T>(dec + 4));
        } else {
          st_v4(out + i0, pack16<T>(dec));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (i0 + j < n) out[i0 + j] = DT<T>::from_float(dec[j]);
      }
    }
  }
  // zero the (<16 B) padding after the last group
  const uint32_t used = groups * (uint32_t)bits;
  const uint32_t tot = block_payload_bytes(n, bits);
  if (threadIdx.x < tot - used) pay[used + threadIdx.x] = 0;
}

// Raw (uncompressed) block, phase A: dst[i] = T(float(src[i]) * prescale), no staging.
template <typename T>
__device__ __forceinline__ void send_raw(const T* __restrict__ base, const BlockDesc& d, float prescale,
                                         uint8_t* dst) {
  constexpr int V = DT<T>::kVec;
  const uint32_t n = block_n(d);
  const T* src = base + d.elem_off;
  const bool src_vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
  const uint32_t nvec = div_up(n, (uint32_t)V);
  for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[V];
    if (src_vec && (v + 1) * V <= n) {
      unpack16<T>(ld_stream_v4(reinterpret_cast<const uint4*>(src) + v), f);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] = ((size_t)v * V + k < n) ? DT<T>::to_float(src[(size_t)v * V + k]) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) f[k] *= prescale;
    st_v4(dst + ((size_t)v << 4), pack16<T>(f));
  }
}

// Raw block, phase B: sum = own*prescale + sum_q float(peer_rec[q][i]) (fixed
// order), written as T to the owner's gradient and to `ndst` wire slots. One
// pass, registers only -- the classic two-shot allreduce inner loop.
template <typename T>
__device__ __forceinline__ void reduce_raw(T* __restrict__ base, const BlockDesc& d, float prescale,
                                           const uint8_t* const* peer_rec, int npeer, uint8_t* const* dst,
                                           int ndst) {
  constexpr int V = DT<T>::kVec;
  const uint32_t n = block_n(d);
  T* own = base + d.elem_off;
  const bool own_vec = (reinterpret_cast<uintptr_t>(own) & 15u) == 0;
  const uint32_t nvec = div_up(n, (uint32_t)V);
  for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    const bool full = own_vec && (v + 1) * V <= n;
    float f[V];
    if (full) {
      unpack16<T>(ld_stream_v4(reinterpret_cast<const uint4*>(own) + v), f);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] = ((size_t)v * V + k < n) ? DT<T>::to_float(own[(size_t)v * V + k]) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) f[k] *= prescale;
    for (int q = 0; q < npeer; ++q) {
      float g[V];
      unpack16<T>(ld_sys_v4(peer_rec[q] + ((size_t)v << 4)), g);
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] += g[k];
    }
    const uint4 packed = pack16<T>(f);
    if (full) {
      st_v4(own + (size_t)v * V, packed);
    } else {
      const T* pe = reinterpret_cast<const T*>(&packed);
#pragma unroll
      for (int k = 0; k < V; ++k)
        if ((size_t)v * V + k < n) own[(size_t)v * V + k] = pe[k];
    }
    for (int k = 0; k < ndst; ++k) st_v4(dst[k] + ((size_t)v << 4), packed);
  }
}

// ---------------------------------------------------------------------------
// Copy a finished record (meta section + payload section, or raw section) from
// smem to `ndst` wire slots. dst[k] already points at the record start.
__device__ __forceinline__ void store_record(const BucketMeta* meta, const uint8_t* pay, uint32_t meta_bytes,
                                             uint32_t pay_bytes, uint8_t* const* dst, int ndst) {
  const uint32_t mv = meta_bytes >> 4, pv = pay_bytes >> 4;
  const uint4* m4 = reinterpret_cast<const uint4*>(meta);
  const uint4* p4 = reinterpret_cast<const uint4*>(pay);
  for (uint32_t i = threadIdx.x; i < mv + pv; i += blockDim.x) {
    uint4 v = i < mv ? m4[i] : p4[i - mv];
    for (int k = 0; k < ndst; ++k) st_v4(dst[k] + ((size_t)i << 4), v);
  }
}

// ---------------------------------------------------------------------------
// Read one pack group's `bits` bytes from a peer-written record.
__device__ __forceinline__ uint64_t load_group_word(const uint8_t* pay, uint32_t g, int bits) {
  const uint8_t* p = pay + (size_t)g * bits;
  switch (bits) {
    case 8: {
      uint2 v = ld_sys_v2(p);
      return (uint64_t)v.x | ((uint64_t)v.y << 32);
    }
    case 4: return ld_sys_u32(p);
    case 2: return ld_sys_u16(p);
    case 1: return ld_sys_u8(p);
    default: {
      uint64_t w = 0;
      for (int t = 0; t < bits; ++t) w |= (uint64_t)ld_sys_u8(p + t) << (8 * t);
      return w;
    }
  }
}

__device__ __forceinline__ BucketMeta load_meta(const uint8_t* rec, uint32_t b) {
  uint2 v = ld_sys_v2(rec + (size_t)b * 8u);
  BucketMeta m;
  m.unit = __uint_as_float(v.x);
  m.min = __uint_as_float(v.y);
  return m;
}

// Decode 8 values of group g of a compressed record.
__device__ __forceinline__ void decode_group(const uint8_t* rec, const uint8_t* pay, uint32_t g, uint32_t n,
                                             uint32_t B, int bits, float* v) {
  const uint32_t i0 = g * 8u;
  const uint64_t w = load_group_word(pay, g, bits);
  if ((B & 7u) == 0) {
    const BucketMeta m = load_meta(rec, i0 / B);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = decode_level(unpack1(w, j, bits), m.unit, m.min);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t i = min(i0 + j, n - 1);
      const BucketMeta m = load_meta(rec, i / B);
      v[j] = decode_level(unpack1(w, j, bits), m.unit, m.min);
    }
  }
}

// acc[i] += decode(record)[i]   (each thread owns the same groups in every call,
// so successive calls need no barrier between them)
template <typename T>
__device__ __forceinline__ void decode_add(const uint8_t* rec, const BlockDesc& d, float* acc) {
  const uint32_t n = block_n(d);
  const int bits = block_bits(d);
  if (bits >= kRawBits) {
    constexpr int V = DT<T>::kVec;
    const uint32_t nvec = div_up(n, (uint32_t)V);
    for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      uint4 raw = ld_sys_v4(rec + ((size_t)v << 4));
      float f[V];
      unpack16<T>(raw, f);
#pragma unroll
      for (int k = 0; k < V; ++k)
        if ((size_t)v * V + k < n) acc[(size_t)v * V + k] += f[k];
    }
    return;
  }
  const uint32_t B = d.bucket;
  const uint8_t* pay = rec + block_meta_bytes(n, B);
  const uint32_t groups = div_up(n, 8u);
  for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
    float v[8];
    decode_group(rec, pay, g, n, B, bits, v);
    const uint32_t i0 = g * 8u;
    if (i0 + 8u <= n) {
      float4* a = reinterpret_cast<float4*>(acc + i0);
      float4 lo = a[0], hi = a[1];
      lo.x += v[0]; lo.y += v[1]; lo.z += v[2]; lo.w += v[3];
      hi.x += v[4]; hi.y += v[5]; hi.z += v[6]; hi.w += v[7];
      a[0] = lo;
      a[1] = hi;
    } else {
      for (int j = 0; j < 8; ++j)
        if (i0 + j < n) acc[i0 + j] += v[j];
    }
  }
}

// out[elem_off + i] = T(decode(record)[i])
template <typename T>
__device__ __forceinline__ void decode_store(const uint8_t* rec, const BlockDesc& d, T* __restrict__ out_base) {
  const uint32_t n = block_n(d);
  const int bits = block_bits(d);
  T* out = out_base + d.elem_off;
  const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  if (bits >= kRawBits) {
    constexpr int V = DT<T>::kVec;
    const uint32_t nvec = div_up(n, (uint32_t)V);
    for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      uint4 raw = ld_sys_v4(rec + ((size_t)v << 4));
      if (out_vec && (v + 1) * V <= n) {
        st_v4(out + (size_t)v * V, raw);
      } else {
        const T* pe = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k)
          if ((size_t)v * V + k < n) out[(size_t)v * V + k] = pe[k];
      }
    }
    return;
  }
  const uint32_t B = d.bucket;
  const uint8_t* pay = rec + block_meta_bytes(n, B);
  const uint32_t groups = div_up(n, 8u);
  for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
    float v[8];
    decode_group(rec, pay, g, n, B, bits, v);
    const uint32_t i0 = g * 8u;
    if (out_vec && i0 + 8u <= n) {
      if (sizeof(T) == 4) {
        st_v4(out + i0, pack16<T>(v));
        st_v4(out + i0 + 4, pack16<T>(v + 4));
      } else {
        st_v4(out + i0, pack16<T>(v));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j < n) out[i0 + j] = DT<T>::from_float(v[j]);
    }
  }
}

// fp32 accumulate into a global fp32 scratch (generic reducer path)
template <typename T>
__device__ __forceinline__ void decode_add_global_f32(const uint8_t* rec, const BlockDesc& d,
                                                      float* __restrict__ accg) {
  const uint32_t n = block_n(d);
  const int bits = block_bits(d);
  if (bits >= kRawBits) {
    const T* src = reinterpret_cast<const T*>(rec);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) accg[i] += DT<T>::to_float(src[i]);
    return;
  }
  const uint32_t B = d.bucket;
  const uint8_t* pay = rec + block_meta_bytes(n, B);
  const uint32_t groups = div_up(n, 8u);
  for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
    float v[8];
    decode_group(rec, pay, g, n, B, bits, v);
    const uint32_t i0 = g * 8u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < n) accg[i0 + j] += v[j];
  }
}

}  // namespace dev
}  // namespace cgx
