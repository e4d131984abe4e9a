// Standalone block kernels: the same device primitives as the fused kernel,
// one launch over MANY blocks (grid-stride), used by the generic reducers
// (SRA / Ring over NCCL send/recv = the cross-node path and the
// "reference-structure" baseline), by the Python ops and by the tests.
// Covers K1-K5/K7 of SURVEY.md §2.3
// (/root/reference/src/common/compression/cuda_compression_operations.cu).
#include "block_device.cuh"
#include "launch.h"

namespace cgx {
using namespace dev;

namespace {
constexpr int kThreads = 512;

template <typename T, bool F32_SRC>
__global__ void __launch_bounds__(kThreads, 2)
quantize_blocks_kernel(const void* __restrict__ src, uint32_t base_elem, const BlockDesc* __restrict__ blocks,
                       uint32_t first, uint32_t count, uint8_t* __restrict__ wire, float prescale, RngKey rng,
                       T* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile& tile = *reinterpret_cast<Tile*>(smem_raw);
  for (uint32_t k = blockIdx.x; k < count; k += gridDim.x) {
    const uint32_t b = first + k;
    const BlockDesc d = blocks[b];
    const uint32_t n = block_n(d);
    const int bits = block_bits(d);
    uint8_t* rec = wire + d.wire_off;
    if (F32_SRC) {
      load_block_f32(reinterpret_cast<const float*>(src) + (d.elem_off - base_elem), n, tile.acc);
    } else {
      if (block_is_raw(d)) {
        send_raw<T>(reinterpret_cast<const T*>(src), d, prescale, rec);
        continue;
      }
      load_block<T>(reinterpret_cast<const T*>(src), d, prescale, tile.acc);
    }
    __syncthreads();
    if (block_is_raw(d)) {  // only reachable with F32_SRC
      constexpr int V = DT<T>::kVec;
      const uint32_t nvec = div_up(n, (uint32_t)V);
      T* o = out ? out + d.elem_off : nullptr;
      for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[V];
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = ((size_t)v * V + j < n) ? tile.acc[(size_t)v * V + j] : 0.f;
        const uint4 packed = pack16<T>(f);
        st_v4(rec + ((size_t)v << 4), packed);
        if (o) {
          const T* pe = reinterpret_cast<const T*>(&packed);
#pragma unroll
          for (int j = 0; j < V; ++j)
            if ((size_t)v * V + j < n) o[(size_t)v * V + j] = pe[j];
        }
      }
      __syncthreads();
      continue;
    }
    compute_meta(tile.acc, n, d.bucket, bits, tile.meta, tile.inv);
    __syncthreads();
    if (out)
      pack_block<T, true>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, out);
    else
      pack_block<T, false>(tile.acc, d, tile.meta, tile.inv, tile.pay, rng, b, nullptr);
    __syncthreads();
    store_record(tile.meta, tile.pay, block_meta_bytes(n, d.bucket), block_payload_bytes(n, bits), &rec, 1);
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
dequantize_blocks_kernel(const uint8_t* __restrict__ wire, const BlockDesc* __restrict__ blocks, uint32_t first,
                         uint32_t count, T* __restrict__ dst) {
  for (uint32_t k = blockIdx.x; k < count; k += gridDim.x) {
    const BlockDesc d = blocks[first + k];
    decode_store<T>(wire + d.wire_off, d, dst);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
accumulate_blocks_kernel(const uint8_t* __restrict__ wire, const BlockDesc* __restrict__ blocks, uint32_t first,
                         uint32_t count, float* __restrict__ acc, uint32_t base_elem,
                         const T* __restrict__ init_src, float prescale) {
  for (uint32_t k = blockIdx.x; k < count; k += gridDim.x) {
    const BlockDesc d = blocks[first + k];
    const uint32_t n = block_n(d);
    float* a = acc + (d.elem_off - base_elem);
    if (init_src) {
      const T* s = init_src + d.elem_off;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) a[i] = DT<T>::to_float(s[i]) * prescale;
    }
    if (wire) {
      if (init_src) __syncthreads();
      decode_add_global_f32<T>(wire + d.wire_off, d, a);
    }
  }
}

template <typename T>
__global__ void scale_kernel(T* __restrict__ data, uint64_t n, float scale) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    data[i] = DT<T>::from_float(DT<T>::to_float(data[i]) * scale);
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ sum, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    sum[i] = DT<T>::from_float(DT<T>::to_float(x[i]) + DT<T>::to_float(y[i]));
}
template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = DT<D>::from_float(DT<S>::to_float(src[i]));
}

int grid_for_blocks(uint32_t count) {
  static int cap = 0;
  if (cap == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cap = sms > 0 ? sms * 2 : 296;
  }
  return (int)(count < (uint32_t)cap ? (count ? count : 1) : cap);
}
int grid_for_elems(uint64_t n, int threads) {
  uint64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (int)(g > 148 * 16 ? 148 * 16 : g);
}

template <typename K>
cudaError_t set_smem(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile));
}

#define CGX_DISPATCH(dtype, ...)                          \
  switch (dtype) {                                        \
    case kF32: { using T = float; __VA_ARGS__; break; }   \
    case kF16: { using T = __half; __VA_ARGS__; break; }  \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: return cudaErrorInvalidValue;                \
  }

}  // namespace

cudaError_t launch_quantize_blocks(const void* src, int dtype, const BlockDesc* blocks, uint32_t first,
                                   uint32_t count, uint8_t* wire, float prescale, const RngKey& rng,
                                   cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, {
    auto k = quantize_blocks_kernel<T, false>;
    cudaError_t e = set_smem(k);
    if (e != cudaSuccess) return e;
    k<<<grid_for_blocks(count), kThreads, sizeof(Tile), stream>>>(src, 0u, blocks, first, count, wire, prescale,
                                                                  rng, (T*)nullptr);
  });
  return cudaGetLastError();
}

cudaError_t launch_quantize_blocks_f32(const float* src_f32, uint32_t base_elem, int dtype,
                                       const BlockDesc* blocks, uint32_t first, uint32_t count, uint8_t* wire,
                                       const RngKey& rng, void* out, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, {
    auto k = quantize_blocks_kernel<T, true>;
    cudaError_t e = set_smem(k);
    if (e != cudaSuccess) return e;
    k<<<grid_for_blocks(count), kThreads, sizeof(Tile), stream>>>(src_f32, base_elem, blocks, first, count, wire,
                                                                  1.0f, rng, (T*)out);
  });
  return cudaGetLastError();
}

cudaError_t launch_dequantize_blocks(const uint8_t* wire, int dtype, const BlockDesc* blocks, uint32_t first,
                                     uint32_t count, void* dst, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, {
    dequantize_blocks_kernel<T><<<grid_for_blocks(count), kThreads, 0, stream>>>(wire, blocks, first, count, (T*)dst);
  });
  return cudaGetLastError();
}

cudaError_t launch_accumulate_blocks_f32(const uint8_t* wire, int dtype, const BlockDesc* blocks,
                                         uint32_t first, uint32_t count, float* acc_f32, uint32_t base_elem,
                                         const void* init_src, float prescale, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, {
    accumulate_blocks_kernel<T><<<grid_for_blocks(count), kThreads, 0, stream>>>(
        wire, blocks, first, count, acc_f32, base_elem, (const T*)init_src, prescale);
  });
  return cudaGetLastError();
}

cudaError_t launch_scale_inplace(void* data, int dtype, uint64_t n, float scale, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, { scale_kernel<T><<<grid_for_elems(n, 256), 256, 0, stream>>>((T*)data, n, scale); });
  return cudaGetLastError();
}

cudaError_t launch_add(const void* x, const void* y, void* sum, int dtype, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  CGX_DISPATCH(dtype, {
    add_kernel<T><<<grid_for_elems(n, 256), 256, 0, stream>>>((const T*)x, (const T*)y, (T*)sum, n);
  });
  return cudaGetLastError();
}

cudaError_t launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, uint64_t n,
                           cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int g = grid_for_elems(n, 256);
#define CGX_CONV(S, D) convert_kernel<S, D><<<g, 256, 0, stream>>>((const S*)src, (D*)dst, n)
  if (src_dtype == kF32 && dst_dtype == kF16) CGX_CONV(float, __half);
  else if (src_dtype == kF32 && dst_dtype == kBF16) CGX_CONV(float, __nv_bfloat16);
  else if (src_dtype == kF16 && dst_dtype == kF32) CGX_CONV(__half, float);
  else if (src_dtype == kBF16 && dst_dtype == kF32) CGX_CONV(__nv_bfloat16, float);
  else if (src_dtype == kF32 && dst_dtype == kF32) CGX_CONV(float, float);
  else if (src_dtype == kF16 && dst_dtype == kBF16) CGX_CONV(__half, __nv_bfloat16);
  else if (src_dtype == kBF16 && dst_dtype == kF16) CGX_CONV(__nv_bfloat16, __half);
  else if (src_dtype == kF16 && dst_dtype == kF16) CGX_CONV(__half, __half);
  else if (src_dtype == kBF16 && dst_dtype == kBF16) CGX_CONV(__nv_bfloat16, __nv_bfloat16);
  else return cudaErrorInvalidValue;
#undef CGX_CONV
  return cudaGetLastError();
}

}  // namespace cgx
