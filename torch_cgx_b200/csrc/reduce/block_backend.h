// Device abstraction for the generic reducers: the four block primitives on
// either host memory (common/block_ops.h) or device memory (kernels/quantize.cu).
// Reference role: Compressor::{Compress,Decompress,Add} over GPUContext
// (/root/reference/src/common/compressor.cc:62-220, gpu_context.h:67-100).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <memory>

#include "../common/philox.h"
#include "../common/plan.h"

namespace cgx {

class BlockBackend {
 public:
  virtual ~BlockBackend() = default;
  virtual bool is_cuda() const = 0;
  // scratch memory of the backend's kind (zero-initialised)
  virtual void* alloc(size_t bytes) = 0;
  virtual void release(void* p) = 0;
  // make `plan` usable by the calls below (uploads the table for CUDA)
  virtual void bind(const Plan& plan, cudaStream_t stream) = 0;
  // wire = quantize(src[T] * prescale) for blocks [first, first+count)
  virtual void quantize(const void* src, uint32_t first, uint32_t count, uint8_t* wire, float prescale,
                        const RngKey& rng, cudaStream_t stream) = 0;
  // acc_f32[elem_off - base_elem + i] = float(init_src[elem_off + i]) * prescale   (init_src != nullptr)
  // acc_f32[...] += decode(wire)                                                   (wire != nullptr)
  virtual void accumulate(const uint8_t* wire, uint32_t first, uint32_t count, float* acc_f32, uint32_t base_elem,
                          const void* init_src, float prescale, cudaStream_t stream) = 0;
  // wire = quantize(acc_f32); out[T] = decode(wire) when out != nullptr
  virtual void quantize_f32(const float* acc_f32, uint32_t base_elem, uint32_t first, uint32_t count, uint8_t* wire,
                            const RngKey& rng, void* out, cudaStream_t stream) = 0;
  // dst[T] = decode(wire)
  virtual void dequantize(const uint8_t* wire, uint32_t first, uint32_t count, void* dst, cudaStream_t stream) = 0;
  virtual void copy(void* dst, const void* src, size_t bytes, cudaStream_t stream) = 0;
};

std::unique_ptr<BlockBackend> make_cpu_block_backend();
std::unique_ptr<BlockBackend> make_cuda_block_backend();

}  // namespace cgx
