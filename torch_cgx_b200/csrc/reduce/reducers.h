// Transport-agnostic reducers: compressed Scatter-Reduce-AllGather, ring, and
// the debug all-to-all, written against Communicator + BlockBackend so the same
// code runs on host memory over Gloo (CPU tensors, tests) and on device memory
// over NCCL send/recv (cross-node stage, and the "reference-structure" baseline
// CGX_INNER_COMMUNICATOR_TYPE=NCCL).
//
// Reference: Reducer / MPI_Allreduce_ScatterReduceAllgather / MPI_Allreduce_Ring /
// NCCL_Reduce (/root/reference/src/common/reducer.{h,cc},
// scatter_reduce_allgather.cc:94-413, ring.cc:57-226, nccl_reduce.cc:89-213).
// Differences: accumulation is fp32 in rank order (bit-identical to the fused
// kernel and to the CPU oracle); quantize/dequantize run as ONE launch per
// chunk over the plan's block table instead of 2-3 launches per layer slice.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <memory>
#include <unordered_map>
#include <vector>

#include "../comm/communicator.h"
#include "../common/config.h"
#include "../common/plan.h"
#include "../common/sra_sim.h"
#include "block_backend.h"

namespace cgx {

class Reducer {
 public:
  Reducer(Communicator* comm, BlockBackend* ops);
  virtual ~Reducer();
  virtual const char* name() const = 0;
  // in-place allreduce of data[T] restricted to `layers`
  virtual void allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                         float prescale, const RngParams& rng, cudaStream_t stream) = 0;
  // raw broadcast of `bytes` from group rank `root`
  void broadcast(void* data, size_t bytes, int root, cudaStream_t stream);
  // Compressed broadcast of `layers` (reference: Reducer::Broadcast with do_compression,
  // /root/reference/src/common/reducer.cc:96-160): the root quantizes, decodes its own bytes back
  // into `data` (so it holds exactly what every receiver will decode) and sends the SAME packed
  // bytes to every peer; receivers decode. Raw layers travel as T inside the same record.
  void broadcast_compressed(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                            int root, const RngParams& rng, cudaStream_t stream);
  uint64_t bytes_sent() const { return bytes_sent_; }

 protected:
  // world <= 0: one chunk per rank of the communicator; otherwise that many chunks
  const Plan& plan_for(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete, int world = 0);
  uint8_t* scratch(int slot, size_t bytes);  // grow-only scratch buffers of the backend's memory kind
  static size_t row_bytes(const Plan& p) { return ((size_t)p.max_chunk_wire + 255) / 256 * 256; }
  // [first element, one past last element) spanned by the blocks of a chunk / of the whole plan
  static void chunk_span(const Plan& p, int chunk, uint32_t* lo, uint32_t* hi);

  Communicator* comm_;
  BlockBackend* ops_;
  uint64_t bytes_sent_ = 0;

 private:
  std::unordered_map<uint64_t, std::unique_ptr<Plan>> plans_;
  struct Buf {
    uint8_t* p = nullptr;
    size_t cap = 0;
  };
  std::vector<Buf> bufs_;
};

class SraReducer : public Reducer {
 public:
  using Reducer::Reducer;
  const char* name() const override { return "SRA"; }
  void allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete, float prescale,
                 const RngParams& rng, cudaStream_t stream) override;
};

class RingReducer : public Reducer {
 public:
  using Reducer::Reducer;
  const char* name() const override { return "RING"; }
  void allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete, float prescale,
                 const RngParams& rng, cudaStream_t stream) override;
};

// Debug variant (CGX_DEBUG_ALL_TO_ALL_REDUCTION): every rank quantizes the whole
// buffer once, sends it to everybody and adds everything it receives to its own
// raw values -- one quantization error per peer, replicas NOT bit-identical.
class AllToAllReducer : public Reducer {
 public:
  using Reducer::Reducer;
  const char* name() const override { return "ALLTOALL"; }
  void allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete, float prescale,
                 const RngParams& rng, cudaStream_t stream) override;
};

std::unique_ptr<Reducer> make_reducer(ReductionType type, Communicator* comm, BlockBackend* ops);

}  // namespace cgx
