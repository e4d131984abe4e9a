// Communicator over a c10d backend: Gloo send/recv for host buffers, NCCL
// send/recv (one coalesced group per exchange, stream-ordered) for device
// buffers. Takes the place of MPICommunicator / the ncclGroupStart..End blocks
// of the reference (/root/reference/src/common/mpi_communicator.cc:37-84,
// /root/reference/src/common/nccl_reduce.cc:139-182).
#pragma once
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/distributed/c10d/Backend.hpp>

#include "../comm/communicator.h"

namespace cgx {

class C10dCommunicator : public Communicator {
 public:
  C10dCommunicator(c10::intrusive_ptr<c10d::Backend> backend, bool cuda, c10::DeviceIndex device)
      : backend_(std::move(backend)), cuda_(cuda), device_(device) {}
  int rank() const override { return backend_->getRank(); }
  int size() const override { return backend_->getSize(); }
  bool is_cuda() const override { return cuda_; }
  void exchange(const std::vector<P2POp>& ops, cudaStream_t stream) override;
  void barrier() override;

 private:
  c10::intrusive_ptr<c10d::Backend> backend_;
  bool cuda_;
  c10::DeviceIndex device_;
  int seq_ = 0;
};

}  // namespace cgx
