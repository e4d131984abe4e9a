#include "c10d_communicator.h"

#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>

namespace cgx {

void C10dCommunicator::exchange(const std::vector<P2POp>& ops, cudaStream_t stream) {
  // one message per (src, dst) pair and exchange: the exchange counter is a safe tag. It must
  // advance on EVERY logical exchange, including those in which this rank has nothing to move.
  const int tag = (seq_++) & 0x3FFF;
  if (ops.empty()) return;
  if (cuda_) {
    c10::cuda::CUDAGuard dg(device_);
    c10::cuda::CUDAStreamGuard sg(c10::cuda::getStreamFromExternal(stream, device_));
    auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device_);
    backend_->startCoalescing();
    for (const P2POp& op : ops) {
      std::vector<at::Tensor> t{at::from_blob(op.buf, {(int64_t)op.bytes}, opts)};
      if (op.send)
        backend_->send(t, op.peer, tag);
      else
        backend_->recv(t, op.peer, tag);
    }
    auto work = backend_->endCoalescing();
    if (work) work->wait();  // CUDA semantics: the current (= our) stream waits, the host does not
    return;
  }
  auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCPU);
  std::vector<c10::intrusive_ptr<c10d::Work>> works;
  works.reserve(ops.size());
  // post receives first, then sends
  for (int pass = 0; pass < 2; ++pass)
    for (const P2POp& op : ops) {
      if (op.send != (pass == 1)) continue;
      std::vector<at::Tensor> t{at::from_blob(op.buf, {(int64_t)op.bytes}, opts)};
      works.push_back(op.send ? backend_->send(t, op.peer, tag) : backend_->recv(t, op.peer, tag));
    }
  for (auto& w : works) w->wait();
}

void C10dCommunicator::barrier() { backend_->barrier()->wait(); }

}  // namespace cgx
