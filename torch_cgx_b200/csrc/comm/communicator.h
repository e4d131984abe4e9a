// Point-to-point byte transport between the ranks of a group -- the interface
// the generic reducers (SRA / Ring / all-to-all / broadcast) are written
// against. Same role as Communicator / CommunicatorLocal of the reference
// (/root/reference/src/common/communicator.h:28-63) whose implementations were
// MPI (mpi_communicator.cc) and host shared memory (shm_communicator.cc); here
// the implementations wrap a c10d backend: Gloo for host buffers, NCCL
// send/recv for device buffers (the reference's NCCL_Reduce pattern,
// /root/reference/src/common/nccl_reduce.cc:139-182) -- see pg/c10d_communicator.
// The intra-node fast path does NOT use this interface: it is the fused
// peer-memory kernel (reduce/fused_sra).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <vector>

namespace cgx {

struct P2POp {
  bool send;    // true: send `bytes` from buf to peer; false: receive into buf
  void* buf;
  size_t bytes;
  int peer;     // rank inside this communicator's group
};

class Communicator {
 public:
  virtual ~Communicator() = default;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // true: buffers are device memory and exchange() is stream-ordered on `stream`
  // false: buffers are host memory and exchange() returns when all ops completed
  virtual bool is_cuda() const = 0;
  // Post all ops as one group (no deadlock regardless of order) and complete them.
  virtual void exchange(const std::vector<P2POp>& ops, cudaStream_t stream) = 0;
  virtual void barrier() = 0;
};

}  // namespace cgx
