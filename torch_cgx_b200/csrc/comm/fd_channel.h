// File-descriptor passing between the ranks of one node (SCM_RIGHTS over abstract unix
// datagram sockets). CUDA VMM allocations and multicast objects are exported as POSIX file
// descriptors; a descriptor is only meaningful inside the process that owns it, so it has to
// travel through the kernel -- the socket *names* travel through the c10d Store.
// Role in the reference: the cudaIpcEventHandle / shm-name exchange over MPI
// (/root/reference/src/common/shm_communicator.cc:270-330), here for memory handles.
#pragma once
#include <cstdint>
#include <string>

namespace cgx {

struct FdMessage {
  int fd = -1;        // received descriptor (owned by the caller: close() it)
  int32_t kind = 0;   // user tag
  int32_t src = -1;   // sender's rank
};

class FdChannel {
 public:
  // binds an abstract-namespace datagram socket with a unique name
  explicit FdChannel(const std::string& hint);
  ~FdChannel();
  FdChannel(const FdChannel&) = delete;
  FdChannel& operator=(const FdChannel&) = delete;

  const std::string& name() const { return name_; }
  // send `fd` (+ two ints) to the channel called `peer_name`. Never blocks: returns false when
  // the receiver's queue is full (drain your own queue, then retry); throws on real errors
  bool try_send(const std::string& peer_name, int fd, int32_t kind, int32_t src) const;
  // waits up to timeout_ms for the next message; false on timeout; throws on error
  bool try_recv(FdMessage* out, int timeout_ms) const;

 private:
  int sock_ = -1;
  std::string name_;
};

}  // namespace cgx
