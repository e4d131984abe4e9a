#include "symmetric_heap.h"

#include <unistd.h>

#include <cstring>
#include <stdexcept>
#include <string>

#include "../common/config.h"

namespace cgx {

void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::string msg = std::string("cgx: CUDA error in ") + what + ": " + cudaGetErrorString(e);
    throw std::runtime_error(msg);
  }
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

HeapLayout HeapLayout::make(int world, int max_lanes, size_t slot_bytes, size_t os_slot_bytes) {
  HeapLayout l;
  l.world = world;
  l.flag_stride = (uint32_t)round_up((size_t)(max_lanes < 1 ? 1 : max_lanes), 32);
  l.slot_bytes = (uint32_t)round_up(slot_bytes < 256 ? 256 : slot_bytes, 256);
  size_t flags_bytes = round_up((size_t)world * l.flag_stride * sizeof(uint32_t), 256);
  l.flags1_off = 0;
  l.flags2_off = flags_bytes;
  l.recv1_off = 2 * flags_bytes;
  l.recv2_off = l.recv1_off + (size_t)world * l.slot_bytes;
  l.os_slot_bytes = (uint32_t)round_up(os_slot_bytes, 256);
  l.os_off[0] = l.recv2_off + (size_t)world * l.slot_bytes;
  l.os_off[1] = l.os_off[0] + (size_t)world * l.os_slot_bytes;
  l.total = l.os_off[1] + (size_t)world * l.os_slot_bytes;
  return l;
}

SymmetricHeap::SymmetricHeap(int rank, int world, const HeapLayout& layout)
    : rank_(rank), world_(world), layout_(layout), bases_(world, nullptr), ipc_opened_(world, false) {
  cuda_check(cudaGetDevice(&device_), "cudaGetDevice");
  void* p = nullptr;
  cuda_check(cudaMalloc(&p, layout_.total), "cudaMalloc(symmetric heap)");
  cuda_check(cudaMemset(p, 0, layout_.total), "cudaMemset(symmetric heap)");
  bases_[rank_] = static_cast<uint8_t*>(p);
  void* h = nullptr;
  cuda_check(cudaHostAlloc(&h, 64, cudaHostAllocMapped), "cudaHostAlloc(status)");
  std::memset(h, 0, 64);
  status_host_ = static_cast<uint32_t*>(h);
  void* d = nullptr;
  cuda_check(cudaHostGetDevicePointer(&d, h, 0), "cudaHostGetDevicePointer(status)");
  status_dev_ = static_cast<uint32_t*>(d);
  cuda_check(cudaDeviceSynchronize(), "heap init sync");
  if (world_ == 1) connected_ = true;
}

SymmetricHeap::~SymmetricHeap() {
  for (int p = 0; p < world_; ++p) {
    if (p != rank_ && ipc_opened_[p] && bases_[p]) cudaIpcCloseMemHandle(bases_[p]);
  }
  if (bases_[rank_]) cudaFree(bases_[rank_]);
  if (status_host_) cudaFreeHost(status_host_);
}

void SymmetricHeap::connect_local(const std::vector<SymmetricHeap*>& all) {
  if ((int)all.size() != world_) throw std::invalid_argument("cgx: connect_local needs one heap per rank");
  for (int p = 0; p < world_; ++p) {
    if (p == rank_) continue;
    if (all[p]->layout_.total != layout_.total) throw std::invalid_argument("cgx: heap layouts differ");
    bases_[p] = all[p]->bases_[p];
    if (all[p]->device_ != device_) {
      int can = 0;
      cuda_check(cudaDeviceCanAccessPeer(&can, device_, all[p]->device_), "cudaDeviceCanAccessPeer");
      if (!can) throw std::runtime_error("cgx: no P2P access between in-process devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device_, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cuda_check(e, "cudaDeviceEnablePeerAccess");
      (void)cudaGetLastError();
    }
  }
  connected_ = true;
}

namespace {
struct HeapHello {
  cudaIpcMemHandle_t handle;
  uint64_t total;
  int32_t device;
  int32_t pid;
  uint64_t local_ptr;  // valid only inside the same pid
};
}  // namespace

void SymmetricHeap::connect_ipc(KVStore& store, const std::string& prefix) {
  HeapHello me;
  std::memset(&me, 0, sizeof(me));
  cuda_check(cudaIpcGetMemHandle(&me.handle, bases_[rank_]), "cudaIpcGetMemHandle");
  me.total = layout_.total;
  me.device = device_;
  me.pid = (int32_t)getpid();
  me.local_ptr = (uint64_t)(uintptr_t)bases_[rank_];
  std::vector<uint8_t> blob(sizeof(me));
  std::memcpy(blob.data(), &me, sizeof(me));
  store.set(prefix + "/heap/" + std::to_string(rank_), blob);
  for (int p = 0; p < world_; ++p) {
    if (p == rank_) continue;
    std::vector<uint8_t> got = store.get(prefix + "/heap/" + std::to_string(p));
    if (got.size() != sizeof(HeapHello)) throw std::runtime_error("cgx: malformed heap handle from peer");
    HeapHello other;
    std::memcpy(&other, got.data(), sizeof(other));
    if (other.total != layout_.total) throw std::runtime_error("cgx: peer heap layout differs (CGX_FUSION_BUFFER_SIZE_MB / CGX_LANES must match on all ranks)");
    if (other.pid == me.pid) {
      bases_[p] = reinterpret_cast<uint8_t*>((uintptr_t)other.local_ptr);
      continue;
    }
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, other.handle, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      std::string msg = "cgx: cudaIpcOpenMemHandle(rank " + std::to_string(p) + ", device " +
                        std::to_string(other.device) + ") failed: " + cudaGetErrorString(e) +
                        " -- the fused P2P path needs NVLink/PCIe peer access between all local GPUs";
      (void)cudaGetLastError();
      throw std::runtime_error(msg);
    }
    bases_[p] = static_cast<uint8_t*>(ptr);
    ipc_opened_[p] = true;
    log_msg(2, "cgx[%d]: mapped heap of rank %d (device %d) at %p", rank_, p, other.device, ptr);
  }
  // everybody has mapped everybody before anyone starts writing flags
  store.set(prefix + "/heap_ready/" + std::to_string(rank_), {1});
  for (int p = 0; p < world_; ++p)
    if (p != rank_) (void)store.get(prefix + "/heap_ready/" + std::to_string(p));
  connected_ = true;
}

}  // namespace cgx
