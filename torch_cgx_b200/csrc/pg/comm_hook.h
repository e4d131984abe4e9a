// Native DDP communication hook: the C++ twin of cgx_hook/CGXState
// (/root/reference/cgx_utils/allreduce_hooks.py:29-73). Registered straight on
// the c10d::Reducer, so the per-bucket path never enters Python: no GIL, no
// Python future callback -- in launch-bound regimes (small per-GPU batch) the
// Python hook costs ~50 us per bucket, which is visible against NCCL's built-in
// C++ allreduce hook.
#pragma once
#include <torch/csrc/distributed/c10d/comm.hpp>
#include <torch/csrc/distributed/c10d/reducer.hpp>

#include <atomic>
#include <memory>

#include "process_group_cgx.h"

namespace cgx {

struct HookState {
  c10::intrusive_ptr<ProcessGroupCGX> pg;
  int64_t layer_min_size = 1024;
  int bits = 32;
  int bucket_size = 1024;
  int register_step = 2;  // DDP rebuilds its buckets after the first iteration
  std::atomic<int64_t> step{0};
};

class CgxCommHook : public c10d::CommHookInterface {
 public:
  explicit CgxCommHook(std::shared_ptr<HookState> st) : st_(std::move(st)) {}
  c10::intrusive_ptr<c10::ivalue::Future> runHook(c10d::GradBucket& bucket) override;
  at::Tensor parseHookResult(const c10::IValue& result) override {
    return c10d::detail::parseCppCommHookResult(result);
  }

 private:
  std::shared_ptr<HookState> st_;
};

// registers the hook on `reducer` (a torch.distributed.Reducer) and returns the shared state
std::shared_ptr<HookState> register_native_hook(const std::shared_ptr<c10d::Reducer>& reducer,
                                                c10::intrusive_ptr<ProcessGroupCGX> pg, int64_t layer_min_size,
                                                int bits, int bucket_size, int register_step);

}  // namespace cgx
