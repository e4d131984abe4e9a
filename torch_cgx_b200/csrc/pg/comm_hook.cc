#include "comm_hook.h"

#include "../common/layers.h"

namespace cgx {

c10::intrusive_ptr<c10::ivalue::Future> CgxCommHook::runHook(c10d::GradBucket& bucket) {
  HookState& s = *st_;
  const int64_t step = s.step.load(std::memory_order_relaxed);
  const size_t idx = bucket.getIndex();
  if (step == s.register_step) {
    // tell the backend the layer structure of this bucket: 1-D / small layers stay uncompressed
    const std::vector<at::Tensor> grads = bucket.getGradients();
    for (size_t i = 0; i < grads.size(); ++i) {
      const at::Tensor& g = grads[i];
      const bool compress = g.dim() > 1 && g.numel() >= s.layer_min_size;
      LayerRegistry::instance().register_layer((unsigned)idx, (unsigned)i, g.numel(), compress ? s.bits : kRawBits,
                                               s.bucket_size);
    }
  }
  const bool registered = step >= s.register_step;
  if (bucket.isLast()) s.step.fetch_add(1, std::memory_order_relaxed);
  // average == true: the 1/world scale is applied inside the kernel, before quantization
  auto work = s.pg->allreduce_bucket(bucket.getBufferRef(), registered ? (int64_t)idx : -1, true);
  return work->getFuture();
}

std::shared_ptr<HookState> register_native_hook(const std::shared_ptr<c10d::Reducer>& reducer,
                                                c10::intrusive_ptr<ProcessGroupCGX> pg, int64_t layer_min_size,
                                                int bits, int bucket_size, int register_step) {
  auto st = std::make_shared<HookState>();
  st->pg = std::move(pg);
  st->layer_min_size = layer_min_size;
  st->bits = bits;
  st->bucket_size = bucket_size;
  st->register_step = register_step;
  reducer->register_comm_hook(std::make_unique<CgxCommHook>(st));
  return st;
}

}  // namespace cgx
