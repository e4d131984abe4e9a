#include "fused_sra.h"

#include <stdexcept>
#include <string>

#include "../common/config.h"
#include "../kernels/launch.h"

namespace cgx {

FusedSra::FusedSra(SymmetricHeap* heap, int max_lanes, int64_t timeout_ms, uint32_t min_lane_elems)
    : heap_(heap), max_lanes_(max_lanes), timeout_ns_((uint64_t)timeout_ms * 1000000ull),
      min_lane_elems_(min_lane_elems) {
  {
    // CGX_KERNEL: "warp" (default: 16 warps/SM, 128 regs), "warp2" (32 warps/SM, 64 regs), "block" (v1)
    const std::string k = env_str("CGX_KERNEL", "warp");
    variant_ = k == "block" ? 1 : (k == "warp2" ? 2 : 0);
  }
  if (max_lanes_ < 1) max_lanes_ = 1;
  if ((uint32_t)max_lanes_ > heap_->layout().flag_stride) max_lanes_ = (int)heap_->layout().flag_stride;
}

void FusedSra::enable_trace(bool on) {
  trace_on_ = on;
  if (on && d_trace_ == nullptr)
    cuda_check(cudaMalloc((void**)&d_trace_, (size_t)heap_->layout().flag_stride * 8 * sizeof(unsigned long long)),
               "cudaMalloc(trace)");
}

std::vector<uint64_t> FusedSra::read_trace() {
  std::vector<uint64_t> out;
  if (d_trace_ == nullptr || last_lanes_ == 0) return out;
  cuda_check(cudaDeviceSynchronize(), "trace sync");
  out.resize((size_t)last_lanes_ * 8);
  cuda_check(cudaMemcpy(out.data(), d_trace_, out.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost), "trace copy");
  return out;
}

FusedSra::~FusedSra() {
  if (d_trace_) cudaFree(d_trace_);
  for (auto& kv : cache_) {
    if (kv.second->d_blocks) cudaFree(kv.second->d_blocks);
    if (kv.second->d_lane_first) cudaFree(kv.second->d_lane_first);
  }
}

const DevicePlan* FusedSra::prepare(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                    cudaStream_t stream) {
  return prepare_impl(layers, dtype, skip_incomplete, stream, world(), heap_->layout().slot_bytes);
}

const DevicePlan* FusedSra::prepare_oneshot(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                            cudaStream_t stream) {
  if (heap_->layout().os_slot_bytes == 0) return nullptr;
  return prepare_impl(layers, dtype, skip_incomplete, stream, 1, heap_->layout().os_slot_bytes);
}

const DevicePlan* FusedSra::prepare_impl(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                         cudaStream_t stream, int plan_world, uint32_t capacity) {
  PlanOptions opt;
  opt.world = plan_world;
  opt.lanes = max_lanes_;
  opt.dtype = dtype;
  opt.skip_incomplete = skip_incomplete;
  opt.min_lane_elems = min_lane_elems_;
  const uint64_t key = plan_key(layers, opt);
  auto it = cache_.find(key);
  if (it == cache_.end()) {
    if (cache_.size() >= kMaxCachedPlans) {
      // workloads with ever-changing message sizes: drop everything (kernels that still read the
      // old tables are ordered before the frees by the synchronize) and invalidate derived caches
      cuda_check(cudaStreamSynchronize(stream), "sync before trimming the plan cache");
      for (auto& kv : cache_) {
        if (kv.second->d_blocks) cudaFree(kv.second->d_blocks);
        if (kv.second->d_lane_first) cudaFree(kv.second->d_lane_first);
      }
      cache_.clear();
      ++generation_;
    }
    auto dp = std::make_unique<DevicePlan>();
    dp->plan = build_plan(layers, opt);
    int ub = -1;
    for (const BlockDesc& bd : dp->plan.blocks) {
      if (block_is_raw(bd)) continue;
      const int bb = block_bits(bd);
      ub = (ub == -1 || ub == bb) ? bb : 0;
    }
    dp->uniform_bits = ub > 0 ? ub : 0;
    if (dp->plan.max_chunk_wire <= capacity && !dp->plan.blocks.empty()) {
      const size_t bb = dp->plan.blocks.size() * sizeof(BlockDesc);
      const size_t lb = dp->plan.lane_first.size() * sizeof(uint32_t);
      cuda_check(cudaMalloc((void**)&dp->d_blocks, bb), "cudaMalloc(plan blocks)");
      cuda_check(cudaMalloc((void**)&dp->d_lane_first, lb), "cudaMalloc(plan lanes)");
      // pageable source: the runtime stages it before returning, so the vectors may be reused
      cuda_check(cudaMemcpyAsync(dp->d_blocks, dp->plan.blocks.data(), bb, cudaMemcpyHostToDevice, stream),
                 "upload plan blocks");
      cuda_check(cudaMemcpyAsync(dp->d_lane_first, dp->plan.lane_first.data(), lb, cudaMemcpyHostToDevice, stream),
                 "upload plan lanes");
    }
    if (log_level() >= 2) log_msg(2, "cgx[%d]: new plan %s", rank(), describe_plan(dp->plan).c_str());
    it = cache_.emplace(key, std::move(dp)).first;
  }
  const DevicePlan* dp = it->second.get();
  if (dp->plan.max_chunk_wire > capacity) return nullptr;
  return dp;
}

void FusedSra::run(const DevicePlan& dp, void* data, float prescale, const RngParams& rng, cudaStream_t stream) {
  launch(dp, data, prescale, rng, stream, false);
}

void FusedSra::run_oneshot(const DevicePlan& dp, void* data, float prescale, const RngParams& rng,
                           cudaStream_t stream) {
  launch(dp, data, prescale, rng, stream, true);
}

void FusedSra::launch(const DevicePlan& dp, void* data, float prescale, const RngParams& rng, cudaStream_t stream,
                      bool oneshot) {
  if (!heap_->connected()) throw std::runtime_error("cgx: symmetric heap is not connected");
  if (dp.plan.blocks.empty()) return;
  ++epoch_;
  SraParams p;
  p.data = data;
  p.blocks = dp.d_blocks;
  p.lane_first = dp.d_lane_first;
  p.rank = rank();
  p.world = world();
  p.lanes = dp.plan.lanes;
  p.dtype = dp.plan.dtype;
  p.epoch = epoch_;
  p.prescale = prescale;
  p.rng = make_rng_key(rng, rank(), 0);
  // one-shot calls alternate between two dedicated regions: a peer can be at most one call ahead
  const int parity = oneshot ? (int)(oneshot_calls_++ & 1u) : 0;
  p.slot_bytes = oneshot ? heap_->layout().os_slot_bytes : heap_->layout().slot_bytes;
  p.flag_stride = heap_->layout().flag_stride;
  for (int q = 0; q < kMaxPeers; ++q) {
    const bool valid = q < world();
    p.recv1[q] = valid ? (oneshot ? heap_->oneshot(q, parity) : heap_->recv1(q)) : nullptr;
    p.recv2[q] = valid ? heap_->recv2(q) : nullptr;
    p.flags1[q] = valid ? heap_->flags1(q) : nullptr;
    p.flags2[q] = valid ? heap_->flags2(q) : nullptr;
  }
  p.status = heap_->status_device();
  p.timeout_ns = timeout_ns_;
  p.trace = nullptr;
  if (trace_on_ && d_trace_) {
    // slot 0 takes a minimum, the rest maxima: 0xFF.. / 0 initialisation per lane
    std::vector<unsigned long long> init((size_t)dp.plan.lanes * 8, 0ull);
    for (int l = 0; l < dp.plan.lanes; ++l) init[(size_t)l * 8] = ~0ull;
    cuda_check(cudaMemcpyAsync(d_trace_, init.data(), init.size() * sizeof(unsigned long long),
                               cudaMemcpyHostToDevice, stream),
               "trace init");
    p.trace = d_trace_;
  }
  last_lanes_ = dp.plan.lanes;
  p.variant = oneshot ? 3 : variant_;
  p.uniform_bits = dp.uniform_bits;
  cuda_check(launch_sra_fused(p, stream), "launch_sra_fused");
  ++launches_;
}

void FusedSra::check_status() {
  uint32_t s = heap_->status_host();
  if (s == 0) return;
  heap_->clear_status();
  const uint32_t code = s & 0xFF, peer = (s >> 8) & 0xFF, lane = s >> 16;
  throw std::runtime_error("cgx: fused allreduce kernel timed out on rank " + std::to_string(rank()) +
                           " waiting for rank " + std::to_string(peer) + " (phase " + std::to_string(code) +
                           ", lane " + std::to_string(lane) +
                           "); a peer died, is stuck, or issued collectives in a different order "
                           "(raise CGX_TIMEOUT_MS if the job is just slow)");
}

}  // namespace cgx
