#include "fused_sra.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "../common/config.h"
#include "../kernels/launch.h"

namespace cgx {

FusedSra::FusedSra(SymmetricHeap* heap, int max_lanes, int64_t timeout_ms, uint32_t min_lane_elems)
    : heap_(heap), max_lanes_(max_lanes), timeout_ns_((uint64_t)timeout_ms * 1000000ull),
      min_lane_elems_(min_lane_elems) {
  // NVLS: multimem.st for phase B / one-shot, multimem.ld_reduce for raw items (CGX_NVLS=0 disables
  // the heap's multicast mapping altogether; CGX_NVLS_REDUCE=0 keeps the two-shot raw path)
  use_mc_ = heap_->has_multicast() && heap_->world() > 1;
  use_mc_reduce_ = use_mc_ && env_bool("CGX_NVLS_REDUCE", true);
  stages_ = (int)env_int("CGX_STAGES", 0);  // pipeline depth experiment, see pick_stages()
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
    if (kv.second->d_items) cudaFree(kv.second->d_items);
    if (kv.second->d_item_first) cudaFree(kv.second->d_item_first);
  }
}

const DevicePlan* FusedSra::prepare(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                    cudaStream_t stream, int max_lanes) {
  return prepare_impl(layers, dtype, skip_incomplete, stream, world(), heap_->layout().slot_bytes, max_lanes);
}

const DevicePlan* FusedSra::prepare_oneshot(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                            cudaStream_t stream, int max_lanes) {
  if (heap_->layout().os_slot_bytes == 0) return nullptr;
  return prepare_impl(layers, dtype, skip_incomplete, stream, 1, heap_->layout().os_slot_bytes, max_lanes);
}

const DevicePlan* FusedSra::prepare_impl(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete,
                                         cudaStream_t stream, int plan_world, uint32_t capacity, int max_lanes) {
  PlanOptions opt;
  opt.world = plan_world;
  opt.lanes = (max_lanes > 0 && max_lanes < max_lanes_) ? max_lanes : max_lanes_;
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
        if (kv.second->d_items) cudaFree(kv.second->d_items);
        if (kv.second->d_item_first) cudaFree(kv.second->d_item_first);
      }
      cache_.clear();
      ++generation_;
    }
    auto dp = std::make_unique<DevicePlan>();
    dp->plan = build_plan(layers, opt);
    {
      bool ok = true;
      for (const BlockDesc& bd : dp->plan.blocks) {
        const int bb = block_bits(bd);
        ok = ok && (block_is_raw(bd) || bb == 2 || bb == 4 || bb == 8);
      }
      dp->multicast_ok = ok;
    }
    if (dp->plan.max_chunk_wire <= capacity && !dp->plan.items.empty()) {
      const size_t ib = dp->plan.items.size() * sizeof(WarpItem);
      const size_t fb = dp->plan.item_first.size() * sizeof(uint32_t);
      cuda_check(cudaMalloc((void**)&dp->d_items, ib), "cudaMalloc(plan items)");
      cuda_check(cudaMalloc((void**)&dp->d_item_first, fb), "cudaMalloc(plan lanes)");
      // pageable source: the runtime stages it before returning, so the vectors may be reused
      cuda_check(cudaMemcpyAsync(dp->d_items, dp->plan.items.data(), ib, cudaMemcpyHostToDevice, stream),
                 "upload plan items");
      cuda_check(cudaMemcpyAsync(dp->d_item_first, dp->plan.item_first.data(), fb, cudaMemcpyHostToDevice, stream),
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

// Pipeline depth of the fused kernel: every chunk's item list can be cut into S pieces, each with
// its own flag value, so that the NVLink flight of one piece overlaps the computation of the next.
// Measured on 2xB200 (profiles/r2/stages_2gpu.md) this LOSES: every extra stage costs one more
// fence.acq_rel.sys per phase, and a system fence issued while the SM keeps streaming peer stores
// only completes when that traffic drains (64 MB 4-bit: 64 us at S=1, 77 us at S=2, 123 us at S=4).
// The default therefore stays 1; CGX_STAGES=2..4 keeps the experiment reproducible.
int FusedSra::pick_stages(const DevicePlan&) const {
  const int w = world();
  const int cap = w > 0 ? (32 / w < 4 ? 32 / w : 4) : 1;
  int s = stages_ <= 0 ? 1 : stages_;
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}

void FusedSra::launch(const DevicePlan& dp, void* data, float prescale, const RngParams& rng, cudaStream_t stream,
                      bool oneshot) {
  if (!heap_->connected()) throw std::runtime_error("cgx: symmetric heap is not connected");
  if (dp.plan.items.empty()) return;
  ++epoch_;
  SraParams p;
  std::memset(&p, 0, sizeof(p));
  p.data = data;
  p.items = dp.d_items;
  p.item_first = dp.d_item_first;
  p.rank = rank();
  p.world = world();
  p.lanes = dp.plan.lanes;
  p.dtype = dp.plan.dtype;
  p.epoch_hint = epoch_;
  p.prescale = prescale;
  p.rng = make_rng_key(rng, rank(), 0);
  p.slot_bytes = oneshot ? heap_->layout().os_slot_bytes : heap_->layout().slot_bytes;
  p.flag_stride = heap_->layout().flag_stride;
  p.os_parity_stride = (uint32_t)(heap_->layout().os_off[1] - heap_->layout().os_off[0]);
  for (int q = 0; q < kMaxPeers; ++q) {
    const bool valid = q < world();
    // one-shot: recv1 is region 0 of the alternating pair; the kernel adds the parity offset
    p.recv1[q] = valid ? (oneshot ? heap_->oneshot(q, 0) : heap_->recv1(q)) : nullptr;
    p.recv2[q] = valid ? heap_->recv2(q) : nullptr;
    p.flags1[q] = valid ? heap_->flags1(q) : nullptr;
    p.flags2[q] = valid ? heap_->flags2(q) : nullptr;
  }
  const bool mc = use_mc_ && dp.multicast_ok;
  p.mc_recv1 = (mc && oneshot) ? heap_->mc_oneshot(0) : nullptr;
  p.mc_recv2 = (mc && !oneshot) ? heap_->mc_recv2() : nullptr;
  p.mc_reduce = (mc && !oneshot && use_mc_reduce_) ? 1 : 0;
  p.status = heap_->status_device();
  p.abort_word = heap_->abort_device();
  p.timeout_ns = timeout_ns_;
  p.sync = heap_->sync_device();
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
  p.uniform_bits = dp.plan.uniform_bits;
  p.slice_elems = (int)dp.plan.slice_elems;
  p.oneshot = oneshot ? 1 : 0;
  p.stages = oneshot ? 1 : pick_stages(dp);
  cuda_check(launch_sra_fused(p, stream), "launch_sra_fused");
  ++launches_;
}

std::string FusedSra::status_message() const {
  const uint32_t s = heap_->status_host();
  if (s == 0) return std::string();
  const uint32_t code = s & 0xFF, peer = (s >> 8) & 0xFF, lane = s >> 16;
  if (code == kSraAborted)
    return "cgx: fused allreduce kernel on rank " + std::to_string(rank()) + " was aborted while waiting for rank " +
           std::to_string(peer) + " (lane " + std::to_string(lane) + ")";
  return "cgx: fused allreduce kernel timed out on rank " + std::to_string(rank()) + " waiting for rank " +
         std::to_string(peer) + " (phase " + std::to_string(code) + ", lane " + std::to_string(lane) +
         "); a peer died, is stuck, or issued collectives in a different order "
         "(raise CGX_TIMEOUT_MS if the job is just slow)";
}

void FusedSra::check_status() {
  // sticky: a heap whose kernel gave up is out of step with its peers (epochs, flags) and must
  // not be reused; clear_status() exists for tests that inject faults on purpose
  const std::string msg = status_message();
  if (!msg.empty()) throw std::runtime_error(msg);
}

void FusedSra::clear_status() {
  heap_->clear_status();
  heap_->request_abort(false);
}

}  // namespace cgx
