// Python bindings: module torch_cgx_b200._C
//  * the c10d backend + the reference's pybind surface (register_layer,
//    set_quantization_bits, set_quantization_bucket_size;
//    /root/reference/src/ProcessGroupCGX.cc:837-857)
//  * planning / CPU oracle / standalone kernels for ops and tests
//  * LocalSraGroup: W virtual ranks of the fused kernel inside ONE process on
//    ONE GPU (W heaps, W streams) so the full protocol is testable on a single
//    device.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <pybind11/chrono.h>
#include <torch/extension.h>

#include <memory>
#include <tuple>
#include <vector>

#include "comm/fd_channel.h"
#include "common/block_ops.h"
#include "common/config.h"
#include "common/layers.h"
#include "common/plan.h"
#include "common/sra_sim.h"
#include "kernels/launch.h"
#include "pg/comm_hook.h"
#include "pg/process_group_cgx.h"

namespace py = pybind11;
using namespace cgx;

namespace {

using LayerTuple = std::tuple<int64_t, int64_t, int64_t, int64_t>;  // (elem_off, numel, bits, bucket)

std::vector<LayerSpec> to_layers(const std::vector<LayerTuple>& in) {
  std::vector<LayerSpec> out;
  out.reserve(in.size());
  for (const auto& t : in) {
    LayerSpec l;
    l.elem_off = (uint64_t)std::get<0>(t);
    l.numel = (uint64_t)std::get<1>(t);
    l.bits = (int)std::get<2>(t);
    l.bucket = (uint32_t)std::get<3>(t);
    out.push_back(l);
  }
  return out;
}

int cgx_dtype(const at::Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return kF32;
    case at::kHalf: return kF16;
    case at::kBFloat16: return kBF16;
    default: TORCH_CHECK(false, "cgx: unsupported dtype ", t.scalar_type(), " (need float32/float16/bfloat16)");
  }
}

PlanOptions make_opts(int world, int lanes, int dtype, bool skip_incomplete, int64_t min_lane_elems) {
  PlanOptions o;
  o.world = world;
  o.lanes = lanes;
  o.dtype = dtype;
  o.skip_incomplete = skip_incomplete;
  o.min_lane_elems = (uint32_t)min_lane_elems;
  return o;
}

py::dict plan_to_dict(const Plan& p) {
  py::dict d;
  d["world"] = p.world;
  d["lanes"] = p.lanes;
  d["numel"] = p.numel;
  d["max_chunk_wire"] = p.max_chunk_wire;
  d["total_wire"] = p.total_wire;
  auto blocks = at::empty({(int64_t)p.blocks.size(), 5}, at::kLong);
  auto* bp = blocks.data_ptr<int64_t>();
  for (size_t i = 0; i < p.blocks.size(); ++i) {
    bp[i * 5 + 0] = p.blocks[i].elem_off;
    bp[i * 5 + 1] = block_n(p.blocks[i]);
    bp[i * 5 + 2] = block_bits(p.blocks[i]);
    bp[i * 5 + 3] = p.blocks[i].bucket;
    bp[i * 5 + 4] = p.blocks[i].wire_off;
  }
  d["blocks"] = blocks;
  auto items = at::empty({(int64_t)p.items.size(), 7}, at::kLong);
  auto* ip = items.data_ptr<int64_t>();
  for (size_t i = 0; i < p.items.size(); ++i) {
    ip[i * 7 + 0] = p.items[i].elem_off;
    ip[i * 7 + 1] = p.items[i].meta_off;
    ip[i * 7 + 2] = p.items[i].pay_off;
    ip[i * 7 + 3] = item_kind(p.items[i]);
    ip[i * 7 + 4] = item_lpb_log2(p.items[i]);
    ip[i * 7 + 5] = item_bits(p.items[i]);
    ip[i * 7 + 6] = item_n(p.items[i]);
  }
  d["items"] = items;  // rows: elem_off, meta_off, pay_off, kind, log2(bucket/8), bits, n
  d["item_first"] = std::vector<int64_t>(p.item_first.begin(), p.item_first.end());
  d["block_item_first"] = std::vector<int64_t>(p.block_item_first.begin(), p.block_item_first.end());
  d["slice_elems"] = p.slice_elems;
  d["uniform_bits"] = p.uniform_bits;
  d["lane_first"] = std::vector<int64_t>(p.lane_first.begin(), p.lane_first.end());
  d["chunk_wire_bytes"] = std::vector<int64_t>(p.chunk_wire_bytes.begin(), p.chunk_wire_bytes.end());
  d["chunk_elems"] = std::vector<int64_t>(p.chunk_elems.begin(), p.chunk_elems.end());
  return d;
}

RngParams make_rng(bool stochastic, uint64_t seed, uint32_t seq) {
  RngParams r;
  r.stochastic = stochastic;
  r.seed = seed;
  r.seq = seq;
  return r;
}

size_t wire_row_bytes(const Plan& p) { return ((size_t)p.max_chunk_wire + 255) / 256 * 256; }

// ---- CPU oracle ops -------------------------------------------------------
void py_sra_simulate(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers, int lanes,
                     bool average, bool skip_incomplete, bool stochastic, uint64_t seed, uint32_t seq,
                     int64_t min_lane_elems) {
  TORCH_CHECK(!tensors.empty(), "need tensors");
  const int W = (int)tensors.size();
  const int dt = cgx_dtype(tensors[0]);
  std::vector<void*> bufs;
  for (auto& t : tensors) {
    TORCH_CHECK(t.device().is_cpu() && t.is_contiguous() && cgx_dtype(t) == dt, "sra_simulate: CPU contiguous tensors of one dtype");
    bufs.push_back(t.data_ptr());
  }
  Plan plan = build_plan(to_layers(layers), make_opts(W, lanes, dt, skip_incomplete, min_lane_elems));
  sra_simulate(plan, bufs, average ? 1.0f / (float)W : 1.0f, make_rng(stochastic, seed, seq));
}

void py_oneshot_simulate(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers, int lanes,
                         bool average, bool skip_incomplete, bool stochastic, uint64_t seed, uint32_t seq,
                         int64_t min_lane_elems) {
  TORCH_CHECK(!tensors.empty(), "need tensors");
  const int W = (int)tensors.size();
  const int dt = cgx_dtype(tensors[0]);
  std::vector<void*> bufs;
  for (auto& t : tensors) {
    TORCH_CHECK(t.device().is_cpu() && t.is_contiguous() && cgx_dtype(t) == dt, "oneshot_simulate: CPU contiguous tensors");
    bufs.push_back(t.data_ptr());
  }
  Plan plan = build_plan(to_layers(layers), make_opts(1, lanes, dt, skip_incomplete, min_lane_elems));
  oneshot_simulate(plan, bufs, average ? 1.0f / (float)W : 1.0f, make_rng(stochastic, seed, seq));
}

// quantize `t` through the plan; returns uint8 wire tensor [world, row_bytes]
at::Tensor py_quantize(const at::Tensor& t, const std::vector<LayerTuple>& layers, int world, int lanes,
                       bool skip_incomplete, float prescale, bool stochastic, uint64_t seed, uint32_t seq,
                       int rank, int phase, int64_t min_lane_elems) {
  TORCH_CHECK(t.is_contiguous(), "quantize: tensor must be contiguous");
  const int dt = cgx_dtype(t);
  Plan plan = build_plan(to_layers(layers), make_opts(world, lanes, dt, skip_incomplete, min_lane_elems));
  const size_t row = wire_row_bytes(plan);
  RngKey key = make_rng_key(make_rng(stochastic, seed, seq), rank, phase);
  if (t.device().is_cpu()) {
    at::Tensor wire = at::zeros({(int64_t)world, (int64_t)row}, at::kByte);
    std::vector<float> acc(kMaxBlockElems);
    for (int c = 0; c < world; ++c)
      for (uint32_t b = plan.chunk_begin(c); b < plan.chunk_end(c); ++b) {
        const BlockDesc& d = plan.blocks[b];
        cpu::load_block(t.data_ptr(), dt, d, prescale, acc.data());
        cpu::quantize_block(acc.data(), dt, d, wire.data_ptr<uint8_t>() + (size_t)c * row + d.wire_off, key, b);
      }
    return wire;
  }
  c10::cuda::CUDAGuard g(t.device());
  auto stream = c10::cuda::getCurrentCUDAStream();
  at::Tensor wire = at::zeros({(int64_t)world, (int64_t)row}, t.options().dtype(at::kByte));
  at::Tensor ditems = at::empty({(int64_t)(plan.items.size() * sizeof(WarpItem) + 16)}, t.options().dtype(at::kByte));
  cuda_check(cudaMemcpyAsync(ditems.data_ptr(), plan.items.data(), plan.items.size() * sizeof(WarpItem),
                             cudaMemcpyHostToDevice, stream),
             "upload items");
  for (int c = 0; c < world; ++c) {
    ItemKernelArgs a;
    a.items = (const WarpItem*)ditems.data_ptr();
    a.first = plan.item_first[(size_t)c * plan.lanes];
    a.count = plan.item_first[(size_t)(c + 1) * plan.lanes] - a.first;
    a.dtype = dt;
    a.slice_elems = (int)plan.slice_elems;
    a.uniform_bits = plan.uniform_bits;
    cuda_check(launch_quantize_items(a, t.data_ptr(), wire.data_ptr<uint8_t>() + (size_t)c * row, prescale, key, stream),
               "quantize_items");
  }
  return wire;
}

// inverse of py_quantize: decode `wire` into a tensor shaped like `like`
at::Tensor py_dequantize(const at::Tensor& wire, const at::Tensor& like, const std::vector<LayerTuple>& layers,
                         int world, int lanes, bool skip_incomplete, int64_t min_lane_elems) {
  const int dt = cgx_dtype(like);
  Plan plan = build_plan(to_layers(layers), make_opts(world, lanes, dt, skip_incomplete, min_lane_elems));
  const size_t row = wire_row_bytes(plan);
  TORCH_CHECK(wire.is_contiguous() && wire.scalar_type() == at::kByte && (size_t)wire.numel() == row * (size_t)world,
              "dequantize: wire tensor has the wrong shape");
  at::Tensor out = at::zeros_like(like);
  if (like.device().is_cpu()) {
    for (int c = 0; c < world; ++c)
      for (uint32_t b = plan.chunk_begin(c); b < plan.chunk_end(c); ++b) {
        const BlockDesc& d = plan.blocks[b];
        cpu::decode_block_store(wire.data_ptr<uint8_t>() + (size_t)c * row + d.wire_off, dt, d, out.data_ptr());
      }
    return out;
  }
  c10::cuda::CUDAGuard g(like.device());
  auto stream = c10::cuda::getCurrentCUDAStream();
  at::Tensor ditems = at::empty({(int64_t)(plan.items.size() * sizeof(WarpItem) + 16)}, like.options().dtype(at::kByte));
  cuda_check(cudaMemcpyAsync(ditems.data_ptr(), plan.items.data(), plan.items.size() * sizeof(WarpItem),
                             cudaMemcpyHostToDevice, stream),
             "upload items");
  for (int c = 0; c < world; ++c) {
    ItemKernelArgs a;
    a.items = (const WarpItem*)ditems.data_ptr();
    a.first = plan.item_first[(size_t)c * plan.lanes];
    a.count = plan.item_first[(size_t)(c + 1) * plan.lanes] - a.first;
    a.dtype = dt;
    a.slice_elems = (int)plan.slice_elems;
    a.uniform_bits = plan.uniform_bits;
    cuda_check(launch_dequantize_items(a, wire.data_ptr<uint8_t>() + (size_t)c * row, out.data_ptr(), stream),
               "dequantize_items");
  }
  return out;
}

// ---- pre-planned codec: the kernels alone, no per-call planning or allocation ----------------
// (kernel microbenchmarks, and users who quantize the same layout repeatedly)
class PreparedCodec {
 public:
  PreparedCodec(const std::vector<LayerTuple>& layers, at::ScalarType dtype, int64_t device, bool skip_incomplete) {
    at::Tensor probe = at::empty({0}, at::TensorOptions().dtype(dtype));
    dt_ = cgx_dtype(probe);
    plan_ = build_plan(to_layers(layers), make_opts(1, 1, dt_, skip_incomplete, 1 << 30));
    c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
    items_ = at::empty({(int64_t)(plan_.items.size() * sizeof(WarpItem) + 16)},
                       at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device));
    cuda_check(cudaMemcpy(items_.data_ptr(), plan_.items.data(), plan_.items.size() * sizeof(WarpItem),
                          cudaMemcpyHostToDevice),
               "upload items");
  }
  int64_t wire_bytes() const { return (int64_t)wire_row_bytes(plan_); }
  int64_t numel() const { return (int64_t)plan_.numel; }
  int64_t num_items() const { return (int64_t)plan_.items.size(); }
  void quantize(const at::Tensor& x, at::Tensor& wire, double prescale, bool stochastic, uint64_t seed, uint32_t seq) {
    check(x, wire);
    c10::cuda::CUDAGuard g(x.device());
    cuda_check(launch_quantize_items(args(), x.data_ptr(), wire.data_ptr<uint8_t>(), (float)prescale,
                                     make_rng_key(make_rng(stochastic, seed, seq), 0, 0),
                                     c10::cuda::getCurrentCUDAStream()),
               "quantize_items");
  }
  void dequantize(const at::Tensor& wire, at::Tensor& out) {
    check(out, wire);
    c10::cuda::CUDAGuard g(out.device());
    cuda_check(launch_dequantize_items(args(), wire.data_ptr<uint8_t>(), out.data_ptr(),
                                       c10::cuda::getCurrentCUDAStream()),
               "dequantize_items");
  }

 private:
  void check(const at::Tensor& x, const at::Tensor& wire) const {
    TORCH_CHECK(x.is_cuda() && x.is_contiguous() && cgx_dtype(x) == dt_ && (uint64_t)x.numel() == plan_.numel,
                "PreparedCodec: tensor does not match the prepared layout");
    TORCH_CHECK(wire.is_cuda() && wire.is_contiguous() && wire.scalar_type() == at::kByte &&
                    (size_t)wire.numel() >= wire_row_bytes(plan_),
                "PreparedCodec: wire buffer too small");
  }
  ItemKernelArgs args() const {
    ItemKernelArgs a;
    a.items = (const WarpItem*)items_.data_ptr();
    a.first = 0;
    a.count = (uint32_t)plan_.items.size();
    a.dtype = dt_;
    a.slice_elems = (int)plan_.slice_elems;
    a.uniform_bits = plan_.uniform_bits;
    return a;
  }
  Plan plan_;
  int dt_ = kF32;
  at::Tensor items_;
};

// ---- single-process multi-rank harness --------------------------------------
class LocalSraGroup {
 public:
  LocalSraGroup(int world, int lanes, int64_t slot_bytes, int64_t timeout_ms, int64_t min_lane_elems)
      : world_(world) {
    TORCH_CHECK(world >= 1 && world <= kMaxPeers, "bad world size");
    HeapLayout layout = HeapLayout::make(world, lanes, (size_t)slot_bytes, (size_t)slot_bytes);
    for (int r = 0; r < world; ++r) heaps_.push_back(std::make_unique<SymmetricHeap>(r, world, layout));
    std::vector<SymmetricHeap*> all;
    for (auto& h : heaps_) all.push_back(h.get());
    for (auto& h : heaps_) h->connect_local(all);
    for (int r = 0; r < world; ++r) {
      fused_.push_back(std::make_unique<FusedSra>(heaps_[r].get(), lanes, timeout_ms, (uint32_t)min_lane_elems));
      streams_.push_back(c10::cuda::getStreamFromPool(false));
    }
  }

  // in-place allreduce of W same-device tensors, one per virtual rank
  void allreduce_oneshot(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers, bool average,
                         bool skip_incomplete, bool stochastic, uint64_t seed, uint32_t seq) {
    run(std::move(tensors), layers, average, skip_incomplete, stochastic, seed, seq, true);
  }
  void allreduce(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers, bool average,
                 bool skip_incomplete, bool stochastic, uint64_t seed, uint32_t seq) {
    run(std::move(tensors), layers, average, skip_incomplete, stochastic, seed, seq, false);
  }
  // fault injection: virtual rank `absent` never launches its kernel (a dead / stuck peer)
  void allreduce_with_absent_rank(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers,
                                  int absent) {
    run(std::move(tensors), layers, false, false, false, 0, 1, false, absent);
  }
  void abort_all() {
    for (auto& h : heaps_) h->request_abort(true);
  }
  std::string status(int r) { return fused_[r]->status_message(); }
  void clear() {
    for (auto& f : fused_) f->clear_status();
  }
  void run(std::vector<at::Tensor> tensors, const std::vector<LayerTuple>& layers, bool average,
           bool skip_incomplete, bool stochastic, uint64_t seed, uint32_t seq, bool oneshot, int absent = -1) {
    TORCH_CHECK((int)tensors.size() == world_, "need one tensor per virtual rank");
    const int dt = cgx_dtype(tensors[0]);
    auto specs = to_layers(layers);
    auto cur = c10::cuda::getCurrentCUDAStream();
    at::cuda::CUDAEvent ready(cudaEventDisableTiming);
    ready.record(cur);
    const RngParams rng = make_rng(stochastic, seed, seq);
    std::vector<const DevicePlan*> plans(world_);
    bool fresh_plan = false;
    for (int r = 0; r < world_; ++r) {
      TORCH_CHECK(tensors[r].is_cuda() && tensors[r].is_contiguous() && cgx_dtype(tensors[r]) == dt, "bad tensor");
      ready.block(streams_[r]);
      const size_t before = fused_[r]->num_plans();
      plans[r] = oneshot ? fused_[r]->prepare_oneshot(specs, dt, skip_incomplete, streams_[r].stream())
                         : fused_[r]->prepare(specs, dt, skip_incomplete, streams_[r].stream());
      TORCH_CHECK(plans[r] != nullptr, "plan does not fit the heap slots (raise slot_bytes)");
      fresh_plan = fresh_plan || fused_[r]->num_plans() != before;
    }
    // make sure every plan upload has landed before any virtual rank starts spinning
    if (fresh_plan)
      for (int r = 0; r < world_; ++r) streams_[r].synchronize();
    for (int r = 0; r < world_; ++r) {
      if (r == absent) continue;
      const float ps = average ? 1.0f / (float)world_ : 1.0f;
      if (oneshot)
        fused_[r]->run_oneshot(*plans[r], tensors[r].data_ptr(), ps, rng, streams_[r].stream());
      else
        fused_[r]->run(*plans[r], tensors[r].data_ptr(), ps, rng, streams_[r].stream());
    }
    for (int r = 0; r < world_; ++r) {
      at::cuda::CUDAEvent done(cudaEventDisableTiming);
      done.record(streams_[r]);
      done.block(cur);
    }
  }

  void check() {
    for (auto& f : fused_) f->check_status();
  }
  int64_t lanes_used(const std::vector<LayerTuple>& layers, const at::Tensor& like, bool skip_incomplete) {
    auto* dp = fused_[0]->prepare(to_layers(layers), cgx_dtype(like), skip_incomplete,
                                  c10::cuda::getCurrentCUDAStream().stream());
    return dp ? dp->plan.lanes : -1;
  }

 private:
  int world_;
  std::vector<std::unique_ptr<SymmetricHeap>> heaps_;
  std::vector<std::unique_ptr<FusedSra>> fused_;
  std::vector<c10::cuda::CUDAStream> streams_;
};

c10::intrusive_ptr<c10d::Backend> create_backend(const c10::intrusive_ptr<c10d::Store>& store, int rank, int size,
                                                 const std::chrono::milliseconds& timeout,
                                                 c10::intrusive_ptr<c10d::Backend> cpu_delegate,
                                                 c10::intrusive_ptr<c10d::Backend> cuda_delegate, int local_size,
                                                 c10::intrusive_ptr<c10d::Backend> cpu_local,
                                                 c10::intrusive_ptr<c10d::Backend> cpu_cross,
                                                 c10::intrusive_ptr<c10d::Backend> cuda_local,
                                                 c10::intrusive_ptr<c10d::Backend> cuda_cross) {
  ProcessGroupCGX::Topology topo;
  topo.local_size = local_size;
  topo.cpu_local = std::move(cpu_local);
  topo.cpu_cross = std::move(cpu_cross);
  topo.cuda_local = std::move(cuda_local);
  topo.cuda_cross = std::move(cuda_cross);
  return c10::make_intrusive<ProcessGroupCGX>(store, rank, size, timeout, std::move(cpu_delegate),
                                              std::move(cuda_delegate), std::move(topo));
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torch_cgx_b200 native core (sm_100a)";
  py::module::import("torch._C._distributed_c10d");

  py::class_<ProcessGroupCGX, c10d::Backend, c10::intrusive_ptr<ProcessGroupCGX>>(m, "ProcessGroupCGX")
      .def("allreduce_bucket", &ProcessGroupCGX::allreduce_bucket, py::arg("tensor"), py::arg("bucket_idx"),
           py::arg("average") = false, py::call_guard<py::gil_scoped_release>())
      .def("init_cuda", &ProcessGroupCGX::init_cuda, py::call_guard<py::gil_scoped_release>())
      .def("p2p_ready", &ProcessGroupCGX::p2p_ready)
      .def("lanes", &ProcessGroupCGX::lanes)
      .def("local_size", &ProcessGroupCGX::local_size)
      .def("stats", &ProcessGroupCGX::stats)
      .def("reset_stats", &ProcessGroupCGX::reset_stats)
      .def("check_health", &ProcessGroupCGX::check_health)
      .def("failure", &ProcessGroupCGX::failure)
      .def("last_lanes", &ProcessGroupCGX::last_lanes)
      .def("abort", &ProcessGroupCGX::abort, py::call_guard<py::gil_scoped_release>())
      .def("uses_multicast", [](ProcessGroupCGX& pg) { return pg.uses_multicast(); })
      .def("heap_kind", [](ProcessGroupCGX& pg) { return pg.heap_kind(); })
      .def("enable_trace", &ProcessGroupCGX::enable_trace)
      .def("read_trace", &ProcessGroupCGX::read_trace);

  py::class_<HookState, std::shared_ptr<HookState>>(m, "NativeHookState")
      .def_property_readonly("step", [](const HookState& h) { return h.step.load(); })
      .def_readonly("layer_min_size", &HookState::layer_min_size)
      .def_readonly("bits", &HookState::bits)
      .def_readonly("bucket_size", &HookState::bucket_size)
      .def_readonly("register_step", &HookState::register_step);
  m.def("register_native_hook",
        [](const std::shared_ptr<c10d::Reducer>& reducer, const c10::intrusive_ptr<ProcessGroupCGX>& pg,
           int64_t layer_min_size, int bits, int bucket_size, int register_step) {
          return register_native_hook(reducer, pg, layer_min_size, bits, bucket_size, register_step);
        },
        py::arg("reducer"), py::arg("backend"), py::arg("layer_min_size"), py::arg("bits"), py::arg("bucket_size"),
        py::arg("register_step") = 2);

  m.def("create_backend", &create_backend, py::arg("store"), py::arg("rank"), py::arg("size"), py::arg("timeout"),
        py::arg("cpu_delegate"), py::arg("cuda_delegate"), py::arg("local_size") = 0,
        py::arg("cpu_local") = nullptr, py::arg("cpu_cross") = nullptr, py::arg("cuda_local") = nullptr,
        py::arg("cuda_cross") = nullptr, py::call_guard<py::gil_scoped_release>());

  // reference pybind surface
  m.def("register_layer", [](unsigned bucket_idx, unsigned layer_idx, int64_t numel, int bits, int bucket_size) {
    LayerRegistry::instance().register_layer(bucket_idx, layer_idx, numel, bits, bucket_size);
  });
  m.def("set_quantization_bits", [](unsigned bucket_idx, unsigned layer_idx, int bits) {
    LayerRegistry::instance().set_bits(bucket_idx, layer_idx, bits);
  });
  m.def("set_quantization_bucket_size", [](unsigned bucket_idx, unsigned layer_idx, int bucket_size) {
    LayerRegistry::instance().set_bucket_size(bucket_idx, layer_idx, bucket_size);
  });
  m.def("reset_layers", []() { LayerRegistry::instance().reset(); });
  m.def("num_registered_buckets", []() { return (int64_t)LayerRegistry::instance().num_buckets(); });
  m.def("registered_bucket", [](unsigned bucket_idx) {
    std::vector<std::tuple<int64_t, int, int>> out;
    auto sizes = LayerRegistry::instance().bucket_layer_sizes(bucket_idx);
    for (size_t i = 0; i < sizes.size(); ++i) {
      auto c = LayerRegistry::instance().layer_config(bucket_idx, (unsigned)i);
      out.emplace_back(sizes[i], c.bits, c.bucket_size);
    }
    return out;
  });
  m.def("extract_layers", [](int64_t numel, int explicit_bucket) {
    CompressionEnv env = CompressionEnv::read();
    EngineConfig cfg = EngineConfig::read();
    int resolved = -1;
    auto ls = LayerRegistry::instance().extract(numel, env, cfg.min_compress_elems, explicit_bucket, &resolved);
    std::vector<LayerTuple> out;
    for (auto& l : ls) out.emplace_back((int64_t)l.elem_off, (int64_t)l.numel, (int64_t)l.bits, (int64_t)l.bucket);
    return std::make_pair(out, resolved);
  }, py::arg("numel"), py::arg("explicit_bucket") = -1);

  m.def("engine_config", []() {
    EngineConfig c = EngineConfig::read();
    CompressionEnv e = CompressionEnv::read();
    py::dict d;
    d["fusion_bytes"] = c.fusion_bytes;
    d["min_compress_elems"] = c.min_compress_elems;
    d["fake_ratio"] = c.fake_ratio;
    d["inner_comm"] = to_string(c.inner_comm);
    d["cross_comm"] = to_string(c.cross_comm);
    d["inner_reduction"] = to_string(c.inner_reduction);
    d["cross_reduction"] = to_string(c.cross_reduction);
    d["intra_broadcast"] = c.intra_broadcast;
    d["intra_compress"] = c.intra_compress;
    d["dummy_compression"] = c.dummy_compression;
    d["remote_buf"] = c.remote_buf;
    d["oneshot_max_bytes"] = c.oneshot_max_bytes;
    d["lanes"] = c.lanes;
    d["timeout_ms"] = c.timeout_ms;
    d["local_size"] = c.local_size;
    d["bits"] = e.bits;
    d["bucket_size"] = e.bucket_size;
    d["skip_incomplete"] = e.skip_incomplete;
    d["stochastic"] = e.stochastic;
    d["seed"] = e.seed;
    return d;
  });

  m.def("build_plan", [](const std::vector<LayerTuple>& layers, int world, int lanes, int dtype, bool skip_incomplete,
                         int64_t min_lane_elems) {
    return plan_to_dict(build_plan(to_layers(layers), make_opts(world, lanes, dtype, skip_incomplete, min_lane_elems)));
  }, py::arg("layers"), py::arg("world"), py::arg("lanes"), py::arg("dtype") = 0, py::arg("skip_incomplete") = false,
        py::arg("min_lane_elems") = 2048);
  m.def("split_for_fusion", [](const std::vector<LayerTuple>& layers, int elsize, int64_t fusion_bytes) {
    std::vector<std::vector<LayerTuple>> out;
    for (auto& g : split_for_fusion(to_layers(layers), elsize, fusion_bytes)) {
      std::vector<LayerTuple> o;
      for (auto& l : g) o.emplace_back((int64_t)l.elem_off, (int64_t)l.numel, (int64_t)l.bits, (int64_t)l.bucket);
      out.push_back(std::move(o));
    }
    return out;
  });

  m.def("sra_simulate", &py_sra_simulate, py::arg("tensors"), py::arg("layers"), py::arg("lanes") = 1,
        py::arg("average") = false, py::arg("skip_incomplete") = false, py::arg("stochastic") = false,
        py::arg("seed") = 0, py::arg("seq") = 0, py::arg("min_lane_elems") = 2048);
  m.def("oneshot_simulate", &py_oneshot_simulate, py::arg("tensors"), py::arg("layers"), py::arg("lanes") = 1,
        py::arg("average") = false, py::arg("skip_incomplete") = false, py::arg("stochastic") = false,
        py::arg("seed") = 0, py::arg("seq") = 0, py::arg("min_lane_elems") = 2048);
  m.def("quantize", &py_quantize, py::arg("tensor"), py::arg("layers"), py::arg("world") = 1, py::arg("lanes") = 1,
        py::arg("skip_incomplete") = false, py::arg("prescale") = 1.0f, py::arg("stochastic") = false,
        py::arg("seed") = 0, py::arg("seq") = 0, py::arg("rank") = 0, py::arg("phase") = 0,
        py::arg("min_lane_elems") = 2048);
  m.def("dequantize", &py_dequantize, py::arg("wire"), py::arg("like"), py::arg("layers"), py::arg("world") = 1,
        py::arg("lanes") = 1, py::arg("skip_incomplete") = false, py::arg("min_lane_elems") = 2048);

  m.def("scale_", [](at::Tensor t, double s) {
    TORCH_CHECK(t.is_cuda() && t.is_contiguous(), "scale_: contiguous CUDA tensor");
    c10::cuda::CUDAGuard g(t.device());
    cuda_check(launch_scale_inplace(t.data_ptr(), cgx_dtype(t), (uint64_t)t.numel(), (float)s,
                                    c10::cuda::getCurrentCUDAStream()),
               "scale");
    return t;
  });
  m.def("add", [](const at::Tensor& x, const at::Tensor& y) {
    TORCH_CHECK(x.is_cuda() && y.is_cuda() && x.is_contiguous() && y.is_contiguous() && x.numel() == y.numel() &&
                    x.scalar_type() == y.scalar_type(),
                "add: matching contiguous CUDA tensors");
    c10::cuda::CUDAGuard g(x.device());
    at::Tensor out = at::empty_like(x);
    cuda_check(launch_add(x.data_ptr(), y.data_ptr(), out.data_ptr(), cgx_dtype(x), (uint64_t)x.numel(),
                          c10::cuda::getCurrentCUDAStream()),
               "add");
    return out;
  });
  m.def("convert", [](const at::Tensor& x, at::ScalarType dst) {
    TORCH_CHECK(x.is_cuda() && x.is_contiguous(), "convert: contiguous CUDA tensor");
    c10::cuda::CUDAGuard g(x.device());
    at::Tensor out = at::empty_like(x, x.options().dtype(dst));
    cuda_check(launch_convert(x.data_ptr(), cgx_dtype(x), out.data_ptr(), cgx_dtype(out), (uint64_t)x.numel(),
                              c10::cuda::getCurrentCUDAStream()),
               "convert");
    return out;
  });
  m.def("max_resident_ctas", [](int dtype) { return sra_max_resident_ctas(dtype); }, py::arg("dtype") = 0);

  py::class_<FdChannel>(m, "FdChannel")
      .def(py::init<const std::string&>(), py::arg("hint"))
      .def("name", &FdChannel::name)
      .def("send", &FdChannel::try_send, py::arg("peer_name"), py::arg("fd"), py::arg("kind"), py::arg("src"))
      .def("recv", [](FdChannel& ch, int timeout_ms) -> py::object {
        FdMessage m;
        bool ok;
        {
          py::gil_scoped_release rel;
          ok = ch.try_recv(&m, timeout_ms);
        }
        if (!ok) return py::none();
        return py::make_tuple(m.fd, m.kind, m.src);
      }, py::arg("timeout_ms"));
  py::class_<PreparedCodec>(m, "PreparedCodec")
      .def(py::init<const std::vector<LayerTuple>&, at::ScalarType, int64_t, bool>(), py::arg("layers"),
           py::arg("dtype"), py::arg("device") = 0, py::arg("skip_incomplete") = false)
      .def("wire_bytes", &PreparedCodec::wire_bytes)
      .def("numel", &PreparedCodec::numel)
      .def("num_items", &PreparedCodec::num_items)
      .def("quantize", &PreparedCodec::quantize, py::arg("x"), py::arg("wire"), py::arg("prescale") = 1.0,
           py::arg("stochastic") = false, py::arg("seed") = 0, py::arg("seq") = 0)
      .def("dequantize", &PreparedCodec::dequantize, py::arg("wire"), py::arg("out"));
  py::class_<LocalSraGroup>(m, "LocalSraGroup")
      .def(py::init<int, int, int64_t, int64_t, int64_t>(), py::arg("world"), py::arg("lanes"),
           py::arg("slot_bytes"), py::arg("timeout_ms") = 5000, py::arg("min_lane_elems") = 2048)
      .def("allreduce", &LocalSraGroup::allreduce, py::arg("tensors"), py::arg("layers"), py::arg("average") = false,
           py::arg("skip_incomplete") = false, py::arg("stochastic") = false, py::arg("seed") = 0,
           py::arg("seq") = 0)
      .def("allreduce_oneshot", &LocalSraGroup::allreduce_oneshot, py::arg("tensors"), py::arg("layers"),
           py::arg("average") = false, py::arg("skip_incomplete") = false, py::arg("stochastic") = false,
           py::arg("seed") = 0, py::arg("seq") = 0)
      .def("check", &LocalSraGroup::check)
      .def("allreduce_with_absent_rank", &LocalSraGroup::allreduce_with_absent_rank, py::arg("tensors"),
           py::arg("layers"), py::arg("absent"))
      .def("abort_all", &LocalSraGroup::abort_all)
      .def("status", &LocalSraGroup::status)
      .def("clear", &LocalSraGroup::clear)
      .def("lanes_used", &LocalSraGroup::lanes_used);

  m.attr("MAX_BLOCK_ELEMS") = (int64_t)kMaxBlockElems;
  m.attr("MAX_PEERS") = (int64_t)kMaxPeers;
  m.attr("RAW_BITS") = (int64_t)kRawBits;
}
