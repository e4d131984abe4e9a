#include "reducers.h"

#include <algorithm>
#include <stdexcept>

namespace cgx {

Reducer::Reducer(Communicator* comm, BlockBackend* ops) : comm_(comm), ops_(ops) {}

Reducer::~Reducer() {
  for (Buf& b : bufs_)
    if (b.p) ops_->release(b.p);
}

const Plan& Reducer::plan_for(const std::vector<LayerSpec>& layers, int dtype, bool skip_incomplete, int world) {
  PlanOptions opt;
  opt.world = world > 0 ? world : comm_->size();
  opt.lanes = 1;  // rank chunks only: one kernel launch per chunk handles all of its blocks
  opt.dtype = dtype;
  opt.skip_incomplete = skip_incomplete;
  const uint64_t key = plan_key(layers, opt);
  auto it = plans_.find(key);
  if (it == plans_.end()) it = plans_.emplace(key, std::make_unique<Plan>(build_plan(layers, opt))).first;
  return *it->second;
}

uint8_t* Reducer::scratch(int slot, size_t bytes) {
  if ((int)bufs_.size() <= slot) bufs_.resize(slot + 1);
  Buf& b = bufs_[slot];
  if (b.cap < bytes) {
    if (b.p) ops_->release(b.p);
    b.cap = std::max(bytes, b.cap * 2);
    b.p = static_cast<uint8_t*>(ops_->alloc(b.cap));
  }
  return b.p;
}

void Reducer::chunk_span(const Plan& p, int chunk, uint32_t* lo, uint32_t* hi) {
  uint32_t a = ~0u, b = 0;
  const uint32_t first = chunk < 0 ? 0 : p.chunk_begin(chunk);
  const uint32_t last = chunk < 0 ? (uint32_t)p.blocks.size() : p.chunk_end(chunk);
  for (uint32_t i = first; i < last; ++i) {
    a = std::min(a, p.blocks[i].elem_off);
    b = std::max(b, p.blocks[i].elem_off + block_n(p.blocks[i]));
  }
  if (a == ~0u) a = b = 0;
  *lo = a;
  *hi = b;
}

void Reducer::broadcast(void* data, size_t bytes, int root, cudaStream_t stream) {
  const int W = comm_->size(), r = comm_->rank();
  if (W == 1 || bytes == 0) return;
  std::vector<P2POp> ops;
  if (r == root) {
    for (int p = 0; p < W; ++p)
      if (p != root) ops.push_back({true, data, bytes, p});
    bytes_sent_ += bytes * (uint64_t)(W - 1);
  } else {
    ops.push_back({false, data, bytes, root});
  }
  comm_->exchange(ops, stream);
}

void Reducer::broadcast_compressed(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                                   int root, const RngParams& rng, cudaStream_t stream) {
  const int W = comm_->size(), r = comm_->rank();
  if (W == 1) return;
  const Plan& plan = plan_for(layers, dtype, skip_incomplete, 1);  // the whole buffer is one chunk
  if (plan.blocks.empty()) return;
  ops_->bind(plan, stream);
  const uint32_t nb = (uint32_t)plan.blocks.size();
  const size_t bytes = plan.chunk_wire_bytes[0];
  uint8_t* wire = scratch(4, row_bytes(plan));
  std::vector<P2POp> x;
  if (r == root) {
    ops_->quantize(data, 0, nb, wire, 1.0f, make_rng_key(rng, root, 2), stream);
    ops_->dequantize(wire, 0, nb, data, stream);  // the root keeps what the receivers will decode
    for (int p = 0; p < W; ++p)
      if (p != root) x.push_back({true, wire, bytes, p});
    bytes_sent_ += bytes * (uint64_t)(W - 1);
    comm_->exchange(x, stream);
  } else {
    x.push_back({false, wire, bytes, root});
    comm_->exchange(x, stream);
    ops_->dequantize(wire, 0, nb, data, stream);
  }
}

// ---------------------------------------------------------------------- SRA --
void SraReducer::allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                           float prescale, const RngParams& rng, cudaStream_t stream) {
  const Plan& plan = plan_for(layers, dtype, skip_incomplete);
  if (plan.blocks.empty()) return;
  const int W = comm_->size(), r = comm_->rank();
  ops_->bind(plan, stream);
  const size_t row = row_bytes(plan);
  uint8_t* sendb = scratch(0, row * W);
  uint8_t* recvb = scratch(1, row * W);
  uint8_t* mine = scratch(2, row);
  uint32_t lo, hi;
  chunk_span(plan, r, &lo, &hi);
  float* acc = reinterpret_cast<float*>(scratch(3, (size_t)(hi - lo) * sizeof(float) + 16));
  const uint32_t cb = plan.chunk_begin(r), cn = plan.chunk_end(r) - cb;

  // phase 1: my quantized copy of every peer's chunk -> that peer
  std::vector<P2POp> x;
  for (int p = 0; p < W; ++p) {
    if (p == r) continue;
    const uint32_t pb = plan.chunk_begin(p), pn = plan.chunk_end(p) - pb;
    if (pn) ops_->quantize(data, pb, pn, sendb + row * p, prescale, make_rng_key(rng, r, 0), stream);
    if (plan.chunk_wire_bytes[p]) x.push_back({true, sendb + row * p, plan.chunk_wire_bytes[p], p});
    if (plan.chunk_wire_bytes[r]) x.push_back({false, recvb + row * p, plan.chunk_wire_bytes[r], p});
    bytes_sent_ += plan.chunk_wire_bytes[p];
  }
  comm_->exchange(x, stream);
  // reduce my chunk in fp32: own raw values first, then the peers in rank order
  if (cn) {
    ops_->accumulate(nullptr, cb, cn, acc, lo, data, prescale, stream);
    for (int q = 0; q < W; ++q)
      if (q != r) ops_->accumulate(recvb + row * q, cb, cn, acc, lo, nullptr, 1.0f, stream);
    // requantize + self-decode so that every rank ends with identical values
    ops_->quantize_f32(acc, lo, cb, cn, mine, make_rng_key(rng, r, 1), data, stream);
  }
  // phase 2: all-gather of the reduced chunks
  x.clear();
  for (int p = 0; p < W; ++p) {
    if (p == r) continue;
    if (plan.chunk_wire_bytes[r]) x.push_back({true, mine, plan.chunk_wire_bytes[r], p});
    if (plan.chunk_wire_bytes[p]) x.push_back({false, recvb + row * p, plan.chunk_wire_bytes[p], p});
    bytes_sent_ += plan.chunk_wire_bytes[r];
  }
  comm_->exchange(x, stream);
  for (int p = 0; p < W; ++p) {
    if (p == r) continue;
    const uint32_t pb = plan.chunk_begin(p), pn = plan.chunk_end(p) - pb;
    if (pn) ops_->dequantize(recvb + row * p, pb, pn, data, stream);
  }
}

// --------------------------------------------------------------------- Ring --
void RingReducer::allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                            float prescale, const RngParams& rng, cudaStream_t stream) {
  const Plan& plan = plan_for(layers, dtype, skip_incomplete);
  if (plan.blocks.empty()) return;
  const int W = comm_->size(), r = comm_->rank();
  ops_->bind(plan, stream);
  const size_t row = row_bytes(plan);
  uint8_t* bufa = scratch(0, row);
  uint8_t* bufb = scratch(1, row);
  uint32_t lo, hi;
  chunk_span(plan, -1, &lo, &hi);
  float* acc = reinterpret_cast<float*>(scratch(3, (size_t)(hi - lo) * sizeof(float) + 16));
  const uint32_t nblocks = (uint32_t)plan.blocks.size();
  auto cb = [&](int c) { return plan.chunk_begin(c); };
  auto cn = [&](int c) { return plan.chunk_end(c) - plan.chunk_begin(c); };
  ops_->accumulate(nullptr, 0, nblocks, acc, lo, data, prescale, stream);
  const int right = (r + 1) % W, left = (r + W - 1) % W;
  // reduce-scatter: W-1 hops, the travelling partial sum is requantized at every hop
  for (int s = 0; s + 1 < W; ++s) {
    const int sseg = (r - s + W) % W, rseg = (r - s - 1 + 2 * W) % W;
    RngParams hop = rng;
    hop.seq = rng.seq * 16u + (uint32_t)s;
    if (cn(sseg)) ops_->quantize_f32(acc, lo, cb(sseg), cn(sseg), bufa, make_rng_key(hop, r, 0), nullptr, stream);
    std::vector<P2POp> x;
    if (plan.chunk_wire_bytes[sseg]) x.push_back({true, bufa, plan.chunk_wire_bytes[sseg], right});
    if (plan.chunk_wire_bytes[rseg]) x.push_back({false, bufb, plan.chunk_wire_bytes[rseg], left});
    bytes_sent_ += plan.chunk_wire_bytes[sseg];
    comm_->exchange(x, stream);
    if (cn(rseg)) ops_->accumulate(bufb, cb(rseg), cn(rseg), acc, lo, nullptr, 1.0f, stream);
  }
  // the fully reduced segment of this rank: requantize once more and self-decode
  int cur = (r + 1) % W;
  if (W == 1) cur = 0;
  RngParams fin = rng;
  fin.seq = rng.seq * 16u + 15u;
  uint8_t* have = bufa;
  uint8_t* next = bufb;
  if (cn(cur)) ops_->quantize_f32(acc, lo, cb(cur), cn(cur), have, make_rng_key(fin, r, 1), data, stream);
  // all-gather: forward the already-compressed bytes around the ring, decode locally
  for (int s = 0; s + 1 < W; ++s) {
    const int rseg = (cur + W - 1) % W;
    std::vector<P2POp> x;
    if (plan.chunk_wire_bytes[cur]) x.push_back({true, have, plan.chunk_wire_bytes[cur], right});
    if (plan.chunk_wire_bytes[rseg]) x.push_back({false, next, plan.chunk_wire_bytes[rseg], left});
    bytes_sent_ += plan.chunk_wire_bytes[cur];
    comm_->exchange(x, stream);
    if (cn(rseg)) ops_->dequantize(next, cb(rseg), cn(rseg), data, stream);
    std::swap(have, next);
    cur = rseg;
  }
}

// ----------------------------------------------------------------- AllToAll --
void AllToAllReducer::allreduce(void* data, int dtype, const std::vector<LayerSpec>& layers, bool skip_incomplete,
                                float prescale, const RngParams& rng, cudaStream_t stream) {
  const Plan& plan = plan_for(layers, dtype, skip_incomplete);
  if (plan.blocks.empty()) return;
  const int W = comm_->size(), r = comm_->rank();
  ops_->bind(plan, stream);
  const size_t row = row_bytes(plan);
  const size_t whole = row * W;  // chunk c of the wire image lives at row * c
  uint8_t* sendb = scratch(0, whole);
  uint8_t* recvb = scratch(1, whole * W);
  uint32_t lo, hi;
  chunk_span(plan, -1, &lo, &hi);
  float* acc = reinterpret_cast<float*>(scratch(3, (size_t)(hi - lo) * sizeof(float) + 16));
  for (int c = 0; c < W; ++c) {
    const uint32_t b0 = plan.chunk_begin(c), bn = plan.chunk_end(c) - b0;
    if (bn) ops_->quantize(data, b0, bn, sendb + row * c, prescale, make_rng_key(rng, r, 0), stream);
  }
  std::vector<P2POp> x;
  for (int p = 0; p < W; ++p) {
    if (p == r) continue;
    x.push_back({true, sendb, whole, p});
    x.push_back({false, recvb + whole * p, whole, p});
    bytes_sent_ += whole;
  }
  comm_->exchange(x, stream);
  ops_->accumulate(nullptr, 0, (uint32_t)plan.blocks.size(), acc, lo, data, prescale, stream);
  for (int q = 0; q < W; ++q) {
    if (q == r) continue;
    for (int c = 0; c < W; ++c) {
      const uint32_t b0 = plan.chunk_begin(c), bn = plan.chunk_end(c) - b0;
      if (bn) ops_->accumulate(recvb + whole * q + row * c, b0, bn, acc, lo, nullptr, 1.0f, stream);
    }
  }
  // store the fp32 sums back as T: a "raw" plan of the same layers does exactly that
  std::vector<LayerSpec> raw = layers;
  for (LayerSpec& l : raw) l.bits = kRawBits;
  const Plan& rp = plan_for(raw, dtype, false);
  ops_->bind(rp, stream);
  uint8_t* tmp = scratch(2, row_bytes(rp) * W);
  for (int c = 0; c < W; ++c) {
    const uint32_t b0 = rp.chunk_begin(c), bn = rp.chunk_end(c) - b0;
    if (bn) ops_->quantize_f32(acc, lo, b0, bn, tmp + row_bytes(rp) * c, make_rng_key(rng, r, 1), data, stream);
  }
}

std::unique_ptr<Reducer> make_reducer(ReductionType type, Communicator* comm, BlockBackend* ops) {
  switch (type) {
    case ReductionType::kRing: return std::make_unique<RingReducer>(comm, ops);
    case ReductionType::kAllToAll: return std::make_unique<AllToAllReducer>(comm, ops);
    default: return std::make_unique<SraReducer>(comm, ops);
  }
}

}  // namespace cgx
