#include "plan.h"

#include <algorithm>
#include <sstream>
#include <stdexcept>

namespace cgx {
namespace {

struct Segment {
  uint64_t elem_off;  // in the tensor
  uint64_t n;
  int bits;
  uint32_t bucket;      // snapping granularity == quantization bucket
  uint32_t block_elems; // multiple of bucket
  uint64_t start;       // position in the concatenated space
};

constexpr uint32_t kRawGranule = 512;

void push_segment(std::vector<Segment>& segs, uint64_t off, uint64_t n, int bits, uint32_t bucket,
                  uint32_t max_block_elems) {
  if (n == 0) return;
  Segment s;
  s.elem_off = off;
  s.n = n;
  s.bits = bits >= kRawBits ? kRawBits : bits;
  if (s.bits >= kRawBits) bucket = kRawGranule;
  bucket = std::max<uint32_t>(1u, std::min<uint32_t>(bucket, max_block_elems));
  s.bucket = bucket;
  uint32_t per_block = std::max<uint32_t>(1u, max_block_elems / bucket);
  per_block = std::min<uint32_t>(per_block, kMaxBlockBuckets);
  s.block_elems = per_block * bucket;
  s.start = 0;
  segs.push_back(s);
}

// Flatten one block into warp items (wire.h). Full slices need whole power-of-two buckets and a
// group-aligned start; everything else becomes one generic item per bucket.
void emit_items(const BlockDesc& d, int elsize, uint32_t slice, std::vector<WarpItem>& out) {
  const uint32_t n = block_n(d);
  const int bits = block_bits(d);
  if (bits >= kRawBits) {
    for (uint32_t off = 0; off < n; off += kRawItemElems) {
      const uint32_t cnt = std::min(kRawItemElems, n - off);
      WarpItem it;
      it.elem_off = d.elem_off + off;
      it.meta_off = it.pay_off = d.wire_off + off * (uint32_t)elsize;
      it.info = make_item_info(cnt == kRawItemElems ? kItemRaw : kItemRawTail, 0, kRawBits, cnt);
      out.push_back(it);
    }
    return;
  }
  const uint32_t B = d.bucket;
  const uint32_t mb = block_meta_bytes(n, B);
  const uint32_t gpb = bucket_groups(B);
  const bool pow2 = B >= 8 && (B & (B - 1)) == 0 && B <= slice;
  uint32_t off = 0;
  if (pow2) {
    uint32_t lg = 0;
    while ((8u << lg) < B) ++lg;
    for (; off + slice <= n; off += slice) {
      WarpItem it;
      it.elem_off = d.elem_off + off;
      it.meta_off = d.wire_off + (off / B) * 8u;
      it.pay_off = d.wire_off + mb + (off / 8u) * (uint32_t)bits;
      it.info = make_item_info(kItemFull, lg, bits, slice);
      out.push_back(it);
    }
  }
  for (; off < n; off += B) {  // off is a multiple of B: blocks start on bucket boundaries
    const uint32_t cnt = std::min(B, n - off);
    WarpItem it;
    it.elem_off = d.elem_off + off;
    it.meta_off = d.wire_off + (off / B) * 8u;
    it.pay_off = d.wire_off + mb + (off / B) * gpb * (uint32_t)bits;
    it.info = make_item_info(kItemBucket, 0, bits, cnt);
    out.push_back(it);
  }
}

}  // namespace

Plan build_plan(const std::vector<LayerSpec>& layers, const PlanOptions& opt) {
  if (opt.world < 1 || opt.world > kMaxPeers) throw std::invalid_argument("cgx plan: bad world size");
  if (opt.lanes < 1) throw std::invalid_argument("cgx plan: lanes must be >= 1");
  if (opt.max_block_elems < 8 || opt.max_block_elems > kMaxBlockElems)
    throw std::invalid_argument("cgx plan: max_block_elems out of range");

  std::vector<Segment> segs;
  segs.reserve(layers.size() * 2);
  uint64_t prev_end = 0;
  for (const LayerSpec& l : layers) {
    if (l.bits < 1 || (l.bits > 8 && l.bits < kRawBits))
      throw std::invalid_argument("cgx plan: quantization bits must be in 1..8 (or 32 for none)");
    if (l.elem_off < prev_end) throw std::invalid_argument("cgx plan: layers overlap or are unsorted");
    prev_end = l.elem_off + l.numel;
    if (prev_end >= (1ull << 32)) throw std::invalid_argument("cgx plan: buffer exceeds 2^32 elements");
    if (l.bits >= kRawBits) {
      push_segment(segs, l.elem_off, l.numel, kRawBits, 0, opt.max_block_elems);
      continue;
    }
    uint32_t bucket = std::max<uint32_t>(1u, l.bucket);
    uint64_t main = l.numel;
    if (opt.skip_incomplete) main = l.numel / bucket * bucket;
    push_segment(segs, l.elem_off, main, l.bits, bucket, opt.max_block_elems);
    push_segment(segs, l.elem_off + main, l.numel - main, kRawBits, 0, opt.max_block_elems);
  }
  uint64_t total = 0;
  for (Segment& s : segs) {
    s.start = total;
    total += s.n;
  }

  Plan p;
  p.world = opt.world;
  p.dtype = opt.dtype;
  p.numel = total;
  {
    // 1024-element slices when every compressed layer quantizes in 1024-element buckets (the
    // Python hook's default), else 512; common bit width of the compressed layers
    bool any = false, all1024 = true;
    int ub = -1;
    for (const Segment& s : segs) {
      if (s.bits >= kRawBits) continue;
      any = true;
      all1024 = all1024 && s.bucket == 1024;
      ub = (ub == -1 || ub == s.bits) ? s.bits : 0;
    }
    p.slice_elems = (any && all1024) ? 1024u : 512u;
    p.uniform_bits = ub > 0 ? ub : 0;
  }
  uint64_t per_rank = total / (uint64_t)opt.world;
  uint64_t want = per_rank / std::max<uint32_t>(1u, opt.min_lane_elems);
  p.lanes = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)opt.lanes, want));
  const int W = p.world, G = p.lanes;
  const uint64_t S = (uint64_t)W * G;

  // slot boundaries in the concatenated space, snapped down to bucket starts
  std::vector<uint64_t> cut(S + 1);
  size_t si = 0;
  for (uint64_t k = 0; k <= S; ++k) {
    // 128-bit safe: total < 2^32, k < 2^12
    uint64_t t = (k == S) ? total : (total * k) / S;
    if (t >= total) {
      cut[k] = total;
      continue;
    }
    while (si + 1 < segs.size() && t >= segs[si].start + segs[si].n) ++si;
    uint64_t rel = t - segs[si].start;
    rel = rel / segs[si].bucket * segs[si].bucket;
    cut[k] = segs[si].start + rel;
  }

  const int elsize = dtype_size(opt.dtype);
  p.lane_first.assign(S + 1, 0);
  p.item_first.assign(S + 1, 0);
  p.chunk_wire_bytes.assign(W, 0);
  p.chunk_elems.assign(W, 0);
  size_t seg = 0;
  for (uint64_t k = 0; k < S; ++k) {
    p.lane_first[k] = (uint32_t)p.blocks.size();
    p.item_first[k] = (uint32_t)p.items.size();
    const int chunk = (int)(k / G);
    uint64_t a = cut[k], b = cut[k + 1];
    while (a < b) {
      while (seg + 1 < segs.size() && a >= segs[seg].start + segs[seg].n) ++seg;
      const Segment& s = segs[seg];
      uint64_t seg_end = s.start + s.n;
      uint64_t piece_end = std::min(b, seg_end);
      uint64_t rel = a - s.start;  // multiple of s.bucket by construction
      // at most block_elems per block, counted from the start of the lane's piece
      uint64_t next = std::min<uint64_t>(piece_end - s.start, rel + (uint64_t)s.block_elems);
      uint32_t n = (uint32_t)(next - rel);
      BlockDesc d;
      d.elem_off = (uint32_t)(s.elem_off + rel);
      uint64_t woff = p.chunk_wire_bytes[chunk];
      d.wire_off = (uint32_t)woff;
      d.n_bits = (n & 0xFFFFFFu) | ((uint32_t)s.bits << 24);
      d.bucket = s.bucket;
      uint64_t wb = block_wire_bytes(n, s.bits, s.bucket, elsize);
      if (woff + wb >= (1ull << 32)) throw std::invalid_argument("cgx plan: chunk wire size exceeds 4 GiB");
      p.chunk_wire_bytes[chunk] = (uint32_t)(woff + wb);
      p.chunk_elems[chunk] += n;
      p.block_item_first.push_back((uint32_t)p.items.size());
      emit_items(d, elsize, p.slice_elems, p.items);
      p.blocks.push_back(d);
      a = s.start + next;
    }
  }
  p.lane_first[S] = (uint32_t)p.blocks.size();
  p.item_first[S] = (uint32_t)p.items.size();
  p.block_item_first.push_back((uint32_t)p.items.size());
  for (int c = 0; c < W; ++c) {
    p.max_chunk_wire = std::max(p.max_chunk_wire, p.chunk_wire_bytes[c]);
    p.total_wire += p.chunk_wire_bytes[c];
  }
  return p;
}

uint64_t plan_key(const std::vector<LayerSpec>& layers, const PlanOptions& opt) {
  uint64_t h = 0xcbf29ce484222325ull;
  auto mix = [&h](uint64_t v) {
    for (int i = 0; i < 8; ++i) {
      h ^= (v >> (i * 8)) & 0xFF;
      h *= 0x100000001b3ull;
    }
  };
  mix((uint64_t)opt.world);
  mix((uint64_t)opt.lanes);
  mix((uint64_t)opt.dtype);
  mix((uint64_t)opt.skip_incomplete);
  mix(opt.min_lane_elems);
  mix(opt.max_block_elems);
  mix(layers.size());
  for (const LayerSpec& l : layers) {
    mix(l.elem_off);
    mix(l.numel);
    mix((uint64_t)l.bits);
    mix(l.bucket);
  }
  return h;
}

std::string describe_plan(const Plan& p) {
  std::ostringstream os;
  os << "Plan{world=" << p.world << ", lanes=" << p.lanes << ", numel=" << p.numel
     << ", blocks=" << p.blocks.size() << ", max_chunk_wire=" << p.max_chunk_wire
     << ", total_wire=" << p.total_wire << "}";
  return os.str();
}

}  // namespace cgx
