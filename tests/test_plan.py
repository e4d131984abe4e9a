"""Plan builder: layer-aware split into blocks / rank chunks / lanes.
Behavioural counterpart of Quantizer::GetSizesAndOffsets and the per-slice
walk of the reference (/root/reference/src/common/compressor.cc:62-194,265-299)."""
import pytest
import torch

import torch_cgx_b200 as cgx

C = cgx._C


def blocks_of(plan):
    return [tuple(int(v) for v in row) for row in plan["blocks"]]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("lanes", [1, 7, 32])
def test_blocks_tile_layers_exactly(world, lanes):
    layers = [(0, 100_000, 4, 512), (100_000, 300, 32, 512), (100_300, 5_001, 8, 64), (105_301, 77, 2, 128)]
    plan = C.build_plan(layers, world=world, lanes=lanes, min_lane_elems=256)
    blocks = sorted(blocks_of(plan))
    # blocks cover every layer exactly once, in order, without overlap
    pos = 0
    for off, n, bits, bucket, woff in blocks:
        assert off == pos
        assert 0 < n <= C.MAX_BLOCK_ELEMS
        pos += n
    assert pos == 105_378 == plan["numel"]
    # every block lies inside one layer and starts on a bucket boundary of it
    for off, n, bits, bucket, woff in blocks:
        owner = [l for l in layers if l[0] <= off < l[0] + l[1]]
        assert len(owner) == 1
        lo, ln, lbits, lbucket = owner[0]
        assert off + n <= lo + ln
        assert bits == lbits
        if bits < 32:
            assert bucket == lbucket
            assert (off - lo) % bucket == 0
        assert woff % 16 == 0


def test_chunks_are_balanced_and_wire_offsets_disjoint():
    n = 6_500_000
    plan = C.build_plan([(0, n, 4, 512)], world=8, lanes=148)
    assert plan["lanes"] == 148
    elems = plan["chunk_elems"]
    assert sum(elems) == n
    assert max(elems) - min(elems) <= 2 * 512
    # per chunk: records are laid out back to back
    lf = plan["lane_first"]
    G = plan["lanes"]
    blocks = blocks_of(plan)
    for c in range(8):
        cur = 0
        for b in range(lf[c * G], lf[(c + 1) * G]):
            off, bn, bits, bucket, woff = blocks[b]
            assert woff == cur
            nb = -(-bn // bucket)
            cur += (nb * 8 + 15) // 16 * 16 + ((-(-bn // 8)) * bits + 15) // 16 * 16
        assert cur == plan["chunk_wire_bytes"][c]
    # lanes are balanced to within one block
    per_lane = []
    for s in range(8 * G):
        per_lane.append(sum(blocks[b][1] for b in range(lf[s], lf[s + 1])))
    assert max(per_lane) - min(per_lane) <= 1024


def test_small_tensors_use_few_lanes():
    plan = C.build_plan([(0, 4096, 4, 512)], world=4, lanes=148)
    assert plan["lanes"] == 1
    plan = C.build_plan([(0, 8, 32, 512)], world=8, lanes=148)
    assert plan["lanes"] == 1
    assert plan["numel"] == 8


def test_skip_incomplete_tail_is_raw():
    plan = C.build_plan([(0, 1025, 4, 512)], world=1, lanes=1, skip_incomplete=True)
    bl = blocks_of(plan)
    assert [(b[0], b[1], b[2]) for b in bl] == [(0, 1024, 4), (1024, 1, 32)]
    plan = C.build_plan([(0, 1025, 4, 512)], world=1, lanes=1, skip_incomplete=False)
    assert [(b[0], b[1], b[2]) for b in blocks_of(plan)] == [(0, 1025, 4)]


def test_compression_ratio_matches_baseline_table():
    # BASELINE.md class-B rows use T-sized meta; ours is fp32 meta (8 B per bucket)
    n = 512 * 1000
    plan = C.build_plan([(0, n, 4, 512)], world=1, lanes=1)
    assert plan["total_wire"] == n // 2 + 8 * 1000
    assert abs(4 * n / plan["total_wire"] - 32 / (4 + 64 / 512)) < 1e-9


def test_invalid_inputs():
    with pytest.raises(Exception):
        C.build_plan([(0, 10, 9, 512)], world=1, lanes=1)
    with pytest.raises(Exception):
        C.build_plan([(0, 10, 4, 512), (5, 10, 4, 512)], world=1, lanes=1)
    with pytest.raises(Exception):
        C.build_plan([(0, 10, 4, 512)], world=0, lanes=1)


def test_split_for_fusion_never_drops_layers():
    layers = [(0, 1000, 4, 512), (1000, 5_000_000, 4, 512), (5_001_000, 10, 32, 512), (5_001_010, 3000, 8, 64)]
    groups = C.split_for_fusion(layers, 4, 2 << 20)  # 2 MB => 512K elements
    flat = [l for g in groups for l in g]
    assert sum(l[1] for l in flat) == sum(l[1] for l in layers)
    # order and contiguity preserved
    pos = 0
    for off, n, bits, bucket in flat:
        assert off == pos
        pos += n
    for g in groups:
        assert sum(l[1] for l in g) <= 512 * 1024
    # slices of the big layer start on bucket boundaries
    for off, n, bits, bucket in flat:
        if 1000 <= off < 5_001_000:
            assert (off - 1000) % 512 == 0
