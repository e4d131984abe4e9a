"""Property tests (hypothesis) of the plan builder and the CPU quantizer on random layer tables."""
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import torch_cgx_b200 as cgx

C = cgx._C

layer = st.tuples(st.integers(1, 40_000), st.sampled_from([1, 2, 3, 4, 5, 6, 7, 8, 32]),
                  st.sampled_from([8, 64, 100, 256, 512, 1000, 1024, 2048, 8192, 20_000]))


def make_layers(specs):
    out, off = [], 0
    for n, bits, bucket in specs:
        out.append((off, n, bits, bucket))
        off += n
    return out, off


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(layer, min_size=1, max_size=12), st.integers(1, 8), st.integers(1, 64),
       st.booleans(), st.sampled_from([0, 1, 2]))
def test_plan_invariants(specs, world, lanes, skip_incomplete, dtype):
    layers, total = make_layers(specs)
    plan = C.build_plan(layers, world=world, lanes=lanes, dtype=dtype, skip_incomplete=skip_incomplete,
                        min_lane_elems=64)
    blocks = [tuple(int(v) for v in row) for row in plan["blocks"]]
    assert plan["numel"] == total
    assert 1 <= plan["lanes"] <= lanes
    # blocks tile [0, total) exactly once
    pos = 0
    for off, n, bits, bucket, woff in sorted(blocks):
        assert off == pos and 0 < n <= C.MAX_BLOCK_ELEMS and woff % 16 == 0
        pos += n
    assert pos == total
    # each block inside one layer, starting on one of its bucket boundaries
    for off, n, bits, bucket, woff in blocks:
        lo, ln, lbits, lbucket = next(l for l in layers if l[0] <= off < l[0] + l[1])
        assert off + n <= lo + ln
        if bits < 32:
            assert bits == lbits and (off - lo) % bucket == 0 and bucket <= C.MAX_BLOCK_ELEMS
            assert -(-n // bucket) <= 512
    # per chunk: wire records are back to back and inside the chunk's size
    lf, G = plan["lane_first"], plan["lanes"]
    es = 4 if dtype == 0 else 2
    for c in range(world):
        cur = 0
        for b in range(lf[c * G], lf[(c + 1) * G]):
            off, n, bits, bucket, woff = blocks[b]
            assert woff == cur
            if bits >= 32:
                cur += (n * es + 15) // 16 * 16
            else:
                # every bucket starts a fresh pack group
                groups = (n // bucket) * -(-bucket // 8) + -(-(n % bucket) // 8)
                cur += (-(-n // bucket) * 8 + 15) // 16 * 16 + (groups * bits + 15) // 16 * 16
        assert cur == plan["chunk_wire_bytes"][c] <= plan["max_chunk_wire"]
    assert sum(plan["chunk_elems"]) == total
    # warp items: tile every block exactly once, in order, with consistent wire offsets
    items = [tuple(int(v) for v in row) for row in plan["items"]]
    bif, itf, slice_elems = plan["block_item_first"], plan["item_first"], plan["slice_elems"]
    assert len(bif) == len(blocks) + 1 and bif[-1] == len(items) and itf[-1] == len(items)
    assert slice_elems in (512, 1024)
    for k in range(world * G):
        assert itf[k] == bif[lf[k]]
    for b, (off, n, bits, bucket, woff) in enumerate(blocks):
        pos = off
        mb = (-(-n // bucket) * 8 + 15) // 16 * 16
        gpb = -(-bucket // 8)
        for eo, mo, po, kind, lg, ibits, cnt in items[bif[b]:bif[b + 1]]:
            assert eo == pos and cnt > 0 and ibits == (32 if bits >= 32 else bits)
            rel = eo - off
            if bits >= 32:
                assert kind in (2, 3) and (kind == 2) == (cnt == 512) and mo == po == woff + rel * es
            else:
                assert rel % bucket == 0 and mo == woff + (rel // bucket) * 8
                assert po == woff + mb + (rel // bucket) * gpb * bits
                if kind == 0:
                    assert cnt == slice_elems and bucket == 8 << lg and bucket <= slice_elems
                else:
                    assert kind == 1 and cnt <= bucket
            pos += cnt
        assert pos == off + n


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(layer, min_size=1, max_size=6), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_sra_oracle_replicas_identical_and_bounded(specs, world, seed):
    layers, total = make_layers(specs)
    g = torch.Generator().manual_seed(seed)
    ins = [torch.randn(total, generator=g) * (r + 1) for r in range(world)]
    outs = [t.clone() for t in ins]
    C.sra_simulate(outs, layers, lanes=4, min_lane_elems=64)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    exact = sum(ins)
    for off, n, bits, bucket in layers:
        err = (outs[0][off:off + n] - exact[off:off + n]).abs().max().item()
        if bits >= 32:
            assert err <= 1e-4 * max(1.0, exact.abs().max().item())
        else:
            span = sum(float(x[off:off + n].max() - x[off:off + n].min()) for x in ins)
            # <= half a step per contribution (phase 1) + half a step of the sum (phase 2)
            assert err <= span / ((1 << bits) - 1) * 1.01 + 1e-5
