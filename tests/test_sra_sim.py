"""Numerics of the compressed Scatter-Reduce-AllGather on the CPU simulator
(the bit-exact oracle of the fused CUDA kernel). Mirrors the reference's
test_compressed_exact / test_compressed_non_exact / test_uncompressed
(/root/reference/test/test_cgx.py:69-101) without needing mpirun or GPUs."""
import numpy as np
import pytest
import torch

import torch_cgx_b200 as cgx

C = cgx._C


def run(tensors, layers, **kw):
    ts = [t.clone() for t in tensors]
    C.sra_simulate(ts, layers, **kw)
    return ts


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_compressed_exact(world, dtype):
    for bits in (2, 4, 8):
        for n in (16, 128, 1024, 100_000):
            ins = [torch.full((n,), float(r + 1), dtype=dtype) for r in range(world)]
            outs = run(ins, [(0, n, bits, 512)], lanes=4)
            expect = torch.full((n,), float(world * (world + 1) // 2), dtype=dtype)
            for o in outs:
                assert torch.equal(o, expect)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("bits", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_compressed_non_exact_reference_bound(world, bits, dtype):
    for n in (128, 1025, 16_384):
        for bucket in (64, 512, 2048):
            ar = np.arange(-n / 2, n / 2, 1.0)
            if dtype == torch.float16:
                ar = ar * 1e-3
            ins = [torch.tensor((r + 1) * ar, dtype=dtype) for r in range(world)]
            expected = torch.tensor((world * (world + 1) / 2) * ar, dtype=dtype)
            outs = run(ins, [(0, n, bits, bucket)], lanes=8)
            # all ranks bit-identical
            for o in outs[1:]:
                assert torch.equal(o, outs[0])
            err = (outs[0].float() - expected.float()).abs().max().item()
            coef = world * (world + 1)
            assert err < 2 * min(bucket, n) / ((1 << bits) - 1) * coef


@pytest.mark.parametrize("world", [2, 3, 8])
def test_uncompressed_is_exact_sum(world):
    n = 10_001
    torch.manual_seed(0)
    ins = [torch.randint(-100, 100, (n,)).float() for _ in range(world)]
    outs = run(ins, [(0, n, 32, 512)], lanes=3)
    for o in outs:
        assert torch.equal(o, sum(ins))


def test_average_prescale():
    world, n = 4, 5000
    ins = [torch.full((n,), float(4 * (r + 1))) for r in range(world)]
    outs = run(ins, [(0, n, 4, 512)], lanes=2, average=True)
    for o in outs:
        assert torch.equal(o, torch.full((n,), 10.0))


def test_result_independent_of_lane_count_for_aligned_layers():
    # bucket boundaries are tied to the layer, not to the lane/chunk split
    world, n = 4, 64 * 1024
    torch.manual_seed(1)
    ins = [torch.randn(n) for _ in range(world)]
    a = run(ins, [(0, n, 4, 512)], lanes=1)
    b = run(ins, [(0, n, 4, 512)], lanes=16, min_lane_elems=512)
    assert torch.equal(a[0], b[0])


def test_mixed_layers_only_compress_what_is_marked():
    world = 4
    torch.manual_seed(2)
    layers = [(0, 4096, 4, 512), (4096, 100, 32, 512), (4196, 3000, 8, 64), (7196, 7, 32, 512)]
    n = 7203
    ins = [torch.randn(n) for _ in range(world)]
    outs = run(ins, layers, lanes=4, min_lane_elems=256)
    exact = sum(ins)
    o = outs[0]
    assert torch.allclose(o[4096:4196], exact[4096:4196], rtol=0, atol=1e-5)
    assert torch.allclose(o[7196:], exact[7196:], rtol=0, atol=1e-5)
    assert not torch.allclose(o[:4096], exact[:4096], rtol=0, atol=1e-4)
    # 8-bit layer is much closer than the 4-bit one
    e4 = (o[:4096] - exact[:4096]).abs().mean()
    e8 = (o[4196:7196] - exact[4196:7196]).abs().mean()
    assert e8 < e4 / 4


def test_stochastic_sra_ranks_identical_and_seeded():
    world, n = 4, 20_000
    torch.manual_seed(3)
    ins = [torch.randn(n) for _ in range(world)]
    a = run(ins, [(0, n, 2, 512)], lanes=4, stochastic=True, seed=5, seq=9)
    b = run(ins, [(0, n, 2, 512)], lanes=4, stochastic=True, seed=5, seq=9)
    c = run(ins, [(0, n, 2, 512)], lanes=4, stochastic=True, seed=6, seq=9)
    for o in a[1:]:
        assert torch.equal(o, a[0])
    assert torch.equal(a[0], b[0])
    assert not torch.equal(a[0], c[0])


def test_oneshot_oracle_properties():
    # one quantization per contribution: exact on constants, error below SRA's two rounds
    world, n = 4, 50_000
    ins = [torch.full((n,), float(r + 1)) for r in range(world)]
    outs = [t.clone() for t in ins]
    C.oneshot_simulate(outs, [(0, n, 4, 512)], lanes=4)
    for o in outs:
        assert torch.equal(o, torch.full((n,), 10.0))
    torch.manual_seed(5)
    ins = [torch.randn(n) for _ in range(world)]
    a = [t.clone() for t in ins]
    b = [t.clone() for t in ins]
    C.oneshot_simulate(a, [(0, n, 4, 512)], lanes=4)
    C.sra_simulate(b, [(0, n, 4, 512)], lanes=4)
    exact = sum(ins)
    for o in a[1:]:
        assert torch.equal(o, a[0])
    assert (a[0] - exact).abs().mean() < (b[0] - exact).abs().mean()
    # raw layers are summed exactly
    c = [t.clone() for t in ins]
    C.oneshot_simulate(c, [(0, n, 32, 512)], lanes=2)
    assert torch.allclose(c[0], exact, atol=1e-5)
