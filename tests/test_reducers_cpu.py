"""Generic reducers (SRA / Ring / all-to-all) and the hierarchical allreduce on
host memory over the Gloo transport, world sizes 2-4, no GPU. The SRA result
must be BIT-IDENTICAL to the single-process oracle (and therefore to the fused
CUDA kernel). Reference: scatter_reduce_allgather.cc / ring.cc / reducer.cc and
the intra -> cross -> broadcast hierarchy of mpi_allreduce_operations.cc:139-185."""
import os

import pytest
import torch
import torch.distributed as dist

from _dist_utils import spawn


def _inputs(world, n, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(n, generator=g) * (r + 1)).to(dtype) for r in range(world)]


def _sra(rank, world, dtype_name):
    import torch_cgx_b200 as cgx

    dtype = getattr(torch, dtype_name)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        for n, bits, bucket in [(100_003, 4, 512), (5_000, 8, 64), (70_000, 2, 2048), (33, 4, 512)]:
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(bits)
            os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = str(bucket)
            ins = _inputs(world, n, dtype, seed=n)
            t = ins[rank].clone()
            dist.all_reduce(t)
            ref = [x.clone() for x in ins]
            eff_bits = bits if n > 16 else 32
            cgx._C.sra_simulate(ref, [(0, n, eff_bits, bucket)], 1, False, False, False, 0, 0, 2048)
            assert torch.equal(t, ref[rank]), f"n={n} bits={bits}"
        # AVG is folded into the prescale
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        t = torch.full((1000,), float(rank + 1), dtype=dtype)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        assert torch.allclose(t.float(), torch.full((1000,), (world + 1) / 2), atol=1e-2)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_generic_sra_matches_oracle_fp32(world):
    spawn(_sra, world, args=("float32",), env={"CGX_COMPRESS_CPU": "1"})


def test_generic_sra_matches_oracle_bf16():
    spawn(_sra, 2, args=("bfloat16",), env={"CGX_COMPRESS_CPU": "1"})


def _bounded(rank, world, identical):
    import torch_cgx_b200  # noqa: F401  (registers the backend)

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        for bits in (4, 8):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(bits)
            os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = "512"
            n = 40_000
            ins = _inputs(world, n, torch.float32, seed=bits)
            t = ins[rank].clone()
            dist.all_reduce(t)
            exact = sum(ins)
            span = max(float(x.max() - x.min()) for x in ins) * world
            tol = span / ((1 << bits) - 1) * (world + 1)
            assert (t - exact).abs().max().item() < tol
            g = [torch.empty_like(t) for _ in range(world)]
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
            dist.all_gather(g, t)
            same = all(torch.equal(g[0], gi) for gi in g)
            assert same == identical
        # uncompressed through the same reducer is exact
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        t = torch.arange(1000.0) * (rank + 1)
        dist.all_reduce(t)
        assert torch.equal(t, torch.arange(1000.0) * (world * (world + 1) // 2))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_ring_reducer():
    spawn(_bounded, 4, args=(True,), env={"CGX_COMPRESS_CPU": "1", "CGX_INNER_REDUCTION_TYPE": "RING"})


def test_all_to_all_debug_reducer():
    spawn(_bounded, 3, args=(False,), env={"CGX_COMPRESS_CPU": "1", "CGX_DEBUG_ALL_TO_ALL_REDUCTION": "1"})


def test_dummy_compression_is_exact():
    spawn(_exact, 2, env={"CGX_COMPRESS_CPU": "1", "CGX_DEBUG_DUMMY_COMPRESSION": "1",
                           "CGX_COMPRESSION_QUANTIZATION_BITS": "4"})


def _exact(rank, world):
    import torch_cgx_b200  # noqa: F401

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        t = torch.arange(5000.0) * (rank + 1)
        dist.all_reduce(t)
        assert torch.equal(t, torch.arange(5000.0) * (world * (world + 1) // 2))
    finally:
        dist.destroy_process_group()


def _hier(rank, world):
    import torch_cgx_b200 as cgx

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        be = cgx.get_backend(device="cpu")
        assert be.local_size() == 2
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "8"
        n = 30_000
        ins = _inputs(world, n, torch.float32, seed=7)
        t = ins[rank].clone()
        be.reset_stats()
        dist.all_reduce(t)
        if os.environ.get("CGX_INTRA_BROADCAST") == "1" and rank % 2 == 0:
            # leader: intra SRA + cross ring + COMPRESSED broadcast -- a raw broadcast alone would be n * 4 bytes
            assert be.stats()[3] < 0.8 * n * 4, be.stats()
        exact = sum(ins)
        span = max(float(x.max() - x.min()) for x in ins) * world
        assert (t - exact).abs().max().item() < span / 255 * 8
        g = [torch.empty_like(t) for _ in range(world)]
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        dist.all_gather(g, t)
        assert all(torch.equal(g[0], gi) for gi in g), "replicas must end bit-identical"
        # exact when nothing is compressed, AVG over the WHOLE world
        t = torch.full((4096,), float(rank + 1))
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        assert torch.allclose(t, torch.full((4096,), (world + 1) / 2))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("intra_broadcast", ["1", "0"])
def test_hierarchical_two_simulated_nodes(intra_broadcast):
    spawn(_hier, 4, env={"CGX_COMPRESS_CPU": "1", "CGX_LOCAL_SIZE": "2", "CGX_INTRA_BROADCAST": intra_broadcast,
                          "CGX_CROSS_REDUCTION_TYPE": "RING"})
