"""The `cgx` backend on CUDA tensors: one process per rank. With >= 2 GPUs each
rank gets its own device and the full reference test-suite
(/root/reference/test/test_cgx.py) runs, plus DDP + cgx_hook. With a single GPU
only the world_size == 1 path is exercised here (the multi-rank protocol is
covered by the single-process virtual-rank tests in test_kernels_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from _dist_utils import spawn

pytestmark = pytest.mark.gpu


def _world1(rank, world):
    import torch_cgx_b200 as cgx

    torch.cuda.set_device(0)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "4"
        x = torch.randn(100_000, device="cuda")
        ref = cgx.ops.fake_quantize(x.cpu(), 4, 512)
        y = x.clone()
        dist.all_reduce(y)
        torch.cuda.synchronize()
        # world 1: the owner's requantize + self-decode is all that happens
        assert torch.equal(y.cpu(), ref)
        be = cgx.get_backend()
        assert be is not None and be.p2p_ready() and be.stats()[1] >= 1
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        z = x.clone()
        dist.all_reduce(z)
        assert torch.equal(z, x)
        t = torch.tensor([3], dtype=torch.int32, device="cuda")
        dist.all_reduce(t)  # NCCL delegate
        assert t.item() == 3
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_backend_world1():
    spawn(_world1, 1)


def _reference_suite(rank, world):
    import torch_cgx  # noqa: F401

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        # test_compressed_exact
        for q in (2, 4, 8):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(q)
            for dtype in (torch.float16, torch.float32, torch.int32):
                for n in (1, 2, 8, 128, 1024, 1_000_000):
                    for _ in range(3):
                        t = torch.tensor([rank + 1] * n, dtype=dtype, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.SUM)
                        exp = torch.tensor([world * (world + 1) // 2] * n, dtype=dtype, device=dev)
                        assert torch.equal(t, exp), f"bits {q} dtype {dtype} n {n}"
        # test_compressed_non_exact
        for q in (2, 3, 4, 6, 8):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(q)
            for dtype in (torch.float16, torch.float32):
                for n in (128, 1024, 1025, 16_384, 1_000_000):
                    ar = np.arange(-n / 2, n / 2, 1.0)
                    if dtype == torch.float16:
                        ar = ar * 1e-3
                    for bucket in (64, 512, 2048):
                        os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = str(bucket)
                        t = torch.tensor((rank + 1) * ar, dtype=dtype, device=dev)
                        exp = torch.tensor((world * (world + 1) / 2) * ar, dtype=dtype, device=dev)
                        dist.all_reduce(t)
                        err = torch.norm((t - exp).float(), p=float("inf")).item()
                        assert err < 2 * min(bucket, n) / ((1 << q) - 1) * world * (world + 1)
                        # replicas are bit-identical
                        g = [torch.empty_like(t) for _ in range(world)]
                        dist.all_gather(g, t)
                        assert all(torch.equal(g[0], gi) for gi in g)
        # test_uncompressed
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        for dtype in (torch.float16, torch.float32, torch.bfloat16, torch.int32):
            for n in (1, 8, 1024, 1_000_000):
                t = torch.tensor([rank + 1] * n, dtype=dtype, device=dev)
                dist.all_reduce(t)
                assert torch.equal(t, torch.tensor([world * (world + 1) // 2] * n, dtype=dtype, device=dev))
        # AVG is fused into the kernel
        t = torch.full((10_000,), float(rank + 1), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        assert torch.allclose(t, torch.full_like(t, (world + 1) / 2))
        # async work + future
        t = torch.ones(1 << 20, device=dev)
        w = dist.all_reduce(t, async_op=True)
        w.wait()
        assert torch.equal(t, torch.full_like(t, float(world)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
def test_reference_suite_two_gpus():
    spawn(_reference_suite, 2, timeout=600)


def _ddp_gpu(rank, world, hook="python"):
    import torch_cgx_b200 as cgx
    from cgx_utils import CGXState, cgx_hook
    from torch.nn.parallel import DistributedDataParallel as DDP

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        from torch_cgx_b200.models import resnet18

        torch.manual_seed(0)
        model = resnet18(num_classes=10, cifar_stem=True).cuda()
        ddp = DDP(model, device_ids=[rank])
        state = CGXState(None, layer_min_size=1024, compression_params={"bits": 4, "bucket_size": 512})
        if hook == "native":
            assert cgx.register_cgx_hook(ddp, state) is not None
        else:
            ddp.register_comm_hook(state, cgx_hook)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.05, momentum=0.9)
        losses = []
        torch.manual_seed(1234 + rank)
        x = torch.randn(32, 3, 32, 32, device="cuda")
        y = torch.randint(0, 10, (32,), device="cuda")
        for step in range(12):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(ddp(x), y)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert losses[-1] < losses[0], losses  # it learns (memorises the fixed batch) through 4-bit gradients
        be = cgx.get_backend()
        assert be.stats()[1] > 0 and be.stats()[3] < be.stats()[4]  # compressed bytes < raw bytes
        flat = torch.cat([p.detach().reshape(-1) for p in ddp.parameters()])
        g = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(g, flat)
        assert all(torch.equal(g[0], gi) for gi in g)  # replicas bit-identical
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
def test_reference_suite_all_gpus():
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    spawn(_reference_suite, min(n, 8), timeout=900)


@pytest.mark.multigpu
@pytest.mark.parametrize("hook", ["python", "native"])
def test_ddp_hook_two_gpus(hook):
    spawn(_ddp_gpu, 2, args=(hook,), timeout=600)


def _nccl_transport(rank, world):
    import torch_cgx_b200 as cgx

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        for n, bits, bucket, dtype in [(100_003, 4, 512, torch.float32), (50_000, 8, 64, torch.float16),
                                       (300_000, 2, 1024, torch.bfloat16)]:
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(bits)
            os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = str(bucket)
            g = torch.Generator().manual_seed(n)
            ins = [(torch.randn(n, generator=g) * (r + 1)).to(dtype) for r in range(world)]
            t = ins[rank].to(dev)
            dist.all_reduce(t)
            ref = [x.clone() for x in ins]
            cgx._C.sra_simulate(ref, [(0, n, bits, bucket)], 1, False, False, False, 0, 0, 2048)
            # generic SRA over NCCL send/recv + standalone kernels == oracle == fused kernel
            assert torch.equal(t.cpu(), ref[rank])
        be = cgx.get_backend()
        assert not be.p2p_ready()  # the peer-memory heap was never created in this mode
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
def test_nccl_transport_generic_sra():
    spawn(_nccl_transport, 2, env={"CGX_INNER_COMMUNICATOR_TYPE": "NCCL"}, timeout=600)


def _hier_gpu(rank, world):
    import torch_cgx_b200 as cgx

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        be = cgx.get_backend()
        assert be.local_size() == 2
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "8"
        n = 200_000
        g = torch.Generator().manual_seed(3)
        ins = [torch.randn(n, generator=g) * (r + 1) for r in range(world)]
        t = ins[rank].to(dev)
        dist.all_reduce(t)
        exact = sum(ins).to(dev)
        span = max(float(x.max() - x.min()) for x in ins) * world
        assert (t - exact).abs().max().item() < span / 255 * 8
        gathered = [torch.empty_like(t) for _ in range(world)]
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        dist.all_gather(gathered, t)
        assert all(torch.equal(gathered[0], gi) for gi in gathered)
        t = torch.full((10_000,), float(rank + 1), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        assert torch.allclose(t, torch.full_like(t, (world + 1) / 2))
        assert be.p2p_ready() and be.stats()[1] > 0  # the intra-node stage ran the fused kernel
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
@pytest.mark.parametrize("intra_broadcast", ["1", "0"])
def test_hierarchical_simulated_nodes(intra_broadcast):
    if torch.cuda.device_count() < 4:
        pytest.skip("needs >= 4 GPUs")
    spawn(_hier_gpu, 4, env={"CGX_LOCAL_SIZE": "2", "CGX_INTRA_BROADCAST": intra_broadcast}, timeout=600)


# ---- generic reducers on device memory (NCCL send/recv transport): Ring and all-to-all ----
def _generic_reduction(rank, world, kind):
    import torch_cgx_b200 as cgx

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        for n, bits, bucket in [(100_003, 4, 512), (64_000, 8, 64)]:
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(bits)
            os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = str(bucket)
            g = torch.Generator().manual_seed(n)
            ins = [torch.randn(n, generator=g) * (r + 1) for r in range(world)]
            t = ins[rank].to(dev)
            dist.all_reduce(t)
            exact = sum(ins).to(dev)
            span = sum(float(x.max() - x.min()) for x in ins)
            # <= one step per hop of the ring / per contribution
            assert (t - exact).abs().max().item() < span / ((1 << bits) - 1) * (world + 1)
            got = [torch.empty_like(t) for _ in range(world)]
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
            dist.all_gather(got, t)
            if kind == "RING":  # the ring forwards the SAME packed bytes: replicas identical
                assert all(torch.equal(got[0], gi) for gi in got)
        be = cgx.get_backend()
        assert not be.p2p_ready()
        # uncompressed through the same reducer is exact
        t = torch.full((70_001,), float(rank + 1), device=dev)
        dist.all_reduce(t)
        assert torch.equal(t, torch.full_like(t, float(world * (world + 1) // 2)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
@pytest.mark.parametrize("kind", ["RING", "ALLTOALL"])
def test_generic_ring_and_alltoall_reducers_on_gpus(kind):
    env = {"CGX_INNER_COMMUNICATOR_TYPE": "NCCL"}
    if kind == "RING":
        env["CGX_INNER_REDUCTION_TYPE"] = "RING"
    else:
        env["CGX_DEBUG_ALL_TO_ALL_REDUCTION"] = "1"
    spawn(_generic_reduction, min(4, torch.cuda.device_count()), args=(kind,), env=env, timeout=600)


# ---- CUDA graphs: the epoch lives in device memory, so a captured allreduce replays correctly ----
def _graph_replay(rank, world):
    import torch_cgx_b200 as cgx

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "4"
        sizes = [1 << 20, 300_000, 2_000_003]  # three-phase kernel, one-shot kernel, three-phase
        bufs = [torch.zeros(n, device=dev) for n in sizes]
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):  # plans, heap, kernels warm (no allocation may happen under capture)
                for b in bufs:
                    dist.all_reduce(b)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for b in bufs:  # one "step": three buckets
                dist.all_reduce(b)
        for step in range(3):  # replay 3 steps with fresh data each time
            gen = torch.Generator().manual_seed(100 * step + 7)
            ins = [[torch.randn(n, generator=gen) * (r + 1) for r in range(world)] for n in sizes]
            for b, per_rank in zip(bufs, ins):
                b.copy_(per_rank[rank])
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            for b, per_rank, n in zip(bufs, ins, sizes):
                ref = [x.clone() for x in per_rank]
                if n * 4 <= (2 << 20):
                    cgx._C.oneshot_simulate(ref, [(0, n, 4, 512)], 1, False, False, False, 0, 0, 4096)
                else:
                    cgx._C.sra_simulate(ref, [(0, n, 4, 512)], 1, False, False, False, 0, 0, 4096)
                assert torch.equal(b.cpu(), ref[rank]), f"replay {step}: n={n} differs from the oracle"
        cgx.get_backend().check_health()
        # and ordinary (eager) calls still line up with the peers afterwards: epochs stayed in step
        t = torch.full((1 << 20,), float(rank + 1), device=dev)
        dist.all_reduce(t)
        assert torch.equal(t, torch.full_like(t, float(world * (world + 1) // 2)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
def test_cuda_graph_capture_and_replay_of_allreduce():
    spawn(_graph_replay, 2, timeout=600)


def _ddp_graph(rank, world):
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP

    import torch_cgx_b200 as cgx

    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)

        def make():
            torch.manual_seed(0)
            return nn.Sequential(nn.Linear(1024, 2048), nn.ReLU(), nn.Linear(2048, 2048), nn.ReLU(),
                                 nn.Linear(2048, 256)).to(dev)

        def batch(step):
            g = torch.Generator().manual_seed(1000 * step + rank)
            return torch.randn(64, 1024, generator=g).to(dev), torch.randn(64, 256, generator=g).to(dev)

        def run(use_graph):
            cgx.reset_layers()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model = make()
                ddp = DDP(model, device_ids=[rank], gradient_as_bucket_view=True, bucket_cap_mb=4)
                state = cgx.CGXState(None, layer_min_size=1024, compression_params={"bits": 8, "bucket_size": 512})
                cgx.register_cgx_hook(ddp, state)
                opt = torch.optim.SGD(ddp.parameters(), lr=0.01)
                x, y = batch(0)
                sx, sy = x.clone(), y.clone()

                def step():
                    opt.zero_grad(set_to_none=False)
                    loss = torch.nn.functional.mse_loss(ddp(sx), sy)
                    loss.backward()
                    opt.step()

                for _ in range(11):  # DDP rebuilds its buckets, the hook registers layers at step 3
                    step()
                torch.cuda.synchronize()
                graph = None
                if use_graph:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        step()
                for s in range(1, 4):
                    x, y = batch(s)
                    sx.copy_(x)
                    sy.copy_(y)
                    if use_graph:
                        graph.replay()
                    else:
                        step()
                torch.cuda.synchronize()
            out = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
            del ddp
            return out

        eager = run(False)
        graphed = run(True)  # 11 warm-up steps + 1 captured (not executed) + 3 replays == 11 + 3 eager steps
        assert torch.isfinite(graphed).all()
        assert torch.allclose(eager, graphed, rtol=0, atol=1e-6), (eager - graphed).abs().max()
        cgx.get_backend().check_health()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.multigpu
def test_ddp_steps_replay_inside_a_cuda_graph():
    spawn(_ddp_graph, 2, timeout=600)
