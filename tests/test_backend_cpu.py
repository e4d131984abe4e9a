"""The `cgx` torch.distributed backend on CPU tensors, world_size=2, spawned
with torch.multiprocessing over a TCPStore (BASELINE.json configs[0]:
"test_cgx.py allreduce correctness world_size=2 on CPU, bits=32 (plumbing)").
Every collective DDP needs (§3.5 of SURVEY.md) must work."""
import os

import pytest
import torch
import torch.distributed as dist

from _dist_utils import spawn


def _plumbing(rank, world):
    import torch_cgx  # noqa: F401  drop-in module name of the reference

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    assert dist.get_backend() == "cgx"
    try:
        # test_uncompressed of the reference, CPU tensors
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        for dtype in (torch.float16, torch.float32, torch.int32):
            for n in (1, 2, 8, 128, 1024, 100_000):
                t = torch.tensor([rank + 1] * n, dtype=dtype)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                assert torch.equal(t, torch.tensor([world * (world + 1) // 2] * n, dtype=dtype))
        # the rest of the c10d surface
        t = torch.arange(10.0) if rank == 0 else torch.zeros(10)
        dist.broadcast(t, src=0)
        assert torch.equal(t, torch.arange(10.0))
        outs = [torch.zeros(3) for _ in range(world)]
        dist.all_gather(outs, torch.full((3,), float(rank)))
        assert [o[0].item() for o in outs] == [float(r) for r in range(world)]
        t = torch.ones(4) * (rank + 1)
        dist.reduce(t, dst=0)
        if rank == 0:
            assert torch.equal(t, torch.ones(4) * 3)
        if rank == 0:
            dist.send(torch.tensor([42.0]), dst=1)
        else:
            r = torch.zeros(1)
            dist.recv(r, src=0)
            assert r.item() == 42.0
        work = dist.all_reduce(torch.ones(5), async_op=True)
        work.wait()
        out = torch.zeros(world * 2)
        dist.all_gather_into_tensor(out, torch.full((2,), float(rank)))
        assert out.tolist() == [0.0, 0.0, 1.0, 1.0]
        rs = torch.zeros(2)
        dist.reduce_scatter_tensor(rs, torch.arange(4.0))
        assert torch.equal(rs, torch.arange(4.0)[rank * 2 : rank * 2 + 2] * world)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_cpu_plumbing_world2():
    spawn(_plumbing, 2)


def _ddp_cpu(rank, world):
    import torch_cgx_b200 as cgx
    from cgx_utils import CGXState, cgx_hook
    from torch.nn.parallel import DistributedDataParallel as DDP

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10))
        ref = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10))
        ref.load_state_dict(model.state_dict())
        ddp = DDP(model)
        state = CGXState(None, layer_min_size=16, compression_params={"bits": 4, "bucket_size": 64})
        ddp.register_comm_hook(state, cgx_hook)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
        for step in range(5):
            torch.manual_seed(100 + step * world + rank)
            x = torch.randn(8, 64)
            y = torch.randint(0, 10, (8,))
            opt.zero_grad()
            torch.nn.functional.cross_entropy(ddp(x), y).backward()
            opt.step()
        assert state.step == 5
        # layers were registered on the third backward pass
        assert cgx._C.num_registered_buckets() >= 1
        sizes = [s for s, _, _ in cgx._C.registered_bucket(0)]
        assert sum(sizes) > 0
        bits = {b for _, b, _ in cgx._C.registered_bucket(0)}
        assert bits <= {4, 32} and 32 in bits  # biases stay uncompressed
        # replicas stay in sync (CPU tensors are reduced exactly through the Gloo delegate)
        flat = torch.cat([p.detach().reshape(-1) for p in ddp.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert torch.equal(gathered[0], gathered[1])
    finally:
        dist.destroy_process_group()


def test_ddp_hook_cpu_world2():
    spawn(_ddp_cpu, 2)


def _ddp_cpu_native_hook(rank, world):
    import torch_cgx_b200 as cgx
    from torch.nn.parallel import DistributedDataParallel as DDP

    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    try:
        def make():
            torch.manual_seed(0)
            return torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10))

        def train(use_native):
            cgx.reset_layers()
            ddp = DDP(make())
            state = cgx.CGXState(None, layer_min_size=16, compression_params={"bits": 4, "bucket_size": 64})
            handle = None
            if use_native:
                handle = cgx.register_cgx_hook(ddp, state)
                assert handle is not None
            else:
                ddp.register_comm_hook(state, cgx.cgx_hook)
            opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
            for step in range(5):
                torch.manual_seed(100 + step * world + rank)
                x, y = torch.randn(8, 64), torch.randint(0, 10, (8,))
                opt.zero_grad()
                torch.nn.functional.cross_entropy(ddp(x), y).backward()
                opt.step()
            steps = handle.step if handle is not None else state.step
            return torch.cat([p.detach().reshape(-1) for p in ddp.parameters()]), steps

        w_py, s_py = train(False)
        layers_py = cgx._C.registered_bucket(0)
        w_cc, s_cc = train(True)
        assert s_py == s_cc == 5
        assert cgx._C.registered_bucket(0) == layers_py  # same layer table (biases uncompressed)
        assert torch.equal(w_py, w_cc)                   # identical training trajectory
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_native_cpp_hook_matches_python_hook():
    spawn(_ddp_cpu_native_hook, 2)


def test_ddp_hook_cpu_compressed_async_worker():
    # CPU tensors through the compressed generic reducers on the backend's host worker thread
    # (futures complete asynchronously; replicas must still end bit-identical)
    spawn(_ddp_cpu, 2, env={"CGX_COMPRESS_CPU": "1"})
