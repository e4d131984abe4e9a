#!/usr/bin/env python
"""Multi-rank correctness self-test of the `cgx` backend, cheap enough to run inside bench.py at
every N (the driver only runs bench.py on more than one GPU):

  * the reference test-suite semantics (/root/reference/test/test_cgx.py:69-101): constant tensors
    reduce EXACTLY at 2/4/8 bits, ramps stay inside the reference's error bound, 32 bits is exact;
  * replicas are bit-identical after a compressed allreduce;
  * the fused kernel's result equals the CPU oracle (_C.sra_simulate) bit for bit.

    torchrun --nproc-per-node N bench/selftest.py
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run_selftest(dev: torch.device, group=None) -> dict:
    """Must be called by every rank of an initialised `cgx` process group."""
    import torch_cgx_b200 as cgx

    C = cgx._C
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    saved = {k: os.environ.get(k) for k in ("CGX_COMPRESSION_QUANTIZATION_BITS", "CGX_COMPRESSION_BUCKET_SIZE")}
    checks, fails = 0, []

    def check(ok: bool, what: str):
        nonlocal checks
        checks += 1
        if not ok:
            fails.append(what)

    try:
        # ---- exact on constants (unit == 0 -> every level 0 -> decode == min)
        for q in (2, 4, 8):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(q)
            for dtype in (torch.float16, torch.float32):
                for n in (1, 8, 128, 1024, 100_003):
                    t = torch.full((n,), float(rank + 1), dtype=dtype, device=dev)
                    dist.all_reduce(t, group=group)
                    check(bool((t == world * (world + 1) // 2).all()), f"exact q={q} {dtype} n={n}")
        # ---- reference error bound on ramps, replicas identical
        for q in (2, 4, 8):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(q)
            for bucket in (64, 512, 2048):
                os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = str(bucket)
                for n in (1025, 16_384, 300_001):
                    ar = np.arange(-n / 2, n / 2, 1.0)
                    t = torch.tensor((rank + 1) * ar, dtype=torch.float32, device=dev)
                    exp = torch.tensor((world * (world + 1) / 2) * ar, dtype=torch.float32, device=dev)
                    dist.all_reduce(t, group=group)
                    err = (t - exp).abs().max().item()
                    check(err < 2 * min(bucket, n) / ((1 << q) - 1) * world * (world + 1), f"bound q={q} b={bucket} n={n}")
                    g = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(g, t, group=group)
                    check(all(torch.equal(g[0], gi) for gi in g), f"replicas q={q} b={bucket} n={n}")
        # ---- uncompressed is exact
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "32"
        for dtype in (torch.float16, torch.float32, torch.bfloat16):
            for n in (1, 1024, 1_000_000):
                t = torch.full((n,), float(rank + 1), dtype=dtype, device=dev)
                dist.all_reduce(t, group=group)
                check(bool((t == world * (world + 1) // 2).all()), f"raw {dtype} n={n}")
        # ---- fused kernel == CPU oracle, bit for bit (random data, compressed, lanes as planned)
        os.environ["CGX_COMPRESSION_BUCKET_SIZE"] = "512"
        be = cgx.get_backend()
        for q, n in ((4, 1 << 20), (8, 300_000), (2, 77_777)):
            os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(q)
            gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
            x = (torch.randn(n, generator=gen) * (rank + 1)).to(dev)
            ins = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(ins, x, group=group)
            y = x.clone()
            dist.all_reduce(y, group=group)
            if rank == 0:
                cpu = [t.cpu() for t in ins]
                lanes = be.last_lanes() if hasattr(be, "last_lanes") else 0
                # messages up to CGX_ONESHOT_MAX_BYTES take the one-shot kernel (one chunk)
                oneshot = n * 4 <= int(os.environ.get("CGX_ONESHOT_MAX_BYTES", 2 << 20)) and world > 1
                if oneshot:
                    C.oneshot_simulate(cpu, [(0, n, q, 512)], max(1, lanes), False, False, False, 0, 1, 4096)
                else:
                    C.sra_simulate(cpu, [(0, n, q, 512)], max(1, lanes), False, False, False, 0, 1, 4096)
                check(torch.equal(y.cpu(), cpu[0]), f"oracle q={q} n={n} oneshot={oneshot}")
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    t = torch.tensor([len(fails)], device=dev, dtype=torch.int64)
    dist.all_reduce(t, group=group)  # int64 -> NCCL delegate
    return {"status": "pass" if t.item() == 0 else "fail", "checks_per_rank": checks,
            "failures_all_ranks": int(t.item()), "first_failures_rank0": fails[:5]}


def main():
    import torch_cgx_b200 as cgx

    rank, world, local = cgx.map_launcher_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    res = run_selftest(dev)
    be = cgx.get_backend()
    res.update({"world": world, "heap": be.heap_kind(), "multicast": be.uses_multicast()})
    if rank == 0:
        print(json.dumps({"selftest": res}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if res["status"] == "pass" else 1


if __name__ == "__main__":
    sys.exit(main())
