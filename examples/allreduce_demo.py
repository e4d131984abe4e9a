#!/usr/bin/env python
"""Smallest possible use of the backend without DDP:
    torchrun --nproc-per-node 2 examples/allreduce_demo.py
Each rank contributes a random vector; the 4-bit compressed sum is compared with the exact one."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import torch_cgx_b200 as cgx

rank, world, local = cgx.map_launcher_env()
use_cuda = torch.cuda.is_available()
if use_cuda:
    torch.cuda.set_device(local)
dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
torch.manual_seed(rank)
x = torch.randn(1 << 20, device=dev)
exact = x.clone()
dist.all_reduce(exact)                                   # CGX_COMPRESSION_QUANTIZATION_BITS unset: exact sum
for bits in (8, 4, 2):
    y = x.clone()
    cgx.all_reduce(y, bits=bits, bucket_size=512)        # quantized for this call only
    err = ((y - exact).norm() / exact.norm()).item()
    if rank == 0:
        print(f"{bits}-bit allreduce over {world} ranks: relative L2 error {err:.4f}")
dist.barrier()
dist.destroy_process_group()
