#!/usr/bin/env python
"""CIFAR-10/100 data-parallel training with compressed gradients.

Equivalent of the reference's examples/cifar_train.py (ResNet-18, SGD + step LR,
DDP, ``--dist-backend {cgx,nccl,gloo}``, ``--quantization-bits``,
``--quantization-bucket-size``; /root/reference/examples/cifar_train.py:1-239),
but launcher-agnostic: torchrun, mpirun or a plain single process all work.

    torchrun --nproc-per-node 8 examples/cifar_train.py --quantization-bits 4
    python examples/cifar_train.py --synthetic --epochs 1          # no dataset needed

With --synthetic (or when torchvision's CIFAR files are not on disk and there is
no network) random CIFAR-shaped data is used so the script always runs.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.nn.parallel import DistributedDataParallel as DDP

import torch_cgx_b200 as cgx
from torch_cgx_b200.models import resnet18


class Metric:
    """Running average that is averaged across ranks when read (reference: Metric, cifar_train.py:217-229)."""

    def __init__(self, name: str):
        self.name = name
        self.sum = torch.zeros(1)
        self.n = 0

    def update(self, val: torch.Tensor) -> None:
        self.sum += val.detach().cpu().float()
        self.n += 1

    @property
    def avg(self) -> float:
        t = torch.tensor([self.sum.item(), float(self.n)])
        if dist.is_initialized() and dist.get_world_size() > 1:
            dev = torch.device("cuda") if torch.cuda.is_available() and dist.get_backend() != "gloo" else torch.device("cpu")
            t = t.to(dev)
            dist.all_reduce(t)
            t = t.cpu()
        return (t[0] / max(t[1], 1)).item()


def get_loaders(args, rank: int, world: int):
    if not args.synthetic:
        try:
            import torchvision
            import torchvision.transforms as T

            norm = T.Normalize((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616))
            ds_cls = torchvision.datasets.CIFAR100 if args.dataset == "cifar100" else torchvision.datasets.CIFAR10
            train = ds_cls(args.data_dir, train=True, download=False,
                           transform=T.Compose([T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
            test = ds_cls(args.data_dir, train=False, download=False, transform=T.Compose([T.ToTensor(), norm]))
            tsamp = torch.utils.data.distributed.DistributedSampler(train, num_replicas=world, rank=rank)
            return (torch.utils.data.DataLoader(train, batch_size=args.batch_size, sampler=tsamp, num_workers=4, pin_memory=True),
                    torch.utils.data.DataLoader(test, batch_size=args.batch_size, num_workers=4, pin_memory=True), tsamp)
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[cifar_train] dataset unavailable ({e}); falling back to synthetic data")
    ncls = 100 if args.dataset == "cifar100" else 10
    g = torch.Generator().manual_seed(1234 + rank)
    n = args.synthetic_batches * args.batch_size
    x = torch.randn(n, 3, 32, 32, generator=g)
    y = torch.randint(0, ncls, (n,), generator=g)
    ds = torch.utils.data.TensorDataset(x, y)
    mk = lambda: torch.utils.data.DataLoader(ds, batch_size=args.batch_size, pin_memory=torch.cuda.is_available())  # noqa: E731
    return mk(), mk(), None


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", default="cifar10", choices=["cifar10", "cifar100"])
    p.add_argument("--data-dir", default="./data")
    p.add_argument("--log-dir", default="./logs")
    p.add_argument("--batch-size", type=int, default=128, help="per-process batch")
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--base-lr", type=float, default=0.1)
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--wd", type=float, default=5e-4)
    p.add_argument("--dist-backend", default="cgx", choices=["cgx", "nccl", "gloo"])
    p.add_argument("--quantization-bits", type=int, default=4)
    p.add_argument("--quantization-bucket-size", type=int, default=1024)
    p.add_argument("--layer-min-size", type=int, default=1024)
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--synthetic-batches", type=int, default=20)
    p.add_argument("--amp", action="store_true", help="bf16 autocast")
    args = p.parse_args()

    rank, world, local_rank = cgx.map_launcher_env()  # torchrun / mpirun / slurm / single process
    use_cuda = torch.cuda.is_available() and args.dist_backend != "gloo"
    if use_cuda:
        torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    dist.init_process_group(args.dist_backend, init_method="env://", rank=rank, world_size=world)

    ncls = 100 if args.dataset == "cifar100" else 10
    model = resnet18(num_classes=ncls, cifar_stem=True).to(device)
    model = DDP(model, device_ids=[local_rank] if use_cuda else None)
    if args.dist_backend == "cgx":
        state = cgx.CGXState(None, layer_min_size=args.layer_min_size,
                             compression_params={"bits": args.quantization_bits,
                                                 "bucket_size": args.quantization_bucket_size})
        model.register_comm_hook(state, cgx.cgx_hook)
    opt = torch.optim.SGD(model.parameters(), lr=args.base_lr * world ** 0.5, momentum=args.momentum, weight_decay=args.wd)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=max(1, args.epochs // 3), gamma=0.1)
    train_loader, test_loader, sampler = get_loaders(args, rank, world)

    for epoch in range(args.epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        model.train()
        loss_m, acc_m = Metric("loss"), Metric("acc")
        t0 = time.time()
        seen = 0
        for x, y in train_loader:
            x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
            opt.zero_grad(set_to_none=True)
            with torch.autocast(device.type, dtype=torch.bfloat16, enabled=args.amp):
                out = model(x)
                loss = F.cross_entropy(out, y)
            loss.backward()
            opt.step()
            loss_m.update(loss)
            acc_m.update((out.argmax(1) == y).float().mean())
            seen += x.size(0)
        sched.step()
        if use_cuda:
            torch.cuda.synchronize()
        dt = time.time() - t0
        model.eval()
        vacc = Metric("val_acc")
        with torch.no_grad():
            for x, y in test_loader:
                x, y = x.to(device), y.to(device)
                vacc.update((model(x).argmax(1) == y).float().mean())
        tl, ta, va = loss_m.avg, acc_m.avg, vacc.avg
        if rank == 0:
            print(f"epoch {epoch}: loss {tl:.4f} acc {ta:.4f} val_acc {va:.4f} "
                  f"| {seen * world / dt:.0f} img/s ({world} ranks, {args.dist_backend}"
                  f"{', ' + str(args.quantization_bits) + ' bit' if args.dist_backend == 'cgx' else ''})", flush=True)
    if rank == 0 and args.dist_backend == "cgx" and use_cuda:
        st = cgx.get_backend().stats()
        print(f"cgx: {st[0]} allreduces, {st[1]} fused-kernel launches, wire {st[3] / 1e6:.1f} MB vs raw {st[4] / 1e6:.1f} MB")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
