#!/usr/bin/env python
"""NVLink traffic of the fused allreduce, measured by the GPU's own link counters
(`nvidia-smi nvlink -gt d`: per-link data tx/rx KiB) around K back-to-back calls, against the
algorithmic packed bytes the engine accounts for (`stats()[3]`: bytes this rank pushed).
    torchrun --nproc-per-node N bench/nvlink_bytes.py --mb 64 --bits 4 --calls 50
"""
import argparse
import json
import os
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import torch_cgx_b200 as cgx  # noqa: E402


def counters(index):
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
    except Exception:  # noqa: BLE001
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return (tx * 1024, rx * 1024) if (tx or rx) else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--bits", default="4,32")
    ap.add_argument("--calls", type=int, default=50)
    ap.add_argument("--out", default="gpurun_out/nvlink_bytes.json")
    args = ap.parse_args()
    rank, world, local = cgx.map_launcher_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    be = cgx.get_backend()
    n = (args.mb << 20) // 4
    x = torch.randn(n, device=dev)
    rows = []
    for bits in [int(b) for b in args.bits.split(",")]:
        os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(bits)
        for _ in range(3):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        c0 = counters(local)
        be.reset_stats()
        for _ in range(args.calls):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        c1 = counters(local)
        st = be.stats()
        row = {"rank": rank, "bits": bits, "mb": args.mb, "calls": args.calls, "nvls": be.uses_multicast(),
               "algorithmic_pushed_bytes_per_call": st[3] // max(1, st[0])}
        if c0 and c1:
            row["nvlink_tx_bytes_per_call"] = (c1[0] - c0[0]) // args.calls
            row["nvlink_rx_bytes_per_call"] = (c1[1] - c0[1]) // args.calls
        else:
            row["nvlink_counters"] = "unavailable"
        rows.append(row)
    allrows = [None] * world
    dist.all_gather_object(allrows, rows)
    if rank == 0:
        flat = [r for rs in allrows for r in rs]
        for r in flat[: 2 * len(rows)]:
            print(json.dumps(r), flush=True)
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps({"world": world, "rows": flat,
                                              "note": "with NVLS the phase-B result leaves each GPU ONCE (multicast) instead of W-1 times; tx counts what leaves this GPU, rx what arrives"}, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
