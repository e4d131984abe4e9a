#!/usr/bin/env python
"""Per-phase device timeline of the fused allreduce kernel (tracing subsystem):
every lane (CTA) stamps the device globaltimer at its phase boundaries; this
script prints, per message size, where the time goes:
  A      quantize my copies of the peers' chunks and push them            (t1 - t0)
  fenceA system-scope release of those stores + flag stores (warp 0)      (t6 - t1)
  waitB  wait for the W-1 incoming copies of my chunk (includes fenceA)   (t2 - t1)
  B      dequantize-accumulate, requantize, push to all peers             (t3 - t2)
  fenceB release of the phase-B stores                                    (t7 - t3)
  C      wait for + dequantize the peers' reduced chunks                  (t5 - t3)
torchrun --nproc-per-node N bench/trace_phases.py [--bits 4] [--sizes-mb 1,16,64]
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import torch_cgx_b200 as cgx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--sizes-mb", default="1,16,64")
    ap.add_argument("--series", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/trace.json")
    args = ap.parse_args()
    rank, world, local = cgx.map_launcher_env()
    torch.cuda.set_device(local)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = str(args.bits)
    be = cgx.get_backend()
    be.init_cuda(local)
    rows = []
    for mb in [float(x) for x in args.sizes_mb.split(",")]:
        n = int(mb * (1 << 20)) // 4
        xs = [torch.randn(n, device="cuda") for _ in range(4)]
        for x in xs:
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        # the traced call is the last of a back-to-back series: ranks are then paced by each other
        # (steady state), not by the skew of their first launch after a host barrier
        be.enable_trace(True)
        for i in range(args.series):
            dist.all_reduce(xs[i % len(xs)])
        t = be.read_trace().double()
        be.enable_trace(False)
        t0 = t[:, 0].min()
        d = {
            "mb": mb, "rank": rank, "lanes": t.shape[0],
            "A_mean": ((t[:, 1] - t[:, 0]).mean() / 1e3).item(), "A_max": ((t[:, 1] - t0).max() / 1e3).item(),
            "waitB_mean": ((t[:, 2] - t[:, 1]).clamp(min=0).mean() / 1e3).item(),
            "fenceA_mean": ((t[:, 6] - t[:, 1]).clamp(min=0).mean() / 1e3).item(),
            "B_mean": ((t[:, 3] - t[:, 2]).clamp(min=0).mean() / 1e3).item(),
            "fenceB_mean": ((t[:, 7] - t[:, 3]).clamp(min=0).mean() / 1e3).item(),
            "B_done_max": ((t[:, 3] - t0).max() / 1e3).item(),
            "C_mean": ((t[:, 5] - t[:, 3]).mean() / 1e3).item(),
            "lastwaitC_mean": ((t[:, 4] - t[:, 3]).clamp(min=0).mean() / 1e3).item(),
            "total": ((t[:, 5].max() - t0) / 1e3).item(),
        }
        gathered = [None] * world
        dist.all_gather_object(gathered, d)
        if rank == 0:
            for g in gathered[:2]:
                print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in g.items()}), flush=True)
            rows.extend(gathered)
    if rank == 0:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps({"world": world, "bits": args.bits, "unit": "us", "rows": rows}, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
