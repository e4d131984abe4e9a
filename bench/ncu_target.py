#!/usr/bin/env python
"""Tiny target for `ncu`: a handful of launches of ONE kernel on one GPU (ncu replays each
captured launch ~40 times, so keep it short). --op fused|quantize|dequantize."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import torch_cgx_b200 as cgx  # noqa: E402

C = cgx._C
ap = argparse.ArgumentParser()
ap.add_argument("--op", default="fused")
ap.add_argument("--mb", type=int, default=64)
ap.add_argument("--bits", type=int, default=4)
ap.add_argument("--bucket", type=int, default=512)
ap.add_argument("--lanes", type=int, default=296)
args = ap.parse_args()
dev = torch.device("cuda", 0)
n = (args.mb << 20) // 4
x = torch.randn(n, device=dev)
layers = [(0, n, args.bits, args.bucket)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if args.op == "fused":
    g = C.LocalSraGroup(1, args.lanes, n * 4 + (1 << 20), 5000, 4096)
    for _ in range(6):
        flush.zero_()
        g.allreduce([x], layers)
else:
    codec = C.PreparedCodec(layers, torch.float32, 0, False)
    wire = torch.zeros(codec.wire_bytes(), dtype=torch.uint8, device=dev)
    out = torch.empty_like(x)
    for _ in range(6):
        flush.zero_()
        codec.quantize(x, wire)
        codec.dequantize(wire, out)
torch.cuda.synchronize()
print("done")
