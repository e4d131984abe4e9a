#!/usr/bin/env python
"""Single-GPU kernel microbenchmarks (device-timed with CUDA events, L2 flushed
between iterations): standalone quantize / dequantize and the fused kernel in
world=1 mode (load -> min/max -> quantize -> pack -> self-decode), reported as
achieved HBM GB/s against MEASURED_PEAKS.json."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import torch_cgx_b200 as cgx  # noqa: E402

C = cgx._C


def time_op(fn, flush, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()  # > L2 sized write evicts the working set
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernel_bench.json")
    ap.add_argument("--sizes-mb", default="1,8,25,64,256")
    ap.add_argument("--lanes", type=int, default=148)
    ap.add_argument("--bits", default="2,4,8")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for mb in [int(x) for x in args.sizes_mb.split(",")]:
        for dtype in (torch.float32, torch.bfloat16):
            for bits in [int(b) for b in args.bits.split(",")]:
                n = (mb << 20) // (4 if dtype == torch.float32 else 2)
                x = torch.randn(n, device=dev).to(dtype)
                layers = [(0, n, bits, 512)]
                es = x.element_size()
                w = C.quantize(x, layers, 1, 1, False, 1.0, False, 0, 0, 0, 0, 2048)
                wire = cgx.ops.wire_bytes(n, bits, 512, es)
                tq = time_op(lambda: C.quantize(x, layers, 1, 1, False, 1.0, False, 0, 0, 0, 0, 2048), flush)
                td = time_op(lambda: C.dequantize(w, x, layers, 1, 1, False, 2048), flush)
                g = C.LocalSraGroup(1, args.lanes, max(64 << 20, n * es + (1 << 20)), 5000, 2048)
                y = x.clone()
                tf = time_op(lambda: g.allreduce([y], layers), flush)
                row = {
                    "mb": mb, "dtype": str(dtype).split(".")[-1], "bits": bits,
                    "quantize_ms": round(tq, 4), "quantize_gbs": round((n * es + wire) / tq / 1e6, 1),
                    "dequantize_ms": round(td, 4), "dequantize_gbs": round((n * es + wire) / td / 1e6, 1),
                    "fused_w1_ms": round(tf, 4), "fused_w1_gbs": round((2 * n * es) / tf / 1e6, 1),
                    "fused_w1_frac_of_measured_hbm": round((2 * n * es) / tf / 1e6 / hbm, 3),
                }
                rows.append(row)
                print(json.dumps(row), flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({"hbm_gbs_measured": hbm, "note": "quantize/dequantize timings include the op's own output allocation (torch empty/zeros)", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
