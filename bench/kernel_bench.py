#!/usr/bin/env python
"""Single-GPU kernel microbenchmarks: the kernels ALONE (pre-planned, pre-allocated), device-timed
with CUDA events, L2 flushed between iterations, clocks sampled during the run.

  quantize / dequantize   standalone item kernels (_C.PreparedCodec)
  fused_w1                the fused SRA kernel in world=1 mode (load -> min/max -> quantize -> pack
                          -> self-decode): what phase B does per element, minus the peers
reported as achieved HBM GB/s (algorithmic bytes / time) against MEASURED_PEAKS.json."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import torch_cgx_b200 as cgx  # noqa: E402
from torch_cgx_b200.utils.clocks import ClockSampler  # noqa: E402

C = cgx._C


def time_op(fn, flush, iters=10, warmup=3):
    """ONE launch per measurement, L2 flushed before it: includes ~5 us of launch + event overhead."""
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
    return ts[len(ts) // 2], ts[0]


def time_stream(fns, reps=3):
    """Steady state: the launches of `fns` (each on its OWN buffers, > 2x L2 in total, so nothing is
    cache resident) back to back between two events; per-launch time, median of `reps`."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in fns:
            f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / len(fns))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernel_bench.json")
    ap.add_argument("--sizes-mb", default="1,8,25,64,256")
    ap.add_argument("--lanes", type=int, default=296)
    ap.add_argument("--bits", default="2,4,8")
    ap.add_argument("--buckets", default="512")
    ap.add_argument("--dtypes", default="float32,bfloat16")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    sampler = ClockSampler(0).start()
    for mb in [int(x) for x in args.sizes_mb.split(",")]:
        for dname in args.dtypes.split(","):
            dtype = getattr(torch, dname)
            for bucket in [int(b) for b in args.buckets.split(",")]:
                for bits in [int(b) for b in args.bits.split(",")]:
                    n = (mb << 20) // (4 if dtype == torch.float32 else 2)
                    x = torch.randn(n, device=dev).to(dtype)
                    layers = [(0, n, bits, bucket)]
                    es = x.element_size()
                    codec = C.PreparedCodec(layers, dtype, 0, False)
                    wire = torch.zeros(codec.wire_bytes(), dtype=torch.uint8, device=dev)
                    out = torch.empty_like(x)
                    wb = cgx.ops.wire_bytes(n, bits, bucket, es)
                    tq, tq_min = time_op(lambda: codec.quantize(x, wire), flush)
                    td, td_min = time_op(lambda: codec.dequantize(wire, out), flush)
                    # steady state over rotating buffer sets (> 2x L2 in total)
                    nset = max(3, (300 << 20) // (n * es) + 1)
                    xs = [torch.randn(n, device=dev).to(dtype) for _ in range(nset)]
                    ws = [torch.zeros(codec.wire_bytes(), dtype=torch.uint8, device=dev) for _ in range(nset)]
                    os_ = [torch.empty_like(x) for _ in range(nset)]
                    for a, w in zip(xs, ws):
                        codec.quantize(a, w)
                    reps = max(1, 24 // nset)
                    tqs = time_stream([(lambda a=a, w=w: codec.quantize(a, w)) for a, w in zip(xs, ws)] * reps)
                    tds = time_stream([(lambda o=o, w=w: codec.dequantize(w, o)) for o, w in zip(os_, ws)] * reps)
                    del xs, ws, os_
                    g = C.LocalSraGroup(1, args.lanes, max(64 << 20, n * es + (1 << 20)), 5000, 4096)
                    y = x.clone()
                    g.allreduce([y], layers)  # plan + upload outside the timed region
                    tf, tf_min = time_op(lambda: g.allreduce([y], layers), flush)
                    row = {
                        "mb": mb, "dtype": dname, "bits": bits, "bucket": bucket,
                        "quantize_stream_us": round(tqs * 1e3, 1),
                        "quantize_stream_gbs": round((n * es + wb) / tqs / 1e6, 1),
                        "quantize_stream_frac_of_measured_hbm": round((n * es + wb) / tqs / 1e6 / hbm, 3),
                        "dequantize_stream_us": round(tds * 1e3, 1),
                        "dequantize_stream_gbs": round((n * es + wb) / tds / 1e6, 1),
                        "dequantize_stream_frac_of_measured_hbm": round((n * es + wb) / tds / 1e6 / hbm, 3),
                        "quantize_single_launch_us": round(tq * 1e3, 1), "dequantize_single_launch_us": round(td * 1e3, 1),
                        "fused_w1_us": round(tf * 1e3, 1), "fused_w1_min_us": round(tf_min * 1e3, 1),
                        "fused_w1_gbs": round((2 * n * es) / tf / 1e6, 1),
                        "fused_w1_frac_of_measured_hbm": round((2 * n * es) / tf / 1e6 / hbm, 3),
                    }
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                    del g, codec
    clocks = sampler.stop()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({
        "hbm_gbs_measured": hbm, "clocks": clocks,
        "timing": "*_stream_*: CUDA events around back-to-back launches over rotating buffer sets (> 2x L2 in total), per launch; *_single_launch_* and fused_w1: events around ONE launch after a 256 MB L2 flush (includes ~5 us launch + event overhead)",
        "rows": rows}, indent=1))
    print(json.dumps({"clocks": clocks}))


if __name__ == "__main__":
    main()
