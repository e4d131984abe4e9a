#!/bin/bash
# 2-GPU: full GPU test tier (incl. graph / ring / failure tests) + kernel bench on GPU 0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/d_pytest.log
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 600 python bench/kernel_bench.py --sizes-mb 25,64 --bits 4,8 --out gpurun_out/d_kernel_bench.json > gpurun_out/d_kernel_bench.log 2>&1
timeout 300 python bench/kernel_bench.py --sizes-mb 64 --bits 4 --buckets 64,128,1024 --dtypes float32 --out gpurun_out/d_kernel_bench_buckets.json > gpurun_out/d_kernel_bench_buckets.log 2>&1
tail -12 gpurun_out/d_pytest.log
python3 - <<'PY'
import json
for f in ("gpurun_out/d_kernel_bench.json","gpurun_out/d_kernel_bench_buckets.json"):
    try: d=json.load(open(f))
    except Exception as e: print(f,e); continue
    print(d["clocks"])
    for r in d["rows"]:
        print(r["mb"],r["dtype"],r["bits"],r["bucket"],"q",r["quantize_stream_us"],r["quantize_stream_gbs"],"d",r["dequantize_stream_us"],r["dequantize_stream_gbs"],"fused",r["fused_w1_us"])
PY
