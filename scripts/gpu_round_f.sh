#!/bin/bash
# 1-GPU check after a kernel change: GPU test tier, kernel microbench, 8-virtual-rank proxy, one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/f_pytest.log
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python bench/kernel_bench.py --sizes-mb 25,64 --bits 2,4,8 --out gpurun_out/f_kernel_bench.json > gpurun_out/f_kernel_bench.log 2>&1
timeout 300 python bench/kernel_bench.py --sizes-mb 64 --bits 4 --buckets 64,128,1024 --dtypes float32 --out gpurun_out/f_kernel_bench_buckets.json > gpurun_out/f_kernel_bench_buckets.log 2>&1
timeout 300 python bench/virtual_world.py --world 8 --bits 4,8 --out gpurun_out/f_virtual8.json > gpurun_out/f_virtual8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sra_kernel -s 3 -c 1 -o gpurun_out/f_prof_fused_w1 -f python bench/ncu_target.py > gpurun_out/f_ncu.log 2>&1
tail -6 gpurun_out/f_pytest.log
tail -6 gpurun_out/f_virtual8.log | cut -c1-300
python3 - <<'PY'
import json
for f in ("gpurun_out/f_kernel_bench.json","gpurun_out/f_kernel_bench_buckets.json"):
    try: d=json.load(open(f))
    except Exception as e: print(f,e); continue
    print(d["clocks"])
    for r in d["rows"]:
        print(r["mb"],r["dtype"],r["bits"],r["bucket"],"q",r["quantize_stream_us"],r["quantize_stream_gbs"],"d",r["dequantize_stream_us"],r["dequantize_stream_gbs"],"fused",r["fused_w1_us"])
PY
