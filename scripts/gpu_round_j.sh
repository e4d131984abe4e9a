#!/bin/bash
# 1-GPU final check: GPU test tier, one ncu capture of the flagship kernel, kernel microbench, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/j_pytest.log
echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sra_kernel -s 3 -c 1 -o gpurun_out/f_prof_fused_w1 -f python bench/ncu_target.py > gpurun_out/j_ncu.log 2>&1
timeout 200 python bench/kernel_bench.py --sizes-mb 64 --bits 4,8 --dtypes float32 --out gpurun_out/j_kernel_bench.json > gpurun_out/j_kernel_bench.log 2>&1
timeout 200 python __graft_entry__.py smoke > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/j_smoke.log
tail -4 gpurun_out/j_pytest.log; tail -2 gpurun_out/j_smoke.log
python3 - <<'PY'
import json
d=json.load(open("gpurun_out/j_kernel_bench.json"))
for r in d["rows"]:
    print(r["mb"],r["dtype"],r["bits"],r["bucket"],"q",r["quantize_stream_us"],"d",r["dequantize_stream_us"],"fused",r["fused_w1_us"])
PY
