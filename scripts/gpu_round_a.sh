#!/bin/bash
# 1-GPU validation batch: GPU test tier, kernel microbench, one ncu capture of the fused kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/a_pytest.log
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench/kernel_bench.py --sizes-mb 25,64 --bits 4,8 --out gpurun_out/a_kernel_bench.json > gpurun_out/a_kernel_bench.log 2>&1
timeout 300 python bench/kernel_bench.py --sizes-mb 64 --bits 4 --buckets 64,128,1024 --dtypes float32 --out gpurun_out/a_kernel_bench_buckets.json > gpurun_out/a_kernel_bench_buckets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sra_kernel -s 3 -c 1 -o gpurun_out/a_prof_fused_w1 -f python bench/ncu_target.py > gpurun_out/a_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:quantize_items -s 2 -c 2 -o gpurun_out/a_prof_quantize -f python bench/ncu_target.py --op quantize > gpurun_out/a_ncu_q.log 2>&1
timeout 500 bash scripts/sanitize.sh > gpurun_out/a_sanitize.log 2>&1; echo "sanitize rc=$?" >> gpurun_out/a_sanitize.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -5 gpurun_out/a_pytest.log
tail -3 gpurun_out/a_sanitize.log; tail -2 gpurun_out/a_smoke.log; tail -4 gpurun_out/a_kernel_bench.log | cut -c1-400
