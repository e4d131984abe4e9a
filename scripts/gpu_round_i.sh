#!/bin/bash
# 4-GPU: full GPU test tier (incl. the 2x2 hierarchical tests), selftest, sweep, phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -40 > gpurun_out/i_pytest_${N}.log
echo "pytest rc=$?" >> gpurun_out/i_pytest_${N}.log
timeout 200 $TR --master-port 29551 bench/selftest.py > gpurun_out/i_selftest_${N}.log 2>&1
echo "selftest rc=$?" >> gpurun_out/i_selftest_${N}.log
timeout 400 $TR --master-port 29552 bench/allreduce_sweep.py --sizes 1024,16384,65536,262144 --bits 2,4,8,32 --out gpurun_out/i_sweep_${N}.json > gpurun_out/i_sweep_${N}.log 2>&1
timeout 200 $TR --master-port 29553 bench/trace_phases.py --bits 4 --sizes-mb 64 --out gpurun_out/i_trace_${N}.json > gpurun_out/i_trace_${N}.log 2>&1
tail -8 gpurun_out/i_pytest_${N}.log
tail -2 gpurun_out/i_selftest_${N}.log | cut -c1-300
grep -h '"impl"' gpurun_out/i_sweep_${N}.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'))
"
tail -3 gpurun_out/i_trace_${N}.log
