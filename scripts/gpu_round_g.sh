#!/bin/bash
# short N-GPU perf check after a kernel change: selftest + small sweep + phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench/selftest.py > gpurun_out/g_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/g_selftest.log
timeout 400 $TR --master-port 29532 bench/allreduce_sweep.py --sizes 1024,16384,65536 --bits 2,4,8,32 --out gpurun_out/g_sweep_${N}.json > gpurun_out/g_sweep.log 2>&1
timeout 200 $TR --master-port 29534 bench/trace_phases.py --bits 4 --sizes-mb 64 --out gpurun_out/g_trace_${N}.json > gpurun_out/g_trace.log 2>&1
tail -2 gpurun_out/g_selftest.log
grep -h '"impl"' gpurun_out/g_sweep.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'))
"
tail -3 gpurun_out/g_trace.log
