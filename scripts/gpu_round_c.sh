#!/bin/bash
# 8-GPU (or N-GPU) measurement batch: selftest, allreduce sweep (NVLS on/off), phase trace,
# flagship bench (with same-run NCCL arm + allreduce block + selftest).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CGX_LOG_LEVEL=1 timeout 240 $TR --master-port 29521 bench/selftest.py > gpurun_out/c_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/c_selftest.log
timeout 500 $TR --master-port 29522 bench/allreduce_sweep.py --sizes 64,1024,4096,16384,65536,262144 --bits 2,4,8,32 --out gpurun_out/c_sweep_${N}.json > gpurun_out/c_sweep.log 2>&1
CGX_NVLS=0 timeout 300 $TR --master-port 29523 bench/allreduce_sweep.py --sizes 1024,16384,65536,262144 --bits 4,32 --out gpurun_out/c_sweep_${N}_nonvls.json > gpurun_out/c_sweep_nonvls.log 2>&1
timeout 200 $TR --master-port 29524 bench/trace_phases.py --bits 4 --sizes-mb 16,64 --out gpurun_out/c_trace_${N}.json > gpurun_out/c_trace.log 2>&1
CGX_NVLS=0 timeout 200 $TR --master-port 29525 bench/trace_phases.py --bits 4 --sizes-mb 64 --out gpurun_out/c_trace_${N}_nonvls.json > gpurun_out/c_trace_nonvls.log 2>&1
timeout 600 $TR --master-port 29526 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c_bench_${N}.log 2>&1
tail -2 gpurun_out/c_selftest.log
grep -h '"impl"' gpurun_out/c_sweep.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'),'wire',r.get('wire_gbs'))
"
echo "--- no NVLS"
grep -h '"impl": "cgx"' gpurun_out/c_sweep_nonvls.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'),'wire',r.get('wire_gbs'))
"
cat gpurun_out/c_trace.log | tail -4
tail -c 3000 gpurun_out/c_bench_${N}.log
