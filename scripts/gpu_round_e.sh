#!/bin/bash
# Final N-GPU measurement batch (N = visible GPUs): full sweep, flagship + other BASELINE models,
# small-bucket sweep, NVLink counters, phase trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29541 bench/allreduce_sweep.py --min-kb 1 --max-mb 1024 --bits 2,4,8,32 --iters 12 --out gpurun_out/e_sweep_${N}.json > gpurun_out/e_sweep.log 2>&1
timeout 400 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench_resnet50_${N}.log 2>&1
timeout 400 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --model gpt2-medium --no-allreduce --no-selftest > gpurun_out/e_bench_gpt2m_${N}.log 2>&1
timeout 400 $TR --master-port 29544 bench.py --gpus $N --steps 15 --warmup 3 --model gpt2-medium --batch 1 --no-allreduce --no-selftest > gpurun_out/e_bench_gpt2m_b1_${N}.log 2>&1
timeout 400 $TR --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 --model vit-l16 --no-allreduce --no-selftest > gpurun_out/e_bench_vitl_${N}.log 2>&1
for B in 64 128; do
  timeout 200 $TR --master-port 29546 bench/allreduce_sweep.py --sizes 65536 --bits 4 --bucket-size $B --iters 12 --out gpurun_out/e_sweep_${N}_bucket$B.json > gpurun_out/e_sweep_bucket$B.log 2>&1
done
CGX_LANES=148 timeout 200 $TR --master-port 29547 bench/allreduce_sweep.py --sizes 16384,65536 --bits 4 --iters 12 --out gpurun_out/e_sweep_${N}_lanes148.json > gpurun_out/e_sweep_lanes148.log 2>&1
timeout 200 $TR --master-port 29548 bench/nvlink_bytes.py --mb 64 --bits 4,32 --calls 40 --out gpurun_out/e_nvlink_${N}.json > gpurun_out/e_nvlink.log 2>&1
CGX_NVLS=0 timeout 200 $TR --master-port 29549 bench/nvlink_bytes.py --mb 64 --bits 4 --calls 40 --out gpurun_out/e_nvlink_${N}_nonvls.json > gpurun_out/e_nvlink_nonvls.log 2>&1
timeout 200 $TR --master-port 29550 bench/trace_phases.py --bits 4 --sizes-mb 64,64 --out gpurun_out/e_trace_${N}.json > gpurun_out/e_trace.log 2>&1
grep -h '"impl"' gpurun_out/e_sweep.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    if r['bytes'] in (1<<20,16<<20,64<<20,256<<20,1<<30): print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'))
"
for f in gpurun_out/e_bench_*_${N}.log; do echo "== $f"; grep -h '^{"metric"' $f | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); b=r.get('baseline') or {}
    print(r['metric'], r['value'], r['unit'], 'ms/step', r['ms_per_step'], 'nccl', b.get('value'), 'x', r.get('vs_baseline'), 'e2e x', r.get('e2e_vs_baseline'), 'launches', r['gpu_launches'], r['clocks'])
"; done
grep -h '"impl": "cgx"' gpurun_out/e_sweep_bucket*.log gpurun_out/e_sweep_lanes148.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'))
"
tail -4 gpurun_out/e_nvlink.log; tail -2 gpurun_out/e_nvlink_nonvls.log; tail -3 gpurun_out/e_trace.log
