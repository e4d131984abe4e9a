#!/bin/bash
# N-GPU validation batch (N = number of visible GPUs): multicast probe + selftest, GPU test tier,
# short allreduce sweep with and without NVLS.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" > gpurun_out/b_info.txt
nvidia-smi topo -m >> gpurun_out/b_info.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CGX_LOG_LEVEL=1 timeout 300 $TR --master-port 29511 bench/selftest.py > gpurun_out/b_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/b_selftest.log
CGX_NVLS=0 CGX_LOG_LEVEL=1 timeout 300 $TR --master-port 29512 bench/selftest.py > gpurun_out/b_selftest_nonvls.log 2>&1
echo "selftest(no nvls) rc=$?" >> gpurun_out/b_selftest_nonvls.log
CGX_VMM=0 CGX_LOG_LEVEL=1 timeout 300 $TR --master-port 29513 bench/selftest.py > gpurun_out/b_selftest_ipc.log 2>&1
echo "selftest(cudaIpc) rc=$?" >> gpurun_out/b_selftest_ipc.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/b_pytest.log
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 600 $TR --master-port 29514 bench/allreduce_sweep.py --sizes 64,1024,4096,16384,65536,262144 --bits 4,8,32 --out gpurun_out/b_sweep_${N}.json > gpurun_out/b_sweep.log 2>&1
CGX_NVLS=0 timeout 600 $TR --master-port 29515 bench/allreduce_sweep.py --sizes 1024,16384,65536,262144 --bits 4,32 --out gpurun_out/b_sweep_${N}_nonvls.json > gpurun_out/b_sweep_nonvls.log 2>&1
NCCL_DEBUG=INFO timeout 120 $TR --master-port 29516 bench/allreduce_sweep.py --sizes 1024 --bits 32 --iters 2 --out gpurun_out/b_tmp.json 2>&1 | grep -i -E "nvls|multicast|P2P/|channels" | head -20 > gpurun_out/b_nccl_nvls.txt
tail -3 gpurun_out/b_selftest.log; tail -2 gpurun_out/b_selftest_nonvls.log; tail -2 gpurun_out/b_selftest_ipc.log; tail -4 gpurun_out/b_pytest.log
grep -h '"impl": "cgx"' gpurun_out/b_sweep.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB bits',r['bits'],r['time_us'],'us x',r['speedup_vs_nccl'],'wire',r.get('wire_gbs'))
"
