#!/bin/bash
# N-GPU: GPU test tier, then the pipeline-depth comparison (CGX_STAGES=1/2/4/auto) on a few sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/h_pytest.log
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
fi
port=29540
for st in ${STAGES_LIST:-1 2 4 0}; do
  port=$((port+1))
  CGX_STAGES=$st timeout 300 $TR --master-port $port bench/allreduce_sweep.py --sizes ${SIZES:-16384,65536,262144} --bits ${BITS:-4,32} --out gpurun_out/h_sweep_${N}_st${st}.json > gpurun_out/h_sweep_st${st}.log 2>&1
  echo "== CGX_STAGES=$st"
  grep -h '"impl"' gpurun_out/h_sweep_st${st}.log | python3 -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['bytes']>>10,'KB',r['impl'],r['bits'],r['time_us'],'us x',r.get('speedup_vs_nccl'))
"
done
