#!/bin/bash
# Reference: examples/run_cifar.sh (mpirun -np N python cifar_train.py --quantization-bits 8 ...).
# Any launcher works here; torchrun shown. mpirun -np $N python examples/cifar_train.py ... works too.
N=${1:-2}
torchrun --standalone --nnodes=1 --nproc-per-node "$N" --local-addr 127.0.0.1 \
  "$(dirname "$0")/cifar_train.py" --epochs 10 --quantization-bits 8 --quantization-bucket-size 1024 \
  --dist-backend cgx "${@:2}"
