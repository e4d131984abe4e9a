#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench/virtual_world.py --world 8 --bits 4,8 --out gpurun_out/v_virtual8.json 2>&1 | tail -5
timeout 300 python bench/virtual_world.py --world 8 --bits 4 --bucket 64 --out gpurun_out/v_virtual8_b64.json 2>&1 | tail -3
