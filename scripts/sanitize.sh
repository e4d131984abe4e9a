#!/bin/bash
# Race / memory checking tier (the reference has none): run a subset of the GPU
# kernel tests under compute-sanitizer. memcheck covers every global/peer access of the fused
# kernel (W virtual ranks on one GPU); racecheck covers the shared-memory path (CGX_KERNEL=block).
set -e
cd "$(dirname "$0")/.."
SUBSET='test_fused_sra_tiny_and_exact_constant or (test_fused_sra_matches_cpu_oracle and dtype0) or (test_quantize_kernel_bytes_match_cpu and dtype0-4-512) or (test_oneshot_kernel_matches_cpu_oracle and dtype0)'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SUBSET"
CGX_KERNEL=block compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "(test_fused_sra_matches_cpu_oracle and dtype0)"
