#!/bin/bash
# Memory-checking tier (the reference has none): a subset of the GPU kernel tests
# under compute-sanitizer memcheck. The fused kernel runs with W virtual ranks on ONE GPU, so every
# global / "peer" access, the TMA bulk copies and the mbarrier of the item cache are covered.
# (racecheck only tracks shared memory; the kernels synchronise through global flags, and the one
# shared structure -- the item cache -- is written by TMA before a barrier and read-only afterwards.)
set -e
cd "$(dirname "$0")/.."
SUBSET='test_fused_sra_tiny_and_exact_constant or (test_fused_sra_matches_cpu_oracle and 4-dtype0) or test_fused_sra_mixed_layers_unaligned or (test_quantize_kernel_bytes_match_cpu and dtype0-4-512) or (test_oneshot_kernel_matches_cpu_oracle and 2-dtype0) or test_unaligned_layers_and_mixed_config'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SUBSET"
