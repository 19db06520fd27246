#!/bin/bash
# round-2 run G (1 GPU): the tests that failed in run F, with their full output
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_shifted.py "tests/test_gpu_parity.py::test_random_block_parity" "tests/test_gpu_parity.py::test_bench_matrix_parity" -q -m gpu --tb=short 2>&1 | cut -c1-600 ) > gpurun_out/g_tests.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/g_tests.log | cut -c1-400 | head -80
