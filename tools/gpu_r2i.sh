#!/bin/bash
# round-2 run I (1 GPU): re-run of the adjusted tests, shifted-solver throughput
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_shifted.py "tests/test_gpu_parity.py::test_random_block_parity" "tests/test_gpu_parity.py::test_bench_matrix_parity" -q -m gpu --tb=short 2>&1 | cut -c1-500 | tail -30 ) > gpurun_out/i_tests.log 2>&1
( timeout 600 python tools/shifted_perf.py 16 64 256 2>&1 | tail -5 ) > gpurun_out/i_shifted_perf.log 2>&1
cat gpurun_out/i_tests.log | tail -12; cat gpurun_out/i_shifted_perf.log
