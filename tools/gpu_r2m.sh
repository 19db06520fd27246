#!/bin/bash
# round-2 run M (1 GPU): compute-sanitizer memcheck over the K-level kernel tests, the shifted solver and smoke(); racecheck over smoke()
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=20 BICG_AUTOTUNE=0
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 420 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --tb=line 2>&1 | cut -c1-300 | tail -25; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/m_memcheck_kernels.log 2>&1
( timeout 300 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | cut -c1-300 | tail -25; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/m_memcheck_smoke.log 2>&1
( timeout 300 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_shifted.py -q -m gpu -x --tb=line -k "not medium" 2>&1 | cut -c1-300 | tail -25; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/m_memcheck_shifted.log 2>&1
( timeout 300 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | cut -c1-300 | tail -30; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/m_racecheck_smoke.log 2>&1
tail -n 8 gpurun_out/m_*.log
