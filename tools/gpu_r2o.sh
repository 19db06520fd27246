#!/bin/bash
# round-2 run O (1 GPU): resident mode of the persistent kernel -- whole GPU suite + loop times at the per-GPU size of the 8-GPU benchmark
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | cut -c1-400 | tail -40 ) > gpurun_out/o_tests.log 2>&1
for res in 1 0; do
  ( BICG_RESIDENT=$res BICG_MEGA_TRACE=1 BICG_VERBOSE=1 QP_G=58 QP_MODES=mega timeout 200 python tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab 2>&1 | grep -v "^\[bicg create\|^\[bicg entry" | tail -12 | cut -c1-420 ) > gpurun_out/o_perf_g58_resident$res.log 2>&1
done
tail -15 gpurun_out/o_tests.log; tail -n 8 gpurun_out/o_perf_g58_resident*.log
