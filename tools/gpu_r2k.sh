#!/bin/bash
# round-2 run K (1 GPU): the whole GPU suite, no early stop
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 1300 python -m pytest tests -q -m gpu --tb=short 2>&1 | cut -c1-400 | tail -40 ) > gpurun_out/k_tests.log 2>&1
tail -12 gpurun_out/k_tests.log
