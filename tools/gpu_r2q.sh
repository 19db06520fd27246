#!/bin/bash
# round-2 run Q (2 GPUs): multi-GPU parity worker with the resident mode on (ghost columns in the 16-bit encoding)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29761 tests/_mgpu_worker.py 2>&1 | tail -45 ) > gpurun_out/q_parity.log 2>&1
tail -n 40 gpurun_out/q_parity.log | cut -c1-200
