#!/bin/bash
# round-2 run B (N GPUs, default 2): multi-GPU parity worker, loop timings with trace, short bench
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
nvidia-smi -L > gpurun_out/b${N}_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 900 $TR --master-port 29711 tests/_mgpu_worker.py 2>&1 | tail -60 ) > gpurun_out/b${N}_parity.log 2>&1
( BICG_MEGA_TRACE=1 timeout 400 $TR --master-port 29712 tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab pipe_bicgstab_rr 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -40 ) > gpurun_out/b${N}_perf.log 2>&1
( timeout 600 $TR --master-port 29713 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -4 ) > gpurun_out/b${N}_bench.log 2>&1
tail -n 45 gpurun_out/b${N}_parity.log gpurun_out/b${N}_perf.log gpurun_out/b${N}_bench.log
