#!/bin/bash
# round-2 run E (8 GPUs): boundary-weight sweep of the multi-GPU BiCGStab loop, N = 4 and N = 8 timings, bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for bw in 600 2000 4000; do
  ( BICG_BOUNDARY_WEIGHT=$bw BICG_MEGA_TRACE=1 QP_MODES=mega timeout 200 $TR --nproc-per-node 8 --master-port 2972$((bw/1000)) tools/quick_perf.py bicgstab 2>&1 | grep "snap r0\|snap r3\|N=8\|trace r3" | cut -c1-600 | tail -6 ) > gpurun_out/e_bw_$bw.log 2>&1
done
( QP_MODES=mega timeout 200 $TR --nproc-per-node 4 --master-port 29731 tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab 2>&1 | grep "N=4" ) > gpurun_out/e_n4_perf.log 2>&1
( timeout 400 $TR --nproc-per-node 8 --master-port 29732 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | grep "^{" ) > gpurun_out/e_n8_bench.json 2>&1
( timeout 400 $TR --nproc-per-node 4 --master-port 29733 bench.py --gpus 4 --steps 10 --warmup 3 2>&1 | grep "^{" ) > gpurun_out/e_n4_bench.json 2>&1
tail -n 8 gpurun_out/e_bw_*.log gpurun_out/e_n4_perf.log; cut -c1-400 gpurun_out/e_n8_bench.json gpurun_out/e_n4_bench.json
