#!/bin/bash
# round-2 run P (8 GPUs): resident mode at the benchmark's own geometry -- parity worker, loop times with / without, bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29751 tests/_mgpu_worker.py 2>&1 | tail -45 ) > gpurun_out/p_parity.log 2>&1
( BICG_MEGA_TRACE=1 QP_MODES=mega timeout 120 $TR --master-port 29752 tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab 2>&1 | grep "N=8\|trace r3" | cut -c1-600 ) > gpurun_out/p_perf_resident1.log 2>&1
( BICG_RESIDENT=0 QP_MODES=mega timeout 120 $TR --master-port 29753 tools/quick_perf.py bicgstab 2>&1 | grep "N=8" ) > gpurun_out/p_perf_resident0.log 2>&1
( timeout 300 $TR --master-port 29754 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/p_n8_bench.out 2> gpurun_out/p_n8_bench.err )
grep "^{" gpurun_out/p_n8_bench.out > gpurun_out/p_n8_bench.json
tail -n 6 gpurun_out/p_parity.log; cat gpurun_out/p_perf_*.log; cut -c1-400 gpurun_out/p_n8_bench.json
