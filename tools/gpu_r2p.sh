#!/bin/bash
# round-2 run P (8 GPUs): resident mode at the benchmark's own geometry -- bench line (carries the H-level parity at P = 8), then loop times with / without
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 200 $TR --master-port 29754 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/p_n8_bench.out 2> gpurun_out/p_n8_bench.err )
grep "^{" gpurun_out/p_n8_bench.out > gpurun_out/p_n8_bench.json
( BICG_MEGA_TRACE=1 QP_RESIDENT=1,0 timeout 100 $TR --master-port 29752 tools/quick_perf.py bicgstab 2>&1 | grep "N=8\|trace r3" | cut -c1-600 ) > gpurun_out/p_perf_resident_sweep.log 2>&1
cat gpurun_out/p_perf_resident_sweep.log; cut -c1-400 gpurun_out/p_n8_bench.json; tail -3 gpurun_out/p_n8_bench.err | cut -c1-300
