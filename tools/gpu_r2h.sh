#!/bin/bash
# round-2 run H (8 GPUs): parity worker, two switch experiments, bench lines for cfg 4 (T', 8 GPUs) and cfg 5 (random, CA, 8 GPUs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29741 tests/_mgpu_worker.py 2>&1 | tail -50 ) > gpurun_out/h_parity.log 2>&1
( QP_MODES=mega timeout 200 $TR --master-port 29742 tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab 2>&1 | grep "N=8" ) > gpurun_out/h_perf_default.log 2>&1
( BICG_L2_HINT=0 QP_MODES=mega timeout 200 $TR --master-port 29743 tools/quick_perf.py bicgstab 2>&1 | grep "N=8" ) > gpurun_out/h_perf_l2hint0.log 2>&1
( BICG_GATHER_CG=1 BICG_MEGA_TRACE=1 QP_MODES=mega timeout 200 $TR --master-port 29744 tools/quick_perf.py bicgstab 2>&1 | grep "N=8\|trace r3" | cut -c1-600 ) > gpurun_out/h_perf_gathercg.log 2>&1
( BENCH_E2E_VERBOSE=1 timeout 500 $TR --master-port 29745 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/h_n8_bench.out 2> gpurun_out/h_n8_bench.err )
( timeout 600 $TR --master-port 29746 bench.py --gpus 8 --workload random --steps 5 --warmup 3 --no-cpu 2>&1 | grep "^{" ) > gpurun_out/h_n8_bench_random.json 2>&1
grep "^{" gpurun_out/h_n8_bench.out > gpurun_out/h_n8_bench.json
grep "bicg create r0\|bicg entry r0" gpurun_out/h_n8_bench.err | tail -12 > gpurun_out/h_n8_e2e_laps.log
tail -n 40 gpurun_out/h_parity.log; cat gpurun_out/h_perf_*.log; cut -c1-300 gpurun_out/h_n8_bench.json gpurun_out/h_n8_bench_random.json; cat gpurun_out/h_n8_e2e_laps.log
