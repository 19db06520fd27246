#!/bin/bash
# round-2 run C (1 GPU): whole gpu test-suite, tuning sweeps, ncu launch list + one full capture of the persistent kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 ) > gpurun_out/c_tests.log 2>&1
for rw in 216 600 1000 1400; do
  ( BICG_ROW_WEIGHT=$rw BICG_MEGA_TRACE=1 QP_MODES=mega timeout 200 python tools/quick_perf.py bicgstab 2>&1 | tail -3 ) > gpurun_out/c_roww_$rw.log 2>&1
done
( BICG_MEGA_TRACE=1 QP_MODES=mega timeout 300 python tools/quick_perf.py ca_bicgstab pipe_bicgstab pipe_bicgstab_rr 2>&1 | tail -6 ) > gpurun_out/c_perf_methods.log 2>&1
( QP_KIND=laplace5 QP_G=2000 QP_P0=0 timeout 300 python tools/quick_perf.py pipe_bicgstab 2>&1 | tail -3 ) > gpurun_out/c_perf_laplace.log 2>&1
for ln in 4 8 32; do
  ( BICG_MEGA=2 BICG_MEGA_LANES=$ln QP_MODES=mega QP_KIND=random QP_G=2000000 QP_P0=32 QP_ITERS=60 timeout 300 python tools/quick_perf.py ca_bicgstab 2>&1 | tail -2 ) > gpurun_out/c_random_lanes_$ln.log 2>&1
done
( QP_MODES=graph QP_KIND=random QP_G=2000000 QP_P0=32 QP_ITERS=60 timeout 300 python tools/quick_perf.py ca_bicgstab 2>&1 | tail -2 ) > gpurun_out/c_random_graph.log 2>&1
# ncu: launch list of the bench command, then one full capture of the persistent kernel (one launch = one solve)
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/c_ncu_bench.log 2>&1 )
( BICG_MAX_ITER=40 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bicg_mega_kernel -c 1 -o gpurun_out/c_mega_full python tools/quick_perf_one.py > gpurun_out/c_ncu_full.log 2>&1 )
tail -n 12 gpurun_out/c_tests.log gpurun_out/c_roww_*.log gpurun_out/c_perf_methods.log gpurun_out/c_perf_laplace.log gpurun_out/c_random_*.log
tail -n 3 gpurun_out/c_ncu_bench.log gpurun_out/c_ncu_full.log; ls -la gpurun_out/c_mega_full* gpurun_out/c_launches.csv
