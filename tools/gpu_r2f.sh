#!/bin/bash
# round-2 run F (1 GPU): the whole gpu test-suite, bench lines for cfg 2 and cfg 3, ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/f_tests.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "^{" ) > gpurun_out/f_bench_transport.json 2>gpurun_out/f_bench_transport.err
( timeout 600 python bench.py --workload laplace --steps 5 --warmup 3 --no-cpu 2>&1 | grep "^{" ) > gpurun_out/f_bench_laplace.json 2>&1
( timeout 600 python bench.py --workload random --steps 5 --warmup 3 --no-cpu 2>&1 | grep "^{" ) > gpurun_out/f_bench_random_n1.json 2>&1
( BICG_AUTOTUNE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/f_ncu_bench.log 2>&1 )
( python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 ) > gpurun_out/f_smoke.log 2>&1
tail -n 30 gpurun_out/f_tests.log; cut -c1-300 gpurun_out/f_bench_*.json; tail -3 gpurun_out/f_smoke.log; wc -l gpurun_out/f_launches.csv
