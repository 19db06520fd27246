#!/bin/bash
# round-2 run A (1 GPU): kernel-level + parity tests, loop timings with trace, L2-hint A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/a_kernels.log 2>&1
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -40 ) > gpurun_out/a_parity.log 2>&1
( BICG_MEGA_TRACE=1 timeout 300 python tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab pipe_bicgstab_rr 2>&1 | tail -30 ) > gpurun_out/a_perf_hint1.log 2>&1
( BICG_L2_HINT=0 BICG_MEGA_TRACE=1 QP_MODES=mega timeout 300 python tools/quick_perf.py bicgstab ca_bicgstab pipe_bicgstab 2>&1 | tail -30 ) > gpurun_out/a_perf_hint0.log 2>&1
( QP_KIND=laplace5 QP_G=2000 QP_P0=0 timeout 300 python tools/quick_perf.py pipe_bicgstab bicgstab 2>&1 | tail ) > gpurun_out/a_perf_laplace.log 2>&1
( QP_KIND=random QP_G=2000000 QP_P0=32 QP_ITERS=100 timeout 300 python tools/quick_perf.py ca_bicgstab 2>&1 | tail ) > gpurun_out/a_perf_random.log 2>&1
( timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 ) > gpurun_out/a_bench.log 2>&1
tail -n 40 gpurun_out/a_kernels.log gpurun_out/a_parity.log gpurun_out/a_perf_hint1.log gpurun_out/a_perf_hint0.log gpurun_out/a_perf_laplace.log gpurun_out/a_perf_random.log gpurun_out/a_bench.log
