#!/bin/bash
# round-2 run L (1 GPU): ncu --set full of the cfg-5 stand-alone SpMV (random 2 M x 32) and of the shifted solver's kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spmv_ -c 1 -o gpurun_out/l_random_spmv python tools/ncu_targets.py random > gpurun_out/l_ncu_random.log 2>&1 )
( timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sh_vec_shift -c 2 -o gpurun_out/l_shifted_vec python tools/ncu_targets.py shifted > gpurun_out/l_ncu_shifted.log 2>&1 )
tail -n 4 gpurun_out/l_ncu_random.log gpurun_out/l_ncu_shifted.log; ls -la gpurun_out/*.ncu-rep
