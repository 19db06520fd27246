#!/bin/bash
# round-2 run N (1 GPU): racecheck with the full hazard list, synccheck and initcheck over smoke()
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=20 BICG_AUTOTUNE=0
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 400 $CS --tool racecheck --racecheck-report analysis --print-limit 400 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | cut -c1-260 ) > gpurun_out/n_racecheck_full.log 2>&1
( timeout 300 $CS --tool synccheck --print-limit 40 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | cut -c1-260 | tail -30 ) > gpurun_out/n_synccheck.log 2>&1
( timeout 300 $CS --tool initcheck --print-limit 40 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | cut -c1-260 | tail -60 ) > gpurun_out/n_initcheck.log 2>&1
grep -c "Race reported" gpurun_out/n_racecheck_full.log; tail -3 gpurun_out/n_racecheck_full.log; tail -5 gpurun_out/n_synccheck.log; tail -12 gpurun_out/n_initcheck.log
