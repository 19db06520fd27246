#!/bin/bash
# round-2 run J (1 GPU): the whole GPU suite + smoke + the default bench line on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BICG_PEER_TIMEOUT_S=6
( timeout 1100 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | cut -c1-400 | tail -25 ) > gpurun_out/j_tests.log 2>&1
( timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -8 ) > gpurun_out/j_smoke.log 2>&1
( timeout 400 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -3 gpurun_out/j_bench.err | cut -c1-300 ) 2>&1
tail -6 gpurun_out/j_tests.log; cat gpurun_out/j_smoke.log; cut -c1-700 gpurun_out/j_bench.json
