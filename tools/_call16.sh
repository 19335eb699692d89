#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu16.log 2>&1
tail -6 gpurun_out/pytest_gpu16.log
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 2000 --engines 2 --no-trace --reps 2 2> gpurun_out/big16.err | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print('big  cfg4 300x2000 x592', round(j['forward']['gcups'],1), j['forward']['v0'], j['forward']['redo'], round(j['viterbi_score']['gcups'],1))"
tail -3 gpurun_out/big16.err
MB_NO_BIG=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 2000 --engines 2 --no-trace --reps 2 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print('wide cfg4 300x2000 x592', round(j['forward']['gcups'],1), j['forward']['v0'], j['forward']['redo'])"
timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 592 --li 300 --lo 10000 --engines 2 --no-trace --reps 2 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print('big  cfg4 300x10000 x592', round(j['forward']['gcups'],1), j['forward']['v0'], j['forward']['redo'])"
