#!/bin/bash
set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu12.log 2>&1
tail -6 gpurun_out/pytest_gpu12.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python -c "
import json; d=json.loads(open('gpurun_out/bench12.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['em']['pairs_per_s'], d['em']['kernel_ms'])"
MB_WIDE_VERBOSE=1 timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 10000 --engines 2 --no-trace --reps 1 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print('wide cfg4', round(j['forward']['gcups'],1), round(j['viterbi_score']['gcups'],1))"
for cfg in "2 32" "4 32" "2 48" "1 48"; do set -- $cfg; echo "lane 262144 reads R=$1 warps=$2"; MB_LANE_R=$1 MB_LANE_WARPS=$2 timeout 300 python tools/bench_wide.py --machine hmmer_pf00516 --pairs 262144 --li 0 --lo 275 --engines 2 --reps 1 --no-trace 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read())['engine2']; print(round(j['forward']['gcups'],1), round(j['viterbi_score']['gcups'],1))"; done 2>&1 | tee gpurun_out/lane12.log
