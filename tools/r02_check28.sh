#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -k "big or prot2dna or composite" ) > gpurun_out/pytest_gpu28.log 2>&1
tail -4 gpurun_out/pytest_gpu28.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cfg3-pairs 1000 --cfg5-reads 1024 --em-pairs 256 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for c in d['configs']:
    if c['config'] == 'cfg4':
        for leg in ('forward', 'viterbi_score', 'viterbi'): print('cfg4', leg, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in c[leg].items() if k != 'roofline'})
"
