#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), d['clocks'])
"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_golden and (dnapsw_small or hmmer_pf00516_reads or prot2dna_dnapsw-2)" 2>&1 | tail -2
