#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "big_engine or prot2dna or translate" 2>&1 | tail -4
python - <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.dnapsw_machine()
m = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
for L, P in ((1000, 2000), (4000, 400), (10000, 120)):
    x, xo, y, yo = bench.synth_batch(bench.SEED, 0, P, L, L, 4)
    b = capi.Batch(x=x, x_off=xo, y=y, y_off=yo)
    capi.forward(m, b); ll = capi.forward(m, b)
    ms, n = b.last_kernel_ms()
    print("dnapsw %d pairs of %d: forward %.2f ms, %.0f GCUPS, redo %d of %d, ll0 %.6f" % (P, L, ms, b.cell_states(8) / ms / 1e6, b.last_redo(), P, ll[0]))
    b.close()
PY
for w in 5 4; do MB_BIG_WARPS=$w timeout 300 python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 1000 --li 300 --lo 10000 --engines 2 --no-trace --reps 2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())['engine2']
print('cfg4 1000 x 300 x 10000, warps $w:', 'forward', round(j['forward']['gcups'],1), 'ms', round(j['forward']['kernel_ms'],1), 'redo', j['forward']['redo'], '| viterbi score', round(j['viterbi_score']['gcups'],1))"; done
