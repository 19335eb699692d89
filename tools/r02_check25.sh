#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for opt in "" "--no-fold"; do
python - $opt <<'PY' 2>&1 | tail -3
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.eval_machine("prot2dna_dnapsw")
if "--no-fold" in sys.argv: capi.set_option("big_no_fold", 1)
m = bench.make_machine(capi, mj)
for n, lo in ((1000, 10000), (592, 2000), (296, 2000)):
    x, x_off, _, _ = bench.synth_batch(bench.SEED + 4, 0, n, 300, 300, mj["n_in"])
    _, _, y, y_off = bench.synth_batch(bench.SEED + 4, 0, n, lo, lo, mj["n_out"])
    b = capi.Batch(x=x, x_off=x_off, y=y, y_off=y_off)
    cells = b.cell_states(mj["n_states"])
    for rep in range(2):
        ll = capi.forward(m, b)
    ms = b.last_kernel_ms()[0]
    print(sys.argv[1:], n, lo, "forward %.1f ms, %.1f GCUPS, ll0 %.6f" % (ms, cells / ms / 1e6, ll[0]), flush=True)
    b.close()
PY
done
