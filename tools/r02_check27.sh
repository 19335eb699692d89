#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for kb in 192 227; do
python - $kb <<'PY' 2>&1 | tail -4
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.eval_machine("prot2dna_dnapsw")
capi.set_option("big_smem_kb", int(sys.argv[1])); capi.set_option("verbose", 1)
m = bench.make_machine(capi, mj)
capi.set_option("verbose", None)
n, lo = 1000, 4000
x, x_off, _, _ = bench.synth_batch(bench.SEED + 4, 0, n, 300, 300, mj["n_in"])
_, _, y, y_off = bench.synth_batch(bench.SEED + 4, 0, n, lo, lo, mj["n_out"])
b = capi.Batch(x=x, x_off=x_off, y=y, y_off=y_off)
cells = b.cell_states(mj["n_states"])
for rep in range(2): ll = capi.forward(m, b)
f = b.last_kernel_ms()[0]
for rep in range(2): sc = capi.viterbi(m, b, paths=False)
v = b.last_kernel_ms()[0]
for rep in range(2): sc2, plen = capi.viterbi_lengths(m, b)
vt = b.last_kernel_ms()[0]
print(sys.argv[1], "KB: forward %.1f ms (%.0f GCUPS), viterbi score %.1f ms (%.0f), with traceback %.1f ms (%.0f)" % (f, cells / f / 1e6, v, cells / v / 1e6, vt, cells / vt / 1e6), flush=True)
PY
done
