#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "synthetic or column or hmmer" ) > gpurun_out/pytest_gpu17.log 2>&1
tail -12 gpurun_out/pytest_gpu17.log
python - > gpurun_out/col_tb_launches.txt 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.eval_machine("PF00516")
n = 65536
lens = 50 + (np.arange(n, dtype=np.int64) * 7919) % 451
y, y_off = bench.synth_ragged(bench.SEED + 5, lens, mj["n_out"])
b = capi.Batch(x=np.zeros(0, np.uint8), x_off=np.zeros(n + 1, np.int64), y=y, y_off=y_off)
m = bench.make_machine(capi, mj)
import torch
for rep in range(2):
    sc, plen = capi.viterbi_lengths(m, b)
print("viterbi + traceback ms", b.last_kernel_ms())
sc = capi.viterbi(m, b, paths=False); sc = capi.viterbi(m, b, paths=False)
print("viterbi score ms", b.last_kernel_ms())
PY
cat gpurun_out/col_tb_launches.txt | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_coltb.csv python - > /dev/null 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.eval_machine("PF00516")
n = 65536
lens = 50 + (np.arange(n, dtype=np.int64) * 7919) % 451
y, y_off = bench.synth_ragged(bench.SEED + 5, lens, mj["n_out"])
b = capi.Batch(x=np.zeros(0, np.uint8), x_off=np.zeros(n + 1, np.int64), y=y, y_off=y_off)
m = bench.make_machine(capi, mj)
sc, plen = capi.viterbi_lengths(m, b)
PY
grep -v "^==" gpurun_out/launches_coltb.csv | awk -F'","' '{print $5, $(NF)}' | tail -8
