#!/bin/bash
# run under gpurun on ONE GPU at the end of a round: the -m gpu suite and smoke(), then tools/profile_round.sh
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -3
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
mj = bench.dnapsw_machine()
m = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
x, xo, y, yo = bench.synth_batch(bench.SEED, 0, 10000, 1000, 1000, 4)
b = capi.Batch(x=x, x_off=xo, y=y, y_off=yo)
for _ in range(3):
    capi.viterbi(m, b, paths=False)
ms, n = b.last_kernel_ms()
print("viterbi score only: %.2f ms, %.0f GCUPS" % (ms, b.cell_states(8) / ms / 1e6))
PY
bash tools/profile_round.sh
