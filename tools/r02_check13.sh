#!/bin/bash
set -u
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "column or hmmer or lane or strip_widths" ) > gpurun_out/pytest_gpu13.log 2>&1
tail -12 gpurun_out/pytest_gpu13.log
python - <<'PY' 2>&1 | tail -12
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from machineboss_b200 import capi
for preset, n in (("PF00516", 65536), ("PF00516_protpsw", 16384)):
    mj = bench.eval_machine(preset)
    lens = 50 + (np.arange(n, dtype=np.int64) * 7919) % 451
    y, y_off = bench.synth_ragged(bench.SEED + 5, lens, mj["n_out"])
    b = capi.Batch(x=np.zeros(0, np.uint8), x_off=np.zeros(n + 1, np.int64), y=y, y_off=y_off)
    cells = b.cell_states(mj["n_states"])
    for opts in (dict(), dict(col_no_traceback=1)):
        for k, v in opts.items(): capi.set_option(k, v)
        capi.set_option("verbose", 1)
        m = bench.make_machine(capi, mj)
        for k in list(opts) + ["verbose"]: capi.set_option(k, None)
        sc, plen = capi.viterbi_lengths(m, b); sc, plen = capi.viterbi_lengths(m, b)
        ms, nl = b.last_kernel_ms()
        print(preset, n, opts, "viterbi + traceback %.1f ms, %.1f GCUPS, %d launches, path ids %d, score0 %.6f" % (ms, cells / ms / 1e6, nl, int(plen.sum()), sc[0]), flush=True)
        m.close()
PY
