"""Debug: per-pair posterior counts of the JIT engine against the exact-sum oracle on the shapes of
tests/test_gpu_parity.py::test_chunked_traceback_and_counts (peaked dnapsw, unrelated random sequences)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from helpers import FlatMachine, Oracle, load_golden, synth_tokens, LSE_EXACT
from machineboss_b200 import capi

for mname in ("dnapsw_peaked", "dnapsw_synth64"):
    fm = FlatMachine.from_json(load_golden(mname)["machine"])
    shapes = [(300 - 11 * k, 280 + 9 * k) for k in range(12)]
    pairs = [(synth_tokens(93, k, 0, li, 4), synth_tokens(93, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    for opts in ({}, {"jit_no_linear": 1}):
        for k, v in opts.items():
            capi.set_option(k, v)
        capi.set_engine(1)
        m = capi.Machine(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
        capi.set_engine(-1)
        for k in opts:
            capi.set_option(k, None)
        for k, (x, y) in enumerate(pairs):
            b = capi.Batch([(x, y)])
            cnt, ll = capi.counts(m, b)
            redo = b.last_redo()
            f, bl, want = orc.counts(x, y, mode=LSE_EXACT)
            rel = np.abs(cnt - want) / np.maximum(np.abs(want), 1e-9)
            print("%s %s pair %2d %dx%d redo %d ll %.6f (oracle %.6f) max rel count err %.3e sum %.4f want %.4f" %
                  (mname, opts, k, len(x), len(y), redo, ll[0], f, rel.max(), cnt.sum(), want.sum()), flush=True)
