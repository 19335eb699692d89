"""Throughput of the other BASELINE configs' machines (parity is covered by tests/): protpsw (config 3)
and the composed prot2dna => dnapsw machine (config 4 style) on synthetic batches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import FlatMachine, load_golden, synth_tokens
from machineboss_b200 import capi


def batch(n, li, lo, nin, nout, seed=7):
    return [(synth_tokens(seed, k, 0, li, nin), synth_tokens(seed, k, 1, lo, nout)) for k in range(n)]


def run(name, golden, n, li, lo, do_counts=True, reps=2):
    fm = FlatMachine.from_json(load_golden(golden)["machine"])
    m = capi.Machine(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    b = capi.Batch(batch(n, li, lo, max(fm.n_in, 1), max(fm.n_out, 1)))
    cells = b.cell_states(fm.n_states)
    out = {"config": name, "engine": m.engine, "pairs": n, "li": li, "lo": lo, "states": fm.n_states, "trans": fm.n_trans}
    for what, fn in (("forward", lambda: capi.forward(m, b)), ("viterbi+traceback", lambda: capi.viterbi_lengths(m, b)),
                     ("counts", (lambda: capi.counts(m, b)) if do_counts else None)):
        if fn is None:
            continue
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        wall = (time.perf_counter() - t0) / reps
        ms, _ = b.last_kernel_ms()
        out[what] = {"gcups": cells / ms / 1e6, "pairs_per_s": n / wall, "kernel_ms": ms}
    print(json.dumps(out))


if __name__ == "__main__":
    run("protpsw 300x300 (config 3)", "protpsw_synth", int(os.environ.get("N3", "20000")), 300, 300)
    run("prot2dna=>dnapsw 50aa x 600nt (config 4 style, generic engine)", "prot2dna_dnapsw", 16, 50, 600, do_counts=False, reps=1)
