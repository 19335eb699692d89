"""Throughput of the wide engine (and the generic engine beside it) on config-4 / config-5 style workloads.

  python tools/bench_wide.py --machine prot2dna_dnapsw --pairs 148 --li 300 --lo 10000
  python tools/bench_wide.py --machine hmmer_pf00516 --pairs 4096 --li 0 --lo 275
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from helpers import FlatMachine, load_golden, synth_tokens  # noqa: E402
from machineboss_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--machine", default="prot2dna_dnapsw")
    ap.add_argument("--pairs", type=int, default=148)
    ap.add_argument("--li", type=int, default=300)
    ap.add_argument("--lo", type=int, default=10000)
    ap.add_argument("--engines", default="2")
    ap.add_argument("--generic-pairs", type=int, default=4)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-trace", action="store_true")
    a = ap.parse_args()
    fm = FlatMachine.from_json(load_golden(a.machine)["machine"])
    out = {"machine": a.machine, "S": fm.n_states, "T": fm.n_trans, "li": a.li, "lo": a.lo}
    for eng in [int(e) for e in a.engines.split(",")]:
        n = a.pairs if eng != 0 else min(a.pairs, a.generic_pairs)
        pairs = [(synth_tokens(12345, k, 0, a.li if fm.n_in else 0, max(fm.n_in, 1)), synth_tokens(12345, k, 1, a.lo if fm.n_out else 0, max(fm.n_out, 1))) for k in range(n)]
        capi.set_engine(eng)
        m = capi.Machine(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
        capi.set_engine(-1)
        b = capi.Batch(pairs)
        cs = b.cell_states(fm.n_states)
        r = {"pairs": n, "cell_states": cs}
        for what in ["forward", "viterbi_score"] + ([] if a.no_trace else ["viterbi_trace"]):
            best = None
            for _ in range(a.reps):
                t0 = time.time()
                if what == "forward":
                    v = capi.forward(m, b)
                elif what == "viterbi_score":
                    v = capi.viterbi(m, b, paths=False)
                else:
                    v, plen = capi.viterbi_lengths(m, b)
                wall = time.time() - t0
                ms, nl = b.last_kernel_ms()
                if best is None or ms < best[0]:
                    best = (ms, wall, nl)
            r[what] = {"kernel_ms": best[0], "wall_ms": best[1] * 1e3, "launches": best[2], "gcups": cs / best[0] / 1e6, "v0": float(v[0]), "redo": b.last_redo()}
        out["engine%d" % eng] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
