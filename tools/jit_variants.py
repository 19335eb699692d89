"""Kernel time of the JIT engine's score kernels under the tuning options (normalised sums on / off, Viterbi compares
on the FP64 or the integer pipe, unroll of the steady loop, CTAs per SM): one line of JSON per variant, and a check
that every variant returns the same numbers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from machineboss_b200 import capi

P = int(os.environ.get("P", "10000")); L = int(os.environ.get("L", "1000"))
preset = os.environ.get("MACHINE", "dnapsw")
if preset == "dnapsw":
    mj = bench.dnapsw_machine(); nsym = 4
else:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import FlatMachine, load_golden
    fm = FlatMachine.from_json(load_golden("protpsw_synth")["machine"])
    mj = dict(n_states=fm.n_states, n_in=fm.n_in, n_out=fm.n_out, src=fm.src, dst=fm.dst, tin=fm.tin, tout=fm.tout, lw=fm.lw); nsym = 20
x, x_off, y, y_off = bench.synth_batch(bench.SEED, 0, P, L, L, nsym)
cells = float(L + 1) * float(L + 1) * mj["n_states"] * P
batch = capi.Batch(x=x, x_off=x_off, y=y, y_off=y_off)
variants = [dict(), dict(jit_no_norm=1), dict(jit_vit_intcmp=0), dict(jit_vit_intcmp=2), dict(jit_unroll=2), dict(jit_unroll=2, jit_vit_intcmp=2),
            dict(jit_minblocks_v=2), dict(jit_minblocks_v=2, jit_unroll=2), dict(jit_minblocks_v=4)]
if os.environ.get("VARIANTS"):
    variants = json.loads(os.environ["VARIANTS"])
ref = None
for opts in variants:
    for k, v in opts.items():
        capi.set_option(k, v)
    m = capi.Machine(mj["n_states"], mj["n_in"], mj["n_out"], mj["src"], mj["dst"], mj["tin"], mj["tout"], mj["lw"])
    for k in opts:
        capi.set_option(k, None)
    t = {"forward": [], "viterbi": [], "viterbi_score": []}
    for rep in range(4):
        ll = capi.forward(m, batch); t["forward"].append(batch.last_kernel_ms()[0])
        sc, plen = capi.viterbi_lengths(m, batch); t["viterbi"].append(batch.last_kernel_ms()[0])
        sc2 = capi.viterbi(m, batch, paths=False); t["viterbi_score"].append(batch.last_kernel_ms()[0])
    redo = batch.last_redo()
    if ref is None:
        ref = (ll.copy(), sc.copy(), plen.copy())
    same = bool(np.allclose(ll, ref[0], rtol=1e-12) and np.array_equal(sc, ref[1]) and np.array_equal(sc2, ref[1]) and np.array_equal(plen, ref[2]))
    out = {"options": opts, "same_results": same, "ll0": float(ll[0]), "v0": float(sc[0])}
    for k, v in t.items():
        ms = min(v[1:])
        out[k + "_ms"] = round(ms, 3); out[k + "_gcups"] = round(cells / ms / 1e6, 1)
    print(json.dumps(out), flush=True)
    m.close()
