"""Kernel time of the lane engine (batches without input sequences: profile HMMs) under its options: the windowed sweep
at one and two reads per lane, lookahead and block size of its ring, warps per SM, against the first version of the sweep
(state vectors in global memory).  One JSON line per variant; every variant must return the same numbers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from machineboss_b200 import capi

preset = os.environ.get("MACHINE", "PF00516")
n = int(os.environ.get("READS", "65536"))
mj = bench.eval_machine(preset)
lens = 50 + (np.arange(n, dtype=np.int64) * 7919) % 451
if os.environ.get("LEN"):
    lens[:] = int(os.environ["LEN"])
y, y_off = bench.synth_ragged(bench.SEED + 5, lens, mj["n_out"])
batch = capi.Batch(x=np.zeros(0, np.uint8), x_off=np.zeros(n + 1, np.int64), y=y, y_off=y_off)
cells = batch.cell_states(mj["n_states"])
variants = [dict(), dict(lane_r=1), dict(lane_r=2), dict(lane_r=4), dict(lane_r=4, lane_warps_per_cta=2), dict(lane_r=2, lane_bs=16), dict(lane_old=1)]
if os.environ.get("VARIANTS"):
    variants = json.loads(os.environ["VARIANTS"])
ref = None
for opts in variants:
    for k, v in opts.items():
        capi.set_option(k, v)
    capi.set_option("verbose", 1)
    m = bench.make_machine(capi, mj)
    for k in list(opts) + ["verbose"]:
        capi.set_option(k, None)
    out = {"machine": preset, "reads": n, "options": opts}
    ll = capi.forward(m, batch); ll = capi.forward(m, batch)
    ms = batch.last_kernel_ms()[0]
    out["forward_ms"] = round(ms, 2); out["forward_gcups"] = round(cells / ms / 1e6, 1); out["redo"] = batch.last_redo()
    sc = capi.viterbi(m, batch, paths=False); sc = capi.viterbi(m, batch, paths=False)
    ms = batch.last_kernel_ms()[0]
    out["viterbi_score_ms"] = round(ms, 2); out["viterbi_score_gcups"] = round(cells / ms / 1e6, 1)
    if ref is None:
        ref = (ll.copy(), sc.copy())
    out["same_results"] = bool(np.allclose(ll, ref[0], rtol=1e-10) and np.array_equal(sc, ref[1]))
    out["ll0"] = float(ll[0]); out["v0"] = float(sc[0])
    print(json.dumps(out), flush=True)
    m.close()
