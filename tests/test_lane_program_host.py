"""CPU test of the lane engine's windowed program (mb_lane.cu, "lane2"): the program the host builds for a machine
without input alphabet -- shared-memory window of the current cell, ring of previous-cell values, hub sources and
push-form hub destinations -- is EXECUTED ON THE HOST for one read, slot for slot as the kernel does it (every slot
tagged, so a stale window / ring / hub read is an error), and compared with the oracle: Forward as the exact sum,
Viterbi bit for bit, and the back-pointers walked back into the reference's path.  No device involved."""
import math

import numpy as np
import pytest

from helpers import LSE_EXACT, FlatMachine, Oracle, load_golden, synth_tokens

T_INSERT, T_SILENT = 2, 3


def _walk(fm, y, bp, bp_bytes):
    """DPMatrix::traceBack over the engine's (kind, index-in-list) back-pointers [o][state] (mb_wide.cu wide_traceback_kernel)."""
    kb = 6 if bp_bytes == 1 else 14
    S = fm.n_states
    order = np.lexsort((np.arange(fm.n_trans), fm.src))      # incoming lists: source state ascending, then transition index
    o, s, path = len(y), S - 1, []
    while o > 0 or s != 0:
        v = int(bp[o * S + s])
        assert v != 0xffff, "no pointer stored"
        kind, idx = v >> kb, v & ((1 << kb) - 1)
        c = int(y[o - 1]) if (kind == T_INSERT) else 0
        cand = [t for t in order if fm.dst[t] == s and fm.tin[t] == 0 and fm.tout[t] == c and not (c == 0 and fm.dst[t] <= fm.src[t])]
        t = cand[idx]
        path.append(int(t))
        if kind == T_INSERT:
            o -= 1
        s = int(fm.src[t])
    return path[::-1]


@pytest.mark.parametrize("name,lens", [("hmmer_pf00516", [0, 1, 7, 60]), ("unitindel", [0, 3, 9]), ("counter_xxx", [3]), ("hmmer_pf00516_protpsw", [12])])
def test_windowed_program_on_the_host(name, lens):
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    if fm.n_in != 0:      # the lane engine takes batches without input sequences: strip the input side off a transducer
        keep = fm.tin == 0
        fm = FlatMachine(fm.n_states, 0, fm.n_out, fm.src[keep], fm.dst[keep], fm.tin[keep], fm.tout[keep], fm.lw[keep], [], fm.out_alphabet)
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    orc = Oracle(fm)
    x = np.zeros(0, np.uint8)
    for k, lo in enumerate(lens):
        y = synth_tokens(29, k, 1, lo, fm.n_out)
        f, info = capi.lane_emulate(*args, y, 0)
        assert info[0] == 1, "no windowed program for %s" % name
        want = orc.forward(x, y, mode=LSE_EXACT)
        if math.isinf(want):
            assert f == want
        else:
            assert abs(f - want) <= 1e-10 * max(1.0, abs(want)), (name, lo, f, want)
        lse, _ = capi.lane_emulate(*args, y, 2)
        assert (lse == want) if math.isinf(want) else abs(lse - want) <= 1e-10 * max(1.0, abs(want)), (name, lo, lse, want)
        v, info, bp = capi.lane_emulate(*args, y, 1, back_pointers=True)
        want_v, want_p = orc.viterbi(x, y)
        assert v == want_v, (name, lo, v, want_v)
        if math.isfinite(want_v) and fm.n_states <= 3000:
            assert _walk(fm, y, bp, int(info[7])) == want_p.tolist(), (name, lo)
    if name.startswith("hmmer"):
        assert info[3] >= 1 and info[4] >= 1 and info[5] < fm.n_states      # a profile has hub states, and fewer live states than states


def test_window_and_ring_sizes_follow_the_machine():
    """PF00516: every silent edge but those of the begin / end hubs stays within 32 states; the ring is the smallest that the
    schedule check accepts."""
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden("hmmer_pf00516")["machine"])
    _, info = capi.lane_emulate(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw, np.zeros(0, np.uint8), 0)
    assert info[0] == 1 and info[1] == 32 and info[2] <= 64 and info[3] + info[4] <= 4 and info[5] == 975
