"""CPU test of the column engine (mb_col.cu): the period analysis of a generator without input, the per-column weight
tables and the prefix / suffix programs are EXECUTED ON THE HOST for one read (column = period of the machine, row =
read position, carried prefix states, accumulators for the suffix) and compared with the oracle -- Forward as the exact
sum, Viterbi bit for bit -- and the generated strip kernel is compiled for sm_100a.  No device involved."""
import math

import numpy as np
import pytest

from helpers import LSE_EXACT, FlatMachine, Oracle, load_golden, synth_tokens, synthetic_profile


def _generator(name):
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    if fm.n_in != 0:
        keep = fm.tin == 0
        fm = FlatMachine(fm.n_states, 0, fm.n_out, fm.src[keep], fm.dst[keep], fm.tin[keep], fm.tout[keep], fm.lw[keep], [], fm.out_alphabet)
    return fm


@pytest.mark.parametrize("name,lens", [("hmmer_pf00516", [0, 1, 7, 60]), ("hmmer_pf00516_protpsw", [0, 12])])
def test_column_program_on_the_host(name, lens):
    from machineboss_b200 import capi
    fm = _generator(name)
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    orc = Oracle(fm)
    x = np.zeros(0, np.uint8)
    for k, lo in enumerate(lens):
        y = synth_tokens(31, k, 1, lo, fm.n_out)
        f, info = capi.col_emulate(*args, y, 0)
        assert info[0] == 1, "no column program for %s" % name
        want = orc.forward(x, y, mode=LSE_EXACT)
        assert (f == want) if math.isinf(want) else abs(f - want) <= 1e-10 * max(1.0, abs(want)), (name, lo, f, want)
        v, _, walked = capi.col_emulate(*args, y, 1, path=True)
        want_v, want_p = orc.viterbi(x, y)
        assert v == want_v, (name, lo, v, want_v)
        assert walked.tolist() == want_p.tolist(), (name, lo)      # the walk back over the program's pointers is the reference's path
    # the structure found: 5 states per profile node (Mx, M, D, Ix, I), times the error model's states in the composition
    assert info[1] == (5 if name == "hmmer_pf00516" else 25) and info[3] >= 480 and info[6] >= 1 and info[7] >= 1
    assert info[1] * info[3] + info[4] + info[5] == fm.n_states


def test_column_kernel_compiles():
    from machineboss_b200 import capi
    fm = _generator("hmmer_pf00516")
    _, info, log = capi.col_emulate(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw, np.zeros(0, np.uint8), 0, compile_log=True)
    assert info[0] == 1 and "mb_k_col_sum" in log and "mb_k_col_max" in log and "mb_k_col_maxp" in log
    spills = [l for l in log.splitlines() if "spill stores" in l]
    assert len(spills) == 3 and all(" 0 bytes spill stores" in l for l in spills), log


@pytest.mark.parametrize("name", ["unitindel", "counter_xxx", "dnapsw_small"])
def test_machines_without_a_period_are_declined(name):
    from machineboss_b200 import capi
    fm = _generator(name)
    _, info = capi.col_emulate(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw, np.zeros(0, np.uint8), 0)
    assert info[0] == 0


def test_synthetic_profile_with_diagonal_groups_and_flanking_states():
    """A hand-made periodic generator the HMMER import never produces: transitions that consume a token while moving to the next
    node (diagonal groups), token-consuming self-loops, a begin hub entered with a token, and N / C flanking states with
    self-loops (prefix and suffix programs with their own dynamics).  Forward, Viterbi and the walk back, against the oracle."""
    from machineboss_b200 import capi
    fm = synthetic_profile()
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    orc = Oracle(fm)
    x = np.zeros(0, np.uint8)
    for k, lo in enumerate([0, 1, 2, 5, 17, 40, 95]):
        y = synth_tokens(37, k, 1, lo, fm.n_out)
        f, info = capi.col_emulate(*args, y, 0)
        assert info[0] == 1 and info[1] == 3, info
        want = orc.forward(x, y, mode=LSE_EXACT)
        assert (f == want) if math.isinf(want) else abs(f - want) <= 1e-10 * max(1.0, abs(want)), (lo, f, want)
        v, _, walked = capi.col_emulate(*args, y, 1, path=True)
        want_v, want_p = orc.viterbi(x, y)
        assert v == want_v, (lo, v, want_v)
        assert walked.tolist() == want_p.tolist(), lo
    _, info, log = capi.col_emulate(*args, np.zeros(0, np.uint8), 0, compile_log=True)
    assert "mb_k_col_maxp" in log


def test_transitions_that_skip_a_node_are_declined():
    from machineboss_b200 import capi
    fm = synthetic_profile(skip=True)
    _, info = capi.col_emulate(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw, np.zeros(0, np.uint8), 0)
    assert info[0] == 0


@pytest.mark.parametrize("n_nodes,n_out,seed", [(70, 2, 1), (101, 20, 2), (66, 3, 3), (200, 4, 4)])
def test_synthetic_profiles_of_other_shapes(n_nodes, n_out, seed):
    """The same hand-made generator at other sizes, alphabets and weights: the analysis must find the period of 3 whatever the
    number of nodes (also when it is not a multiple of the columns per lane), and the program must reproduce the oracle."""
    from machineboss_b200 import capi
    fm = synthetic_profile(n_nodes=n_nodes, n_out=n_out, seed=seed)
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    orc = Oracle(fm)
    x = np.zeros(0, np.uint8)
    for k, lo in enumerate([1, 9, 33]):
        y = synth_tokens(43 + seed, k, 1, lo, fm.n_out)
        f, info = capi.col_emulate(*args, y, 0)
        assert info[0] == 1 and info[1] == 3 and info[1] * info[3] + info[4] + info[5] == fm.n_states, info
        want = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(f - want) <= 1e-10 * max(1.0, abs(want)), (lo, f, want)
        v, _, walked = capi.col_emulate(*args, y, 1, path=True)
        want_v, want_p = orc.viterbi(x, y)
        assert v == want_v and walked.tolist() == want_p.tolist(), lo
