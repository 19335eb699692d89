"""CPU tests: the C restatement (oracle/mb_oracle.c) against the reference.

Pinned three ways: (1) the reference's own published goldens (t/expect, test-cpu.mjs) carried in
each fixture's "ref_expect"; (2) the unmodified reference binary's outputs stored in the fixtures
(17 digits); (3) internal identities (Forward ll == Backward ll, counts sum rules).
"""
import math

import numpy as np
import pytest

from helpers import (LSE_EXACT, LSE_TABLE, FlatMachine, Oracle, gnum, golden_names, load_golden,
                     oracle_lib, pairs_from_golden, synth_tokens)

CASES = golden_names()


def _same(a, b):
    return (a == b) or (math.isnan(a) and math.isnan(b))


@pytest.mark.parametrize("name", CASES)
def test_oracle_bit_identical_to_reference_binary(name):
    """Forward, rolling Forward, Backward, Viterbi values and paths: identical bits to refdrv."""
    case = load_golden(name)
    m = FlatMachine.from_json(case["machine"])
    orc = Oracle(m)
    counts = np.zeros(m.n_trans)
    for (x, y), p in zip(pairs_from_golden(case), case["pairs"]):
        orc.set_envelope(p.get("env"))
        if "rolling" in p:
            assert _same(orc.forward(x, y), gnum(p["rolling"]))
        if "forward" in p:
            assert _same(orc.forward(x, y), gnum(p["forward"]))
        if "backward" in p:
            assert _same(orc.backward(x, y), gnum(p["backward"]))
        if "viterbi" in p:
            sc, path = orc.viterbi(x, y)
            assert _same(sc, gnum(p["viterbi"]))
            if "path" in p:
                assert path.tolist() == p["path"]
        if "F" in p:
            _, F = orc.forward(x, y, matrix=True)
            _, B = orc.backward(x, y, matrix=True)
            _, _, V = orc.viterbi(x, y, matrix=True)
            for mine, ref in ((F, p["F"]), (B, p["B"]), (V, p["V"])):
                ref = np.array([gnum(v) for v in ref]).reshape(mine.shape)
                assert np.array_equal(mine, ref)
        if case["counts"] is not None:
            orc.counts(x, y, counts=counts)
    orc.set_envelope(None)
    if case["counts"] is not None:
        ref = np.array([gnum(v) for v in case["counts"]])
        # same per-pair arithmetic; summation over pairs is in list order on both sides
        np.testing.assert_allclose(counts, ref, rtol=1e-13, atol=1e-300)


def _round_sf(v, n):
    if v == 0 or math.isinf(v):
        return v
    return round(v, n - 1 - int(math.floor(math.log10(abs(v)))))


def test_reference_published_goldens():
    """The values the reference's own test-suite pins (Makefile:492-572, test-cpu.mjs:129-197)."""
    c = load_golden("bitnoise_tiny")
    m = FlatMachine.from_json(c["machine"])
    orc = Oracle(m)
    x, y = pairs_from_golden(c)[0]
    _, F = orc.forward(x, y, matrix=True)
    _, B = orc.backward(x, y, matrix=True)
    for mat, key in ((F, "forward_matrix_5dp"), (B, "backward_matrix_5dp")):
        for i, o, v in c["ref_expect"][key]:
            v = gnum(v)
            got = mat[o, i, 0]
            assert (got == v) if math.isinf(v) else (float("%.5g" % got) == v), (key, i, o, got, v)
    fll, bll, cnt = orc.counts(x, y)
    assert [round(v, 6) for v in cnt] == c["ref_expect"]["counts"]
    assert _round_sf(fll, 4) == c["ref_expect"]["loglike_4sf"]

    c = load_golden("bitnoise_counts")       # {"p":2,"q":1}: transitions 0,3 carry p, 1,2 carry q
    m = FlatMachine.from_json(c["machine"])
    _, _, cnt = Oracle(m).counts(*pairs_from_golden(c)[0])
    by_w = {}
    for t in range(m.n_trans):
        by_w.setdefault(round(float(np.exp(m.lw[t])), 6), 0.0)
        by_w[round(float(np.exp(m.lw[t])), 6)] += cnt[t]
    assert round(by_w[0.99], 6) == 2 and round(by_w[0.01], 6) == 1

    c = load_golden("counter_xxx")
    _, _, cnt = Oracle(FlatMachine.from_json(c["machine"])).counts(*pairs_from_golden(c)[0])
    assert round(float(cnt.sum()), 6) == 3

    c = load_golden("stutter_noise_difflen")
    m = FlatMachine.from_json(c["machine"])
    sc, path = Oracle(m).viterbi(*pairs_from_golden(c)[0])
    assert [int(m.dst[t]) for t in path] == c["ref_expect"]["path_to"]
    ia, oa = [""] + m.in_alphabet, [""] + m.out_alphabet
    assert [ia[m.tin[t]] for t in path] == c["ref_expect"]["path_in"]
    assert [oa[m.tout[t]] for t in path] == c["ref_expect"]["path_out"]

    c = load_golden("bitstutternoise_0011")
    orc = Oracle(FlatMachine.from_json(c["machine"]))
    x, y = pairs_from_golden(c)[0]
    assert _round_sf(orc.forward(x, y), 3) == c["ref_expect"]["forward_3sf"]
    assert _round_sf(orc.viterbi(x, y, path=False), 3) == c["ref_expect"]["viterbi_3sf"]

    for name in ("bitnoise_p09", "bitnoise_p001"):
        c = load_golden(name)
        orc = Oracle(FlatMachine.from_json(c["machine"]))
        x, y = pairs_from_golden(c)[0]
        assert abs(orc.forward(x, y) - c["ref_expect"]["forward_1e-4"]) < 1e-4
        if "viterbi_1e-4" in c["ref_expect"]:
            assert abs(orc.viterbi(x, y, path=False) - c["ref_expect"]["viterbi_1e-4"]) < 1e-4

    c = load_golden("bitecho")
    orc = Oracle(FlatMachine.from_json(c["machine"]))
    got = [orc.forward(x, y) for x, y in pairs_from_golden(c)]
    assert got == [gnum(v) for v in c["ref_expect"]["forward_exact"]]

    c = load_golden("unitindel")
    orc = Oracle(FlatMachine.from_json(c["machine"]))
    x, y = pairs_from_golden(c)[0]
    assert abs(orc.forward(x, y) - c["ref_expect"]["forward_1e-3"]) < 1e-3
    assert abs(orc.viterbi(x, y, path=False) - c["ref_expect"]["viterbi_1e-3"]) < 1e-3


def test_log_sum_exp_table_properties():
    """logsumexp.h:48-90: cutoff at 10, a==b shortcut, -inf identity, table vs exact."""
    lib = oracle_lib()
    ninf = float("-inf")
    assert lib.mbo_log_sum_exp(ninf, ninf, LSE_TABLE) == ninf
    assert lib.mbo_log_sum_exp(-3.0, ninf, LSE_TABLE) == -3.0
    assert lib.mbo_log_sum_exp(ninf, -3.0, LSE_TABLE) == -3.0
    assert lib.mbo_log_sum_exp(0.0, -10.0, LSE_TABLE) == 0.0          # truncated: true value 4.54e-5
    assert abs(lib.mbo_log_sum_exp(1.0, 1.0, LSE_TABLE) - (1.0 + math.log(2))) < 1e-15
    rng = np.random.default_rng(0)
    worst = 0.0
    for a, b in rng.uniform(-12, 0, size=(2000, 2)):
        worst = max(worst, abs(lib.mbo_log_sum_exp(a, b, LSE_TABLE) - lib.mbo_log_sum_exp(a, b, LSE_EXACT)))
    assert worst < 4.6e-5      # truncation 4.54e-5 + interpolation 3e-10


def test_forward_equals_backward_and_count_identities():
    c = load_golden("dnapsw_synth64")
    m = FlatMachine.from_json(c["machine"])
    orc = Oracle(m)
    for x, y in pairs_from_golden(c):
        f, b, cnt = orc.counts(x, y, mode=LSE_EXACT)
        assert abs(f - b) < 1e-9 * abs(f)
        # every path emits each output symbol exactly once and consumes each input symbol once
        emits = cnt[(m.tout != 0)].sum()
        eats = cnt[(m.tin != 0)].sum()
        assert abs(emits - len(y)) < 1e-6 and abs(eats - len(x)) < 1e-6
        # flow conservation at the end state: exactly one path enters it
        assert abs(cnt[m.dst == m.n_states - 1].sum() - cnt[m.src == m.n_states - 1].sum() - 1) < 1e-6


def test_numpy_synth_matches_c():
    lib = oracle_lib()
    for seed, k, which, n, nsym in ((1, 0, 0, 17, 4), (12345, 9999, 1, 1000, 20), (7, 3, 1, 5, 1)):
        buf = np.zeros(n, dtype=np.uint8)
        lib.mbo_synth(seed, k, which, n, nsym, buf.ctypes.data)
        assert np.array_equal(buf, synth_tokens(seed, k, which, n, nsym))
        assert buf.min() >= 1 and buf.max() <= nsym
