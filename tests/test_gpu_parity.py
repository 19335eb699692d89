"""GPU parity tests (-m gpu): the CUDA engines, called through the C ABI, against
(a) the committed reference outputs in tests/golden and (b) the C oracle on seeded inputs.

Tolerances (BASELINE.json north_star): Forward / Backward log-likelihoods, Viterbi scores and
counts within 1e-4 relative of the reference, whose log-sum-exp table itself carries up to 4.5e-5
absolute error per operation; Viterbi scores are FP64 add/max and must be bit-identical; traceback
paths must be identical under the reference's first-maximum tie-break.
"""
import math

import numpy as np
import pytest

from helpers import (LSE_EXACT, FlatMachine, Oracle, gnum, golden_names, load_golden, pairs_from_golden, synthetic_profile,
                     synth_tokens)

pytestmark = pytest.mark.gpu

REL = 1e-4
ENGINES = [0, 1, 2]      # generic, jit (small machines), wide


def _capi():
    from machineboss_b200 import capi
    return capi


def make_machine(capi, m: FlatMachine, engine: int, **options):
    capi.set_engine(engine)
    for k, v in options.items():
        capi.set_option(k, v)
    try:
        return capi.Machine(m.n_states, m.n_in, m.n_out, m.src, m.dst, m.tin, m.tout, m.lw)
    except capi.MachineBossError as e:
        if engine == 1 and "jit" in str(e).lower():
            pytest.skip("JIT engine does not take this machine: %s" % e)
        if engine == 2 and "wide" in str(e).lower():
            pytest.skip("wide engine does not take this machine: %s" % e)
        raise
    finally:
        capi.set_engine(-1)
        for k in options:
            capi.set_option(k, None)


def close(a, b, rel=REL):
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= rel * max(1.0, abs(b))


def forward_agrees(fm, x, y, got, want):
    """got (device) against want (the reference's table-based value).  Within 1e-4 relative, or, on machines
    with a large fan-in: the reference's log-sum-exp table returns 0 for x >= 10 (logsumexp.h:52), so a sum
    over n terms can come out low by up to log1p(n e^-10) per cell (0.1 for the 2432 sources of the composed
    profile's end state), which exceeds 1e-4 relative on short reads.  The device sums are exact: they must
    then agree with the exact-sum oracle to 1e-6, and the reference may only be lower, by at most that bound."""
    if close(got, want):
        return True
    exact = Oracle(fm).forward(x, y, mode=LSE_EXACT)
    fan_in = int(np.bincount(fm.dst, minlength=fm.n_states).max())
    bound = (len(x) + len(y) + 1) * math.log1p(fan_in * math.exp(-10.0))
    return abs(got - exact) <= 1e-6 * max(1.0, abs(exact)) and -1e-6 <= exact - want <= bound


def counts_agree(fm, pairs, got, want):
    """Posterior counts of a batch against the reference's (table-based) values `want`.  Within 1e-4 relative -- or,
    where the reference's own approximation is larger than that: its log-sum-exp table drops every term more than
    e^-10 below the running sum (logsumexp.h:52), which on peaked parameters moves its own counts by up to 3e-4
    relative (measured here, oracle table mode against oracle exact mode).  The device sums are exact, so they must
    then match the exact-sum oracle ten times closer than the stated tolerance, 1e-5 (the E-step stores Forward
    values rounded to 21 bits: 2e-6), and the reference must sit within 5e-4 of the exact values."""
    got, want = np.asarray(got), np.asarray(want)
    if np.allclose(got, want, rtol=REL, atol=1e-7):
        return True
    orc = Oracle(fm)
    exact = np.zeros(fm.n_trans)
    for x, y in pairs:
        orc.counts(x, y, mode=LSE_EXACT, counts=exact)
    return np.allclose(got, exact, rtol=1e-5, atol=1e-7) and np.allclose(want, exact, rtol=5e-4, atol=1e-7)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", golden_names())
def test_golden(name, engine):
    capi = _capi()
    case = load_golden(name)
    fm = FlatMachine.from_json(case["machine"])
    pairs = pairs_from_golden(case)
    m = make_machine(capi, fm, engine)
    b = capi.Batch(pairs)
    ref = case["pairs"]
    if any("env" in p for p in ref):      # pairs that carried an alignment: the reference's path envelope
        b.set_envelopes([p.get("env") for p in ref])
    if any("rolling" in p or "forward" in p for p in ref):
        ll = capi.forward(m, b)
        for k, p in enumerate(ref):
            want = gnum(p.get("rolling", p.get("forward")))
            assert forward_agrees(fm, pairs[k][0], pairs[k][1], ll[k], want), (name, k, ll[k], want)
    if any("backward" in p for p in ref):
        bl = capi.backward(m, b)
        for k, p in enumerate(ref):
            assert close(bl[k], gnum(p["backward"])), (name, k, bl[k], p["backward"])
    if any("viterbi" in p for p in ref):
        sc, paths = capi.viterbi(m, b)
        sc2 = capi.viterbi(m, b, paths=False)
        for k, p in enumerate(ref):
            want = gnum(p["viterbi"])
            assert sc[k] == want and sc2[k] == want, (name, k, sc[k], want)     # bit-exact
            if "path" in p:
                assert paths[k].tolist() == p["path"], (name, k)
    if case["counts"] is not None:
        c, ll = capi.counts(m, b)
        want = np.array([gnum(v) for v in case["counts"]])
        finite = [gnum(p["forward"]) for p in ref if not math.isinf(gnum(p["forward"]))]
        if len(finite) == len(ref):      # the reference's counts are NaN-poisoned by impossible pairs
            assert counts_agree(fm, pairs, c, want), (name, np.abs(c - want).max())
            assert close(float(ll.sum()), gnum(case["loglike"]))


@pytest.mark.parametrize("engine", ENGINES)
def test_against_oracle_ragged_batch(engine):
    """A ragged batch (empty, short, long, rectangular pairs) against the oracle, table and exact."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_synth64")["machine"])
    # (input lengths that are a multiple of the E-step's 128-column strip, or just above one, put the origin of the
    # mirrored Backward sweep on the strip's last lane: a pair of 256 x 316 once came back with all counts zero)
    shapes = [(0, 0), (1, 0), (0, 1), (1, 1), (5, 40), (40, 5), (33, 33), (64, 64), (127, 129), (200, 150), (31, 32), (257, 3),
              (128, 100), (256, 316), (129, 77), (131, 40), (255, 64), (127, 33), (384, 20), (130, 0)]
    pairs = [(synth_tokens(77, k, 0, li, 4), synth_tokens(77, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, engine)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    bl = capi.backward(m, b)
    sc, paths = capi.viterbi(m, b)
    cnt, cll = capi.counts(m, b)
    want_cnt = np.zeros(fm.n_trans)
    for k, (x, y) in enumerate(pairs):
        f_tab = orc.forward(x, y)
        f_ex = orc.forward(x, y, mode=LSE_EXACT)
        assert close(ll[k], f_tab) and close(cll[k], f_tab)
        assert abs(ll[k] - f_ex) <= 1e-6 * max(1.0, abs(f_ex))     # the device softplus is exact up to the cutoff
        assert close(bl[k], orc.backward(x, y))
        v, p = orc.viterbi(x, y)
        assert sc[k] == v
        assert paths[k].tolist() == p.tolist()
        orc.counts(x, y, counts=want_cnt)
    np.testing.assert_allclose(cnt, want_cnt, rtol=REL, atol=1e-7)
    # the same paths through the caller-owned-buffer entry points, ids as int32, uint16 and bytes (mb_viterbi_paths_narrow)
    for dt in (np.int32, np.uint16, np.uint8):
        score, plen, off = np.empty(len(pairs)), np.zeros(len(pairs), np.int64), np.zeros(len(pairs) + 1, np.int64)
        trans = np.zeros(sum(len(p) for p in paths) + 3, dtype=dt)
        total = capi.viterbi_into(m, b, score, plen, off, trans)
        assert total == sum(len(p) for p in paths) and np.array_equal(score, sc)
        for k, p in enumerate(paths):
            assert trans[off[k]:off[k + 1]].tolist() == p.tolist(), (dt, k)


@pytest.mark.parametrize("engine", ENGINES)
def test_update_weights(engine):
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    flat = FlatMachine.from_json(load_golden("dnapsw_synth64")["machine"])
    pairs = [(synth_tokens(5, k, 0, 50 + k, 4), synth_tokens(5, k, 1, 60 - k, 4)) for k in range(4)]
    m = make_machine(capi, flat, engine)
    b = capi.Batch(pairs)
    before = capi.forward(m, b)
    m.update_weights(fm.lw)
    after = capi.forward(m, b)
    orc = Oracle(fm)
    for k, (x, y) in enumerate(pairs):
        assert close(after[k], orc.forward(x, y))
    assert not np.allclose(before, after)


def test_wide_engine_multi_strip_composite():
    """prot2dna => dnapsw (308 states, 12 silent levels): pairs wider than one strip of columns, against the oracle."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("prot2dna_dnapsw")["machine"])
    shapes = [(40, 130), (17, 60), (16, 33), (0, 5), (3, 0), (35, 90)]
    pairs = [(synth_tokens(11, k, 0, li, fm.n_in), synth_tokens(11, k, 1, lo, fm.n_out)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 2)
    assert m.engine == 2
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    sc, paths = capi.viterbi(m, b)
    sc2 = capi.viterbi(m, b, paths=False)
    for k, (x, y) in enumerate(pairs):
        assert close(ll[k], orc.forward(x, y)), (k, ll[k], orc.forward(x, y))
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and sc2[k] == v, (k, sc[k], v)
        assert paths[k].tolist() == p.tolist(), k


@pytest.mark.parametrize("name", ["bitnoise_tiny", "unitindel", "dnapsw_small", "stutter_noise_difflen", "dnapsw_path_envelope"])
def test_stored_matrices(name):
    """mb_matrix (DPMatrix::cell, dpmatrix.h:128-146): whole Forward / Backward / Viterbi matrices against the
    reference's, cell by cell, including the -inf cells outside a path envelope."""
    capi = _capi()
    case = load_golden(name)
    fm = FlatMachine.from_json(case["machine"])
    pairs = pairs_from_golden(case)
    m = make_machine(capi, fm, -1)
    b = capi.Batch(pairs)
    if any("env" in p for p in case["pairs"]):
        b.set_envelopes([p.get("env") for p in case["pairs"]])
    for k, p in enumerate(case["pairs"]):
        if "F" not in p:
            continue
        for kind, key in ((0, "F"), (1, "B"), (2, "V")):
            want = np.array([gnum(v) for v in p[key]])
            got = capi.matrix(m, b, k, kind).reshape(-1)
            assert got.shape == want.shape, (name, k, key)
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(got), fin), (name, k, key)
            if kind == 2:
                assert np.array_equal(got[fin], want[fin]), (name, k, key)      # add + max only: bit-exact
            else:
                np.testing.assert_allclose(got[fin], want[fin], rtol=1e-7, atol=1e-7)      # the table interpolates, the device evaluates log1p(exp())


@pytest.mark.parametrize("narrow", [0, 1])
def test_jit_strip_widths(narrow):
    """The score-only kernels exist at 4 and at 8 columns per lane (the wider ones with a frame per lane in the
    linear sweeps); the engine picks per call.  Both, forced, on pairs that span several strips of either width."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    shapes = [(300, 280), (0, 7), (129, 40), (257, 300), (520, 64), (31, 530)]
    pairs = [(synth_tokens(91, k, 0, li, 4), synth_tokens(91, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1, jit_narrow=narrow)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    bl = capi.backward(m, b)
    sc, paths = capi.viterbi(m, b)
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(ll[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, ll[k], f)
        assert abs(bl[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, bl[k], f)
        v, p = orc.viterbi(x, y)
        assert sc[k] == v, (k, sc[k], v)
        assert paths[k].tolist() == p.tolist(), k


@pytest.mark.parametrize("name,fit_c,max_li", [("dnapsw_peaked", 10, 319), ("dnapsw_peaked", 5, 150), ("dnapsw_peaked", 12, 383), ("protpsw_synth", 10, 300), ("protpsw_synth", 3, 95)])
def test_jit_strips_fitted_to_the_batch(name, fit_c, max_li):
    """A score module compiled for the batch (mb_jit.cu choose_width / ensure_fit_module): every pair in ONE strip of 32 * C
    columns, C not a power of two, so the Viterbi pointers of a lane sit in a padded 16-byte group (MB_TBPAD).  Forced with
    jit_fit_c on ragged pairs up to the strip's last column; Forward, Backward, Viterbi scores bit for bit and the paths."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    shapes = [(max_li, 120), (max_li - 1, 33), (0, 7), (1, 0), (max_li // 2, 200), (31, 64), (32, 31), (33, 400)]
    pairs = [(synth_tokens(93, k, 0, li, fm.n_in), synth_tokens(93, k, 1, lo, fm.n_out)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1, jit_fit_c=fit_c, verbose=1)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    bl = capi.backward(m, b)
    sc, paths = capi.viterbi(m, b)
    sc2 = capi.viterbi(m, b, paths=False)
    ref = make_machine(capi, fm, 1, jit_fit_c=0)
    assert np.array_equal(capi.viterbi(ref, b, paths=False), sc)
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(ll[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, ll[k], f)
        assert abs(bl[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, bl[k], f)
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and sc2[k] == v, (k, sc[k], v)
        assert paths[k].tolist() == p.tolist(), k


def test_jit_fitted_strips_are_chosen_for_uniform_short_pairs(capfd):
    """300 x 300 pairs in a batch of 512: the cost model prefers one strip of 320 columns to 3 of 128 or 2 of 256, compiles
    the module once, and the numbers do not change."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    pairs = [(synth_tokens(95, k, 0, 300 - (k % 7), 4), synth_tokens(95, k, 1, 290 + (k % 11), 4)) for k in range(512)]
    b = capi.Batch(pairs)
    m = make_machine(capi, fm, 1, verbose=1)
    ll = capi.forward(m, b)
    sc, paths = capi.viterbi(m, b)
    err = capfd.readouterr().err
    assert "fitted to the batch compiled: 10 columns per lane" in err and err.count("fitted to the batch compiled") == 1
    ref = make_machine(capi, fm, 1, jit_fit_c=0)
    ll0 = capi.forward(ref, b)
    sc0, paths0 = capi.viterbi(ref, b)
    np.testing.assert_allclose(ll, ll0, rtol=1e-12)
    assert np.array_equal(sc, sc0) and all(np.array_equal(p, q) for p, q in zip(paths, paths0))


@pytest.mark.parametrize("reads_per_lane", [1, 2, 4])
def test_lane_engine_reads_per_lane(reads_per_lane):
    """Batches without input sequences go through the lane engine (a read per lane, mb_lane.cu): every
    reads-per-lane variant, ragged read lengths filling more than one task, against the oracle."""
    capi = _capi()
    for name, n_reads, max_len in (("unitindel", 150, 12), ("hmmer_pf00516", 140, 9)):
        fm = FlatMachine.from_json(load_golden(name)["machine"])
        lens = [(7 * k + 3) % (max_len + 1) for k in range(n_reads)]
        pairs = [(np.zeros(0, np.uint8), synth_tokens(21, k, 1, lo, fm.n_out)) for k, lo in enumerate(lens)]
        orc = Oracle(fm)
        m = make_machine(capi, fm, 2, lane_r=reads_per_lane, no_col=1)      # (the column engine would take the profile's sweeps)
        b = capi.Batch(pairs)
        ll = capi.forward(m, b)
        sc, paths = capi.viterbi(m, b)
        sc2 = capi.viterbi(m, b, paths=False)
        for k in list(range(0, n_reads, 13)) + [n_reads - 1]:
            x, y = pairs[k]
            f = orc.forward(x, y)
            assert forward_agrees(fm, x, y, ll[k], f), (name, k, ll[k], f)
            v, p = orc.viterbi(x, y)
            assert sc[k] == v and sc2[k] == v, (name, k, sc[k], v)
            if math.isfinite(v):
                assert paths[k].tolist() == p.tolist(), (name, k)
            else:
                assert len(paths[k]) == 0


def test_linear_sweep_long_pairs_stay_in_the_linear_domain():
    """Pairs long enough for a dozen 256-column strips: far from the diagonal the first rows of a strip have underflowed
    to zero, and an empty lane must take its frame from the first boundary row that holds something; nothing may be
    handed to the log-domain kernel, and the values are the exact sums."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_synth64")["machine"])
    shapes = [(3000, 2800), (2100, 2600)]
    pairs = [(synth_tokens(57, k, 0, li, 4), synth_tokens(57, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    assert b.last_redo() == 0
    bl = capi.backward(m, b)
    assert b.last_redo() == 0
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(ll[k] - f) <= 1e-9 * abs(f), (k, ll[k], f)
        assert abs(bl[k] - f) <= 1e-9 * abs(f), (k, bl[k], f)


def test_big_engine_against_wide_engine():
    """Mid-size machines with full matrices run on the generated thread-per-cell sweep (mb_big.cu); the same pairs
    through the table-driven wide engine (option no_big) must give the same Forward values to rounding, bit-identical
    Viterbi scores and identical paths.  Pairs long enough for several strips whose first rows no path reaches
    (the case that needs an empty lane to adopt its frame from the boundary), none of them flagged."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("prot2dna_dnapsw")["machine"])
    shapes = [(120, 700), (70, 150), (33, 400), (3, 10), (0, 4), (64, 64)]
    pairs = [(synth_tokens(31, k, 0, li, fm.n_in), synth_tokens(31, k, 1, lo, fm.n_out)) for k, (li, lo) in enumerate(shapes)]
    b = capi.Batch(pairs)
    m_big = make_machine(capi, fm, 2)
    ll = capi.forward(m_big, b)
    redo = b.last_redo()
    sc, paths = capi.viterbi(m_big, b)
    sc_only = capi.viterbi(m_big, b, paths=False)
    m_wide = make_machine(capi, fm, 2, no_big=1)
    ll_w = capi.forward(m_wide, b)
    sc_w, paths_w = capi.viterbi(m_wide, b)
    assert redo == 0
    for k in range(len(pairs)):
        assert close(ll[k], ll_w[k], rel=1e-11), (k, ll[k], ll_w[k])
        assert sc[k] == sc_w[k] and sc_only[k] == sc_w[k], (k, sc[k], sc_w[k])
        assert paths[k].tolist() == paths_w[k].tolist(), k
    orc = Oracle(fm)
    x, y = pairs[1]
    assert forward_agrees(fm, x, y, ll[1], orc.forward(x, y))
    v, p = orc.viterbi(x, y)
    assert sc[1] == v and paths[1].tolist() == p.tolist()


def test_big_engine_folded_sums_follow_the_weights():
    """The big engine's Forward keeps one folded value per class of proportional insert groups (mb_big.cu); the classes depend on the
    weights' ratios.  New weights that keep the ratios are uploaded; weights that break them (every transition perturbed on its own)
    make the engine generate itself again -- and with folding switched off (big_no_fold) the same numbers come out."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("prot2dna_dnapsw")["machine"])
    shapes = [(40, 260), (12, 90), (33, 140), (0, 5)]
    pairs = [(synth_tokens(41, k, 0, li, fm.n_in), synth_tokens(41, k, 1, lo, fm.n_out)) for k, (li, lo) in enumerate(shapes)]
    b = capi.Batch(pairs)
    m = make_machine(capi, fm, 2, verbose=1)
    plain = make_machine(capi, fm, 2, big_no_fold=1)
    ll = capi.forward(m, b)
    np.testing.assert_allclose(ll, capi.forward(plain, b), rtol=1e-12)
    orc = Oracle(fm)
    for k, (x, y) in enumerate(pairs):
        assert close(ll[k], orc.forward(x, y, mode=LSE_EXACT), rel=1e-9), k
    rng = np.random.default_rng(5)
    for trial, lw in enumerate([fm.lw + 0.25, np.where(np.isfinite(fm.lw), fm.lw + rng.uniform(-0.3, 0.3, fm.lw.shape), fm.lw)]):
        m.update_weights(lw)      # trial 0: every ratio kept; trial 1: none
        fm2 = fm.with_weights(lw)
        orc2 = Oracle(fm2)
        ll2 = capi.forward(m, b)
        assert b.last_redo() == 0
        for k, (x, y) in enumerate(pairs):
            assert close(ll2[k], orc2.forward(x, y, mode=LSE_EXACT), rel=1e-9), (trial, k)
        sc, paths = capi.viterbi(m, b)
        v, p = orc2.viterbi(*pairs[1])
        assert sc[1] == v and paths[1].tolist() == p.tolist()


def test_wide_engine_log_domain_rerun():
    """The trap machine through the wide engine: the scaled sweep flags the long pairs and the log-domain sweep redoes them."""
    capi = _capi()
    fm = _trap_machine()
    ys = [np.array([1] * n + [2], dtype=np.uint8) for n in (10, 250, 400, 900)]
    pairs = [(np.zeros(0, np.uint8), y) for y in ys]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 2)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    assert b.last_redo() >= 2
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y)
        assert math.isfinite(f) and close(ll[k], f), (k, ll[k], f)
    fm2 = fm.with_weights(np.where(fm.lw < 0, -60.0, fm.lw))      # extreme weights: log domain from the start
    m2 = make_machine(capi, fm2, 2)
    ll2 = capi.forward(m2, b)
    for k, (x, y) in enumerate(pairs):
        assert close(ll2[k], Oracle(fm2).forward(x, y))


def _trap_machine():
    """A 'fast but doomed' branch next to a slow real one: state 1 emits a's with weight 1 and can
    never leave; state 2 emits a or b with weight 1/8 and reaches the end.  After n a's the live
    path is 8^-n below the trap in the same cell: the case the scaled linear sweep must hand over."""
    w8 = math.log(0.125)
    #          src dst in out lw
    trans = [(0, 1, 0, 0, 0.0), (0, 2, 0, 0, 0.0), (1, 1, 0, 1, 0.0), (2, 2, 0, 1, w8), (2, 2, 0, 2, w8), (2, 3, 0, 0, 0.0)]
    a = np.array(trans)
    return FlatMachine(4, 0, 2, a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].astype(np.int32),
                       a[:, 3].astype(np.int32), a[:, 4].astype(np.float64), [], ["a", "b"])


def test_linear_sweep_hands_dangerous_pairs_to_log_domain():
    capi = _capi()
    fm = _trap_machine()
    ys = [np.array([1] * n + [2], dtype=np.uint8) for n in (10, 250, 400, 900)]
    pairs = [(np.zeros(0, np.uint8), y) for y in ys]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    assert b.last_redo() >= 3            # 250, 400 and 900 a's exceed the 2^700 spread (or underflow outright)
    bl = capi.backward(m, b)
    cnt, cll = capi.counts(m, b)
    want = np.zeros(fm.n_trans)
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y)
        assert math.isfinite(f) and close(ll[k], f) and close(cll[k], f) and close(bl[k], orc.backward(x, y))
        orc.counts(x, y, counts=want)
    np.testing.assert_allclose(cnt, want, rtol=REL, atol=1e-7)
    # a machine with extreme weights never uses the linear sweep
    fm2 = fm.with_weights(np.where(fm.lw < 0, -60.0, fm.lw))
    m2 = make_machine(capi, fm2, 1)
    ll2 = capi.forward(m2, b)
    assert b.last_redo() == 0
    for k, (x, y) in enumerate(pairs):
        assert close(ll2[k], Oracle(fm2).forward(x, y))


def test_errors():
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_small")["machine"])
    # a silent transition going backwards: "Machine is not topologically sorted" (eval.cpp:44)
    bad_dst = fm.dst.copy()
    silent = np.where((fm.tin == 0) & (fm.tout == 0) & (fm.src >= 1))[0][0]
    bad_dst[silent] = fm.src[silent]
    with pytest.raises(capi.MachineBossError, match="topologically"):
        capi.Machine(fm.n_states, fm.n_in, fm.n_out, fm.src, bad_dst, fm.tin, fm.tout, fm.lw)
    with pytest.raises(capi.MachineBossError):
        capi.Machine(0, 0, 0, [], [], [], [], [])
    m = capi.Machine(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    b = capi.Batch([])
    assert capi.forward(m, b).shape == (0,)


def test_set_b_mutated_pairs_at_config_size():
    """SURVEY 8(d) parameter set B at config size: peaked dnapsw weights, 1 kb pairs whose output is the input with
    10 % substitutions and 2 % indels.  These are the hard inputs for the scaled linear sweeps (the likelihood
    concentrates on a narrow diagonal band, so the dynamic range across a strip is at its largest): every pass
    against the oracle, and NONE of the pairs may be handed to the log-domain kernels."""
    from helpers import mutate_tokens
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked_1k")["machine"])
    xs = [synth_tokens(401, k, 0, 1000 - 7 * k, 4) for k in range(6)]
    pairs = [(x, mutate_tokens(401, k, x, 4)) for k, x in enumerate(xs)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    assert b.last_redo() == 0
    bl = capi.backward(m, b)
    assert b.last_redo() == 0
    sc, paths = capi.viterbi(m, b)
    cnt, cll = capi.counts(m, b)
    assert b.last_redo() == 0
    want = np.zeros(fm.n_trans)
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(ll[k] - f) <= 1e-9 * abs(f) and abs(bl[k] - f) <= 1e-9 * abs(f), (k, ll[k], bl[k], f)
        assert close(ll[k], orc.forward(x, y)) and close(cll[k], f)
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and paths[k].tolist() == p.tolist(), k
        orc.counts(x, y, counts=want)
    np.testing.assert_allclose(cnt, want, rtol=REL, atol=1e-7)


@pytest.mark.parametrize("reads_per_lane", [1, 2, 4])
def test_lane_engine_config5_read_lengths(reads_per_lane):
    """BASELINE config 5's own read lengths (50 - 500 residues, ragged, several lane tasks, the last one partly
    empty) through the lane engine at every reads-per-lane variant: PF00516 and PF00516 => protpsw."""
    capi = _capi()
    for name, n_reads in (("hmmer_pf00516", 150), ("hmmer_pf00516_protpsw", 70)):
        fm = FlatMachine.from_json(load_golden(name)["machine"])
        lens = [50 + (k * 37) % 451 for k in range(n_reads)]
        lens[0], lens[1], lens[2] = 500, 50, 275
        pairs = [(np.zeros(0, np.uint8), synth_tokens(23, k, 1, lo, fm.n_out)) for k, lo in enumerate(lens)]
        orc = Oracle(fm)
        m = make_machine(capi, fm, 2, lane_r=reads_per_lane, no_col=1)      # (the column engine would take the profile's sweeps)
        b = capi.Batch(pairs)
        ll = capi.forward(m, b)
        assert b.last_redo() == 0, (name, b.last_redo())
        sc, paths = capi.viterbi(m, b)
        sc2 = capi.viterbi(m, b, paths=False)
        check = sorted(set([0, 1, 2, 31, 32, 33, n_reads - 1] + list(range(5, n_reads, 29))))
        for k in check:
            x, y = pairs[k]
            f = orc.forward(x, y)
            assert forward_agrees(fm, x, y, ll[k], f), (name, k, ll[k], f)
            v, p = orc.viterbi(x, y)
            assert sc[k] == v and sc2[k] == v, (name, k, sc[k], v)
            assert paths[k].tolist() == p.tolist(), (name, k)


@pytest.mark.parametrize("opts", [dict(), dict(col_c=1), dict(col_c=3, col_minblocks=1), dict(col_r=2, col_bnd_budget_mb=1, col_bp_budget_mb=4), dict(col_threads=64, col_sil_regs=0)])
def test_column_engine_profile_reads(opts):
    """Periodic generators (PF00516 and PF00516 => protpsw) through the column engine (mb_col.cu: column = profile node,
    row = read position): read lengths around the strip-staging block (0, 1, 15 - 17, 31 - 33) and config 5's own 50 - 500,
    one to three columns per lane, several reads per warp, boundary buffers forced into several chunks -- Forward against the
    oracle, Viterbi scores bit for bit, and the same numbers as the lane engine's own sweep."""
    capi = _capi()
    for name, n_reads in (("hmmer_pf00516", 150), ("hmmer_pf00516_protpsw", 40)):
        fm = FlatMachine.from_json(load_golden(name)["machine"])
        lens = [50 + (k * 37) % 451 for k in range(n_reads)]
        lens[:12] = [0, 1, 15, 16, 17, 31, 32, 33, 500, 50, 275, 2]
        pairs = [(np.zeros(0, np.uint8), synth_tokens(27, k, 1, lo, fm.n_out)) for k, lo in enumerate(lens)]
        orc = Oracle(fm)
        m = make_machine(capi, fm, 2, **opts)
        b = capi.Batch(pairs)
        ll = capi.forward(m, b)
        launches = b.last_kernel_ms()[1]
        assert launches >= 3 and launches % 3 == 0, launches      # prefix, strips, suffix per chunk: the column engine ran
        if "col_bnd_budget_mb" in opts:
            assert launches >= 6
        assert b.last_redo() == 0
        sc = capi.viterbi(m, b, paths=False)
        lane = make_machine(capi, fm, 2, no_col=1)
        ll_lane = capi.forward(lane, b)
        sc_lane = capi.viterbi(lane, b, paths=False)
        assert np.array_equal(sc, sc_lane)
        np.testing.assert_allclose(ll, ll_lane, rtol=1e-10)
        sc2, paths = capi.viterbi(m, b)      # with paths: the sweep that stores pointers, and the walk back over them
        assert np.array_equal(sc2, sc)
        launches = b.last_kernel_ms()[1]
        assert launches >= 5 and launches % 5 == 0, launches      # prefix, strips, suffix, two traceback passes per chunk
        if "col_bp_budget_mb" in opts:
            assert launches >= 10
        for k in sorted(set(list(range(12)) + [n_reads - 1] + list(range(14, n_reads, 31)))):
            x, y = pairs[k]
            f = orc.forward(x, y)
            assert forward_agrees(fm, x, y, ll[k], f), (name, k, ll[k], f)
            v, p = orc.viterbi(x, y)
            assert sc[k] == v, (name, k, sc[k], v)
            assert paths[k].tolist() == p.tolist(), (name, k)
        _, paths_lane = capi.viterbi(lane, b)
        assert all(np.array_equal(a, c) for a, c in zip(paths, paths_lane))


@pytest.mark.parametrize("opts", [dict(), dict(col_c=1), dict(col_c=3, col_minblocks=1), dict(col_sil_regs=0, col_r=2)])
def test_column_engine_synthetic_profile(opts):
    """A periodic generator unlike the HMMER import (helpers.synthetic_profile): diagonal groups (a token consumed on the way
    to the next node), token-consuming self-loops, a begin hub entered with a token, N / C flanking states with their own
    dynamics in the prefix and suffix programs -- the kernel paths the PF00516 machines do not reach."""
    capi = _capi()
    fm = synthetic_profile(n_nodes=90)
    lens = [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 64, 150, 301] + [5 + (k * 13) % 120 for k in range(60)]
    pairs = [(np.zeros(0, np.uint8), synth_tokens(39, k, 1, lo, fm.n_out)) for k, lo in enumerate(lens)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 2, **opts)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    launches, redo = b.last_kernel_ms()[1], b.last_redo()
    assert launches - (1 if redo else 0) == 3, launches      # the column engine ran (+ the log-domain sweep for reads without a path: the empty read)
    sc = capi.viterbi(m, b, paths=False)
    sc2, paths = capi.viterbi(m, b)
    assert np.array_equal(sc, sc2)
    n_inf = 0
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        n_inf += int(np.isinf(f))
        assert (ll[k] == f) if np.isinf(f) else abs(ll[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, ll[k], f)
        v, p = orc.viterbi(x, y)
        assert sc[k] == v, (k, sc[k], v)
        assert paths[k].tolist() == p.tolist(), k
    assert redo == n_inf


def test_column_engine_hands_impossible_reads_to_the_log_domain():
    """A profile that cannot emit residue 1: reads containing it have no path.  The column engine's linear sweep returns
    nothing for them; they are counted as re-run and come back -inf from the lane engine's log-domain sweep, the others are
    untouched."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("hmmer_pf00516")["machine"])
    lw = np.where(fm.tout == 1, -np.inf, fm.lw)
    fm2 = FlatMachine(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, lw, fm.in_alphabet, fm.out_alphabet)
    pairs = []
    for k in range(40):
        y = synth_tokens(33, k, 1, 20 + k, fm.n_out)
        if k % 2:
            y = np.where(y == 1, 2, y).astype(np.uint8)
        pairs.append((np.zeros(0, np.uint8), y))
    m = make_machine(capi, fm2, 2)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    sc = capi.viterbi(m, b, paths=False)
    sc_p, paths = capi.viterbi(m, b)
    assert np.array_equal(sc, sc_p)
    assert all((len(paths[k]) == 0) == bool(np.isinf(sc[k])) for k in range(len(pairs)))
    orc = Oracle(fm2)
    n_inf = 0
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        if np.isinf(f):
            n_inf += 1
            assert ll[k] == f and sc[k] == f, (k, ll[k], sc[k])
        else:
            assert abs(ll[k] - f) <= 1e-9 * abs(f), (k, ll[k], f)
            assert sc[k] == orc.viterbi(x, y)[0]
    assert n_inf >= 10
    capi.forward(m, b)
    assert b.last_redo() == n_inf


def test_chunked_traceback_and_counts():
    """A scratch budget of a few MiB (options jit_tb_budget_mb / jit_f_budget_mb) forces the Viterbi back-pointers and
    the E-step's stored Forward values into several chunks of pairs: same results as in one piece, and as the oracle."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    shapes = [(300 - 11 * k, 280 + 9 * k) for k in range(12)]
    pairs = [(synth_tokens(93, k, 0, li, 4), synth_tokens(93, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    b = capi.Batch(pairs)
    m1 = make_machine(capi, fm, 1)
    sc1, paths1 = capi.viterbi(m1, b)
    cnt1, ll1 = capi.counts(m1, b)
    mc = make_machine(capi, fm, 1, jit_tb_budget_mb=1, jit_f_budget_mb=4)
    b2 = capi.Batch(pairs)
    sc, paths = capi.viterbi(mc, b2)
    cnt, ll = capi.counts(mc, b2)
    assert np.array_equal(sc, sc1) and np.array_equal(ll, ll1)
    np.testing.assert_allclose(cnt, cnt1, rtol=1e-12)
    for k in range(len(pairs)):
        assert paths[k].tolist() == paths1[k].tolist(), k
    # ids as bytes through the packed-path entry point (the chunks pack in pair order)
    score, plen, off = np.empty(len(pairs)), np.zeros(len(pairs), np.int64), np.zeros(len(pairs) + 1, np.int64)
    trans = np.zeros(sum(len(p) for p in paths), dtype=np.uint8)
    capi.viterbi_into(mc, b2, score, plen, off, trans)
    for k, p in enumerate(paths):
        assert trans[off[k]:off[k + 1]].tolist() == p.tolist(), k
    orc = Oracle(fm)
    want = np.zeros(fm.n_trans)
    for k, (x, y) in enumerate(pairs):
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and paths[k].tolist() == p.tolist(), k
        orc.counts(x, y, counts=want)
    np.testing.assert_allclose(cnt, want, rtol=REL, atol=1e-7)


def test_tokens_outside_the_alphabet_are_rejected():
    """A pre-tokenised batch is held to Tokenizer::tokenize's rule (eval.h:33-37): a token beyond the machine's
    alphabet, or a 0 (epsilon) in the data, fails the call instead of indexing past the device tables."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_small")["machine"])
    m = make_machine(capi, fm, -1)
    for bad in ([1, 2, 5], [1, 0, 2]):
        b = capi.Batch([(np.array(bad, np.uint8), np.array([1, 2], np.uint8))])
        for call in (capi.forward, capi.backward, capi.counts, lambda mm, bb: capi.viterbi(mm, bb, paths=False)):
            with pytest.raises(capi.MachineBossError, match="alphabet"):
                call(m, b)
    ok = capi.Batch([(np.array([1, 2, 4], np.uint8), np.array([1, 2], np.uint8))])
    assert np.isfinite(capi.forward(m, ok)[0])


@pytest.mark.parametrize("engine", ENGINES)
def test_silent_self_loop_on_the_start_state(engine):
    """isAdvancingMachine exempts state 0 (machine.cpp:759): a silent 0 -> 0 transition is legal.  In the
    reference's fill it reads the cell under construction, still -inf, so it never contributes; every engine
    must agree (the generic engine once read the uninitialised cell there)."""
    capi = _capi()
    base = FlatMachine.from_json(load_golden("unitindel")["machine"])
    ins = int(np.searchsorted(base.src, 1))      # after state 0's transitions
    def with_loop(a, v):
        return np.insert(a, ins, v)
    fm = FlatMachine(base.n_states, base.n_in, base.n_out, with_loop(base.src, 0).astype(np.int32), with_loop(base.dst, 0).astype(np.int32),
                     with_loop(base.tin, 0).astype(np.int32), with_loop(base.tout, 0).astype(np.int32), with_loop(base.lw, math.log(0.5)),
                     base.in_alphabet, base.out_alphabet)
    pairs = [(synth_tokens(3, k, 0, li, max(1, fm.n_in)), synth_tokens(3, k, 1, lo, max(1, fm.n_out))) for k, (li, lo) in enumerate([(2, 3), (0, 0), (5, 4), (40, 37)])]
    orc = Oracle(fm)
    m = make_machine(capi, fm, engine)
    b = capi.Batch(pairs)
    ll, bl = capi.forward(m, b), capi.backward(m, b)
    sc, paths = capi.viterbi(m, b)
    cnt, _ = capi.counts(m, b)
    want = np.zeros(fm.n_trans)
    for k, (x, y) in enumerate(pairs):
        # the reference's FULL matrices (ForwardMatrix, the API's storage): its rolling matrix reads row o-2's stale cell there
        assert close(ll[k], orc.forward(x, y, matrix=True)[0]) and close(bl[k], orc.backward(x, y, matrix=True)[0]), (k, ll[k], bl[k])
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and paths[k].tolist() == p.tolist(), k
        orc.counts(x, y, counts=want)
    # getCounts multiplies the finished F and B cells through the loop (backward.cpp:76-84): a number no path
    # produces; here the loop, which no path can take, counts 0
    keep = np.arange(fm.n_trans) != ins
    np.testing.assert_allclose(cnt[keep], want[keep], rtol=REL, atol=1e-7)
    assert cnt[ins] == 0


def test_machines_from_the_kernel_cache_give_the_same_results(tmp_path):
    """mb_set_kernel_cache_dir: the second machine of a structure is loaded from the cubin the first one left in the directory -- and
    sweeps, traces back and counts like a freshly compiled one (JIT engine, its fitted module, and the column engine)."""
    import os
    import time
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    pairs = [(synth_tokens(97, k, 0, 300 - (k % 5), 4), synth_tokens(97, k, 1, 280 + (k % 9), 4)) for k in range(300)]
    b = capi.Batch(pairs)
    prof = synthetic_profile(n_nodes=70)
    reads = capi.Batch([(np.zeros(0, np.uint8), synth_tokens(98, k, 1, 20 + k, prof.n_out)) for k in range(40)])
    capi.set_kernel_cache_dir(str(tmp_path))
    try:
        out = []
        for trial in range(2):
            t0 = time.time()
            m = make_machine(capi, fm, 1)
            p = make_machine(capi, prof, 2)
            made = time.time() - t0
            ll = capi.forward(m, b)      # (300 uniform pairs: the fitted module is compiled -- or found -- here)
            sc, paths = capi.viterbi(m, b)
            cnt, _ = capi.counts(m, b)
            out.append((ll, sc, paths, cnt, capi.forward(p, reads), capi.viterbi(p, reads, paths=False), made))
            m.close(); p.close()
        n_files = len([f for f in os.listdir(str(tmp_path)) if f.endswith(".cubin")])
    finally:
        capi.set_kernel_cache_dir(None)
    assert n_files >= 4      # the JIT engine's two modules, the fitted one, the column engine's
    a, c = out
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[4], c[4]) and np.array_equal(a[5], c[5])
    np.testing.assert_allclose(a[3], c[3], rtol=1e-12)      # (counts are summed with atomics: equal to rounding, run to run)
    assert all(np.array_equal(x, y) for x, y in zip(a[2], c[2]))
    assert c[6] < 0.5 * a[6], (a[6], c[6])      # creating the machines again took less than half as long


def test_group_handles_may_be_destroyed_in_any_order():
    """An interpreter at exit (or a caller's error path) may destroy the group before its machines and batches: the group is
    kept alive until its last child is gone."""
    capi = _capi()
    fm, pairs = _group_case()
    g = capi.Group()
    gm = capi.GroupMachine(g, fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    gb = capi.GroupBatch(g, pairs[:8])
    ll = capi.group_forward(gm, gb)
    g.close()      # first the group ...
    ll2 = capi.group_forward(gm, gb)      # ... whose children still work
    assert np.array_equal(ll, ll2)
    gb.close()
    gm.close()


def _group_case():
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    shapes = [(40 + (37 * k) % 300, 30 + (53 * k) % 280) for k in range(41)] + [(0, 0), (0, 5), (600, 580)]
    pairs = [(synth_tokens(71, k, 0, li, 4), synth_tokens(71, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    return fm, pairs


def _check_group_against_single(capi, devices):
    fm, pairs = _group_case()
    m1 = make_machine(capi, fm, -1)
    b1 = capi.Batch(pairs)
    ll1, (sc1, paths1), (cnt1, cll1) = capi.forward(m1, b1), capi.viterbi(m1, b1), capi.counts(m1, b1)
    g = capi.Group(devices)
    gm = capi.GroupMachine(g, fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    gb = capi.GroupBatch(g, pairs)
    dev, cells = gb.shard()
    assert len(dev) == len(pairs) and set(dev.tolist()) <= set(range(64))
    if g.n_devices > 1:
        biggest = max((len(x) + 1.0) * (len(y) + 1.0) for x, y in pairs)      # (one pair cannot be split: on 8 devices it outweighs a device's share)
        assert len(set(dev.tolist())) == g.n_devices and cells.max() <= max(1.2 * cells.mean(), 1.0001 * biggest)      # every device has work, evenly
    ll = capi.group_forward(gm, gb)
    sc, paths = capi.group_viterbi(gm, gb)
    cnt, cll = capi.group_counts(gm, gb)
    assert np.array_equal(sc, sc1)      # Viterbi: bit for bit
    # the sums: to rounding (the strip width is chosen per call from the pairs at hand, and a shard is not the whole list:
    # another width adds the same numbers up in another order)
    np.testing.assert_allclose(ll, ll1, rtol=1e-12)
    np.testing.assert_allclose(cll, cll1, rtol=1e-12)
    for k in range(len(pairs)):
        assert paths[k].tolist() == paths1[k].tolist(), k
    np.testing.assert_allclose(cnt, cnt1, rtol=1e-12, atol=1e-300)      # summed in a different order across devices
    assert abs(gm.last_loglike() - float(cll1.sum())) <= 1e-12 * abs(float(cll1.sum()))
    # a second E-step with other weights (what an EM iteration does): replicas updated on every device
    flat = FlatMachine.from_json(load_golden("dnapsw_synth64")["machine"])
    gm.update_weights(flat.lw)
    m1.update_weights(flat.lw)
    cnt, _ = capi.group_counts(gm, gb)
    cnt1, _ = capi.counts(m1, b1)
    np.testing.assert_allclose(cnt, cnt1, rtol=1e-12, atol=1e-300)
    return g


def test_group_of_one_device_equals_single_device_calls():
    """mb_group_* with a single device: the same results as the plain entry points (threads, gather and the count
    reduction are the group's own; no NCCL with one device)."""
    capi = _capi()
    g = _check_group_against_single(capi, [0])
    assert g.n_devices == 1 and not g.uses_nccl


def test_group_over_all_devices_equals_single_device():
    """The list dealt over every GPU of the box: per-pair results identical to one device's, counts equal to 1e-12
    after the NCCL all-reduce.  Skipped below two devices."""
    capi = _capi()
    if capi.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    g = _check_group_against_single(capi, None)
    assert g.n_devices == capi.device_count()


@pytest.mark.parametrize("split", [1, 0])
def test_split_mode_strips_as_work_items(split):
    """With fewer pairs than resident warps the strips of a pair are work items of their own and the warps that claim them
    run as a pipeline down the strips, each waiting for the rows of its left neighbour's boundary (option jit_split: 1
    forces it, 0 forbids it; the default decides by the batch).  Pairs of 1 to 12 strips, ragged: every score pass and the
    traceback against the oracle, and nothing handed to the log domain."""
    capi = _capi()
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    shapes = [(3000, 2800), (600, 700), (257, 300), (100, 50), (1025, 900), (0, 7), (256, 31), (255, 1000), (511, 40)]
    pairs = [(synth_tokens(83, k, 0, li, 4), synth_tokens(83, k, 1, lo, 4)) for k, (li, lo) in enumerate(shapes)]
    orc = Oracle(fm)
    m = make_machine(capi, fm, 1, jit_split=split)
    b = capi.Batch(pairs)
    ll = capi.forward(m, b)
    assert b.last_redo() == 0
    bl = capi.backward(m, b)
    assert b.last_redo() == 0
    sc, paths = capi.viterbi(m, b)
    sc2 = capi.viterbi(m, b, paths=False)
    for k, (x, y) in enumerate(pairs):
        f = orc.forward(x, y, mode=LSE_EXACT)
        assert abs(ll[k] - f) <= 1e-9 * max(1.0, abs(f)) and abs(bl[k] - f) <= 1e-9 * max(1.0, abs(f)), (k, ll[k], bl[k], f)
        v, p = orc.viterbi(x, y)
        assert sc[k] == v and sc2[k] == v, (k, sc[k], v)
        assert paths[k].tolist() == p.tolist(), k
    # the default picks the split for a batch this small: same numbers
    m_auto = make_machine(capi, fm, 1)
    assert np.array_equal(capi.forward(m_auto, b), ll) and np.array_equal(capi.viterbi(m_auto, b, paths=False), sc)
