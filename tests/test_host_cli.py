"""The C++ host mirror (machineboss_b200/host) and its CLI: builds on CPU, fails loudly without a
GPU, and on the GPU reproduces the reference CLI's output (`boss -A`, `-L`, `-V`, `-C` layouts)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import FlatMachine, gnum, load_golden


def _cli():
    from machineboss_b200 import build
    return build.build_host()


def _machine_file(case):
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump(case["machine"], f)
    f.close()
    return f.name


def _pairs_file(case):
    m = case["machine"]
    ia, oa = [""] + m["inAlphabet"], [""] + m["outAlphabet"]
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k, "sequence": [ia[t] for t in p["x"]]},
                "output": {"name": "y%d" % k, "sequence": [oa[t] for t in p["y"]]}} for k, p in enumerate(case["pairs"])], f)
    f.close()
    return f.name


def test_host_cli_builds_and_fails_loudly_without_gpu():
    import torch
    cli = _cli()
    assert os.path.exists(cli)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    case = load_golden("bitnoise_tiny")
    r = subprocess.run([cli, "--evaluated-machine", _machine_file(case), "--input-chars", "001", "--output-chars", "101", "-L"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "CUDA" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_cli_align_matches_reference_golden():
    """Makefile:515-516: boss bitstutter bitnoise -P params -D difflen -A == t/expect/align-stutter-noise-difflen.json"""
    case = load_golden("stutter_noise_difflen")
    r = subprocess.run([_cli(), "--evaluated-machine", _machine_file(case), "--input-chars", "01", "--output-chars", "101", "-A"],
                       capture_output=True, text=True, check=True)
    assert json.loads(r.stdout) == json.loads(case["ref_expect"]["align_output"])
    assert r.stdout.strip() == case["ref_expect"]["align_output"].strip()      # byte-for-byte, as the reference's test harness diffs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dnapsw_small", "unitindel", "bitecho", "protpsw_synth"])
def test_cli_loglike_viterbi_counts(name):
    case = load_golden(name)
    mf, pf = _machine_file(case), _pairs_file(case)
    out = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf, "-L", "-V"], capture_output=True, text=True, check=True).stdout
    dec = json.JSONDecoder()
    fwd, end = dec.raw_decode(out)
    vit, _ = dec.raw_decode(out[end:].lstrip())
    for k, p in enumerate(case["pairs"]):
        assert fwd[k][0] == "x%d" % k and fwd[k][1] == "y%d" % k
        for got, key in ((fwd[k][2], "rolling"), (vit[k][2], "viterbi")):
            want = gnum(p[key])
            if np.isinf(want):
                assert got == "-Infinity"
            else:
                assert float(got) == float("%.6g" % want) or abs(float(got) - want) <= 1e-4 * abs(want)
    if all(not np.isinf(gnum(p["forward"])) for p in case["pairs"]):
        out = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf, "-C"], capture_output=True, text=True, check=True).stdout
        got = np.array([c for row in json.loads(out) for c in row])
        want = np.array([gnum(v) for v in case["counts"]])
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-5)      # 6 printed digits


@pytest.mark.gpu
def test_cli_untokenisable_pair_reports_minus_infinity():
    """boss.cpp:798,805: -L prints "-Infinity" for a pair the machine cannot tokenise; -C throws (boss.cpp:811-816)."""
    case = load_golden("dnapsw_small")
    mf = _machine_file(case)
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "--input-chars", "ACGN", "--output-chars", "ACGT", "-L"], capture_output=True, text=True, check=True)
    assert json.loads(r.stdout) == [["ACGN", "ACGT", "-Infinity"]]
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "--input-chars", "ACGN", "--output-chars", "ACGT", "-C"], capture_output=True, text=True)
    assert r.returncode != 0 and "tokenize" in r.stderr


@pytest.mark.gpu
def test_cli_sample_paths_match_the_reference():
    """ForwardMatrix::samplePath (forward.cpp:17-23) over the device's stored Forward matrix (mb_matrix): with the
    reference's generator (mt19937, seed + pair index) and its random_index, the drawn paths are the reference's."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_sample_paths.json")) as f:
        g = json.load(f)
    mf = _machine_file(g)
    pf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k, "sequence": p["input"]}, "output": {"name": "y%d" % k, "sequence": p["output"]}}
               for k, p in enumerate(g["pairs"])], pf)
    pf.close()
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf.name, "--sample-paths", str(g["seed"])], capture_output=True, text=True, check=True)
    got = json.loads(r.stdout)
    assert got == [p["sample"] for p in g["pairs"]]


@pytest.mark.gpu
def test_cli_post_trans_queue_and_trace_from_match_the_reference():
    """BackwardMatrix::postTransQueue (backward.cpp:52-56) and traceFrom with a TraceTerminator (backward.cpp:98-108) of the
    host mirror, over the device's stored Forward and Backward matrices, against the reference's own output (refdrv): the
    number of (cell, transition) posteriors, the largest ones, and -- where the first is not tied -- the transitions
    visited tracing back and forward from it."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_post_trans.json")) as f:
        g = json.load(f)
    mf = _machine_file(g)
    pf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k, "sequence": p["input"]}, "output": {"name": "y%d" % k, "sequence": p["output"]}}
               for k, p in enumerate(g["pairs"])], pf)
    pf.close()
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf.name, "--post-trans", "12"], capture_output=True, text=True, check=True)
    got = json.loads(r.stdout)
    assert len(got) == len(g["pairs"])
    traced = 0
    # The reference's Forward / Backward sums go through its log-sum-exp lookup table (logsumexp.h:14-48): its posteriors
    # carry that table's error (measured here: 6e-5 relative on these pairs), the device's sums are exact.
    tol = 3e-4
    for mine, ref in zip(got, g["pairs"]):
        assert mine["postTransCount"] == ref["postTransCount"]
        want = {(e[0], e[1], e[2]): e[3] for e in ref["postTrans"]}
        np.testing.assert_allclose([e[3] for e in mine["postTrans"]], [e[3] for e in ref["postTrans"]], rtol=tol)      # the weights, in order
        floor = ref["postTrans"][-1][3] * (1 + 2 * tol)
        for e in mine["postTrans"]:
            if e[3] > floor:      # (entries tied with the last one shown may be others of the same weight)
                assert (e[0], e[1], e[2]) in want and abs(want[(e[0], e[1], e[2])] - e[3]) <= tol * e[3], e
        if ref["postTrans"][0][3] > ref["postTrans"][1][3] * (1 + 1e-9) and mine["postTrans"][0][:3] == ref["postTrans"][0][:3]:      # same starting point
            assert mine["traceFrom"] == ref["traceFrom"]
            traced += 1
    assert traced >= 2


@pytest.mark.gpu
def test_cli_downsample_keeps_the_transitions_the_reference_keeps():
    """Machine::downsample's selection (machine.cpp:2036-2082) in the host mirror (downsampleTransitions: the null machine's Forward and
    Backward for the empty pair on the device, postTransQueue, traceFrom with the stop terminator) on four alignment lattices the
    reference built and toposorted: the same transitions kept, by proportion and by posterior threshold -- and, for the stochastic
    form (paths sampled through the Forward matrix), the same transitions visited for the same mt19937 seed."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_downsample.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        mf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump(c["machine"], mf)
        mf.close()
        if c.get("stochastic"):      # Machine::stochasticDownsample: paths drawn with the reference's mt19937 sequence
            how = ["--downsample-path", str(c["paths"])] if c["paths"] > 0 else ["--downsample-frac", repr(c["size"])]
            r = subprocess.run([_cli(), "--machine", mf.name, "-U", "--seed", str(c["seed"])] + how, capture_output=True, text=True, check=True)
        else:
            r = subprocess.run([_cli(), "--machine", mf.name, "-U", "--downsample-size", repr(c["size"]), "--downsample-prob", repr(c["prob"])],
                               capture_output=True, text=True, check=True)
        os.unlink(mf.name)
        got = json.loads(r.stdout)
        assert got["nTransitions"] == c["nTransitions"], c["what"]
        assert got["kept"] == c["kept"] and got["allowed"] == c["allowed"], (c["what"], got["kept"], c["kept"])


def test_cli_envelopes_match_the_reference_goldens():
    """Makefile:450-462 test-env: Envelope::initFull / initPath / initPathArea of the host mirror against
    t/expect/*_env.json (CPU only: no device involved)."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_envelopes.json")) as f:
        g = json.load(f)
    cases = [("tinypath", "full", "tinypath_full"), ("tinypath", "path", "tinypath_path"), ("smallpath", "path", "smallpath_path"),
             ("smallpath", "0", "smallpath_area0"), ("smallpath", "1", "smallpath_area1"), ("smallpath", "2", "smallpath_area2"),
             ("smallpath", "3", "smallpath_area3"), ("smallpath", "4", "smallpath_area4"), ("smallpath", "5", "smallpath_area4"),
             ("asympath", "0", "asympath_area0"), ("asympath", "1", "asympath_area1")]
    for inp, mode, want in cases:
        pf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump([g[inp]], pf)
        pf.close()
        r = subprocess.run([_cli(), "-D", pf.name, "--envelope", mode], capture_output=True, text=True, check=True)
        assert json.loads(r.stdout) == g["expect"][want], (inp, mode)


@pytest.mark.gpu
def test_cli_pairs_with_alignments_get_the_path_envelope():
    """A SeqPair that carries an alignment restricts every matrix to the path envelope (seqpair.cpp:104-110,
    dpmatrix.defs.h:3): -L, -V and -C through the host mirror against the reference's values for such pairs."""
    case = load_golden("dnapsw_path_envelope")
    pf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k}, "output": {"name": "y%d" % k}, "alignment": p["alignment"]} for k, p in enumerate(case["pairs"])], pf)
    pf.close()
    mf = _machine_file(case)
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf.name, "-L"], capture_output=True, text=True, check=True)
    for row, p in zip(json.loads(r.stdout.replace("-Infinity", '"-Infinity"')), case["pairs"]):
        assert abs(gnum(row[2]) - gnum(p["forward"])) <= 1e-5 * max(1.0, abs(gnum(p["forward"]))), (row, p["forward"])      # 6 significant digits printed
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf.name, "-V"], capture_output=True, text=True, check=True)
    for row, p in zip(json.loads(r.stdout.replace("-Infinity", '"-Infinity"')), case["pairs"]):
        assert abs(gnum(row[2]) - gnum(p["viterbi"])) <= 1e-5 * max(1.0, abs(gnum(p["viterbi"]))), (row, p["viterbi"])
    r = subprocess.run([_cli(), "--evaluated-machine", mf, "-D", pf.name, "-C"], capture_output=True, text=True, check=True)
    got = np.array([v for row in json.loads(r.stdout) for v in row], dtype=np.float64)
    want = np.array([gnum(v) for v in case["counts"]])
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def _protein_files(n, length, tmp):
    aa = "ACDEFGHIKLMNPQRSTVWY"
    rng = np.random.default_rng(3)
    xs = rng.integers(0, 20, size=(n, length))
    ys = rng.integers(0, 20, size=(n, length - 3))
    fa_in, fa_out, js = os.path.join(tmp, "in.fa"), os.path.join(tmp, "out.fa"), os.path.join(tmp, "list.json")
    with open(fa_in, "w") as f:
        for k in range(n):
            f.write(">x%d some description\n%s\n" % (k, "".join(aa[c] for c in xs[k])))
    with open(fa_out, "w") as f:
        for k in range(n):
            s = "".join(aa[c] for c in ys[k])
            f.write(">y%d\n%s\n%s\n" % (k, s[:40], s[40:]))      # wrapped lines
    with open(js, "w") as f:
        json.dump([{"input": {"name": "x%d" % k, "sequence": [aa[c] for c in xs[k]]}, "meta": {"note": [1, {"a": 'b"c'}]},
                    "output": {"sequence": [aa[c] for c in ys[k]], "name": "y%d" % k}} for k in range(n)], f)
    return fa_in, fa_out, js


def test_fast_ingest_reads_lists_and_fasta_without_a_device(tmp_path):
    """boss_b200_ingest.h: a SeqPairList JSON and a pair of FASTA files go straight to packed tokens (no device needed for
    --ingest-only): pair and residue counts, and the failures a Tokenizer would raise."""
    cli = _cli()
    fa_in, fa_out, js = _protein_files(50, 60, str(tmp_path))
    for args in (["-D", js], ["--paired-fasta", fa_in, fa_out]):
        r = subprocess.run([cli, "--preset", "protpsw"] + args + ["--ingest-only"], capture_output=True, text=True, check=True)
        got = json.loads(r.stdout)
        assert got["pairs"] == 50 and got["residues"] == 50 * (60 + 57), got
    bad = os.path.join(str(tmp_path), "bad.fa")
    open(bad, "w").write(">z\nACDZ\n")      # Z is not an amino acid of protpsw
    r = subprocess.run([cli, "--preset", "protpsw", "--paired-fasta", bad, bad, "--ingest-only"], capture_output=True, text=True)
    assert r.returncode != 0 and "Can't tokenize symbol Z" in r.stderr
    aligned = os.path.join(str(tmp_path), "aligned.json")
    json.dump([{"input": {"name": "a"}, "output": {"name": "b"}, "alignment": [["A", "A"]]}], open(aligned, "w"))
    r = subprocess.run([cli, "--preset", "protpsw", "-D", aligned, "--ingest-only"], capture_output=True, text=True)
    assert r.returncode != 0 and "alignment" in r.stderr


@pytest.mark.gpu
def test_fast_ingest_scores_equal_the_general_reader(tmp_path):
    """-L and -V through the packed path (--fast-ingest, --paired-fasta) print what the SeqPairList path prints."""
    cli = _cli()
    fa_in, fa_out, js = _protein_files(40, 50, str(tmp_path))
    want = subprocess.run([cli, "--preset", "protpsw", "-D", js, "-L", "-V"], capture_output=True, text=True, check=True).stdout
    fast = subprocess.run([cli, "--preset", "protpsw", "-D", js, "--fast-ingest", "-L", "-V"], capture_output=True, text=True, check=True).stdout
    fasta = subprocess.run([cli, "--preset", "protpsw", "--paired-fasta", fa_in, fa_out, "-L", "-V"], capture_output=True, text=True, check=True).stdout
    assert fast == want and fasta == want
