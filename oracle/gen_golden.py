#!/usr/bin/env python3
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE: writes tests/golden/*.json.

Runs the UNMODIFIED reference (oracle/_ref/refdrv, built by `make -C oracle ref` from the sources
under /root/reference) on a fixed list of cases and stores, per case: the flat evaluated machine,
the token sequences, and the reference's Forward / rolling Forward / Backward / Viterbi values,
Viterbi traceback (global transition ids) and raw posterior counts at 17 significant digits.
Where the reference's own test-suite pins a value for the case (t/expect/*, Makefile:492-572,
js/webgpu/test/test-cpu.mjs:129-197) that published value is stored too under "ref_expect", so
tests can check the oracle against the reference's goldens and not only against the binary.

This needs /root/reference and therefore runs only in the build container; the fixtures are
committed so nothing on the GPU box reads the reference.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import mutate_tokens, synth_tokens  # noqa: E402

REF = os.environ.get("MB_REFERENCE", "/root/reference")
REFDRV = os.path.join(HERE, "_ref", "refdrv")
OUT = os.path.join(REPO, "tests", "golden")


def run(args):
    r = subprocess.run([REFDRV] + args, check=True, capture_output=True, text=True)
    return json.loads(r.stdout)


def machine_args(specs, params):
    a = []
    for s in specs:
        a += ["--machine", s]
    if params is not None:
        f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump(params, f)
        f.close()
        a += ["--params", f.name]
    return a


def case(name, specs, pairs, params=None, do="forward,rolling,viterbi,path,backward,counts", matrices=False,
         ref_expect=None, note="", gz=False, machine_from=None):
    """pairs: list of (in_symbols, out_symbols), ("synth", N, Li, Lo, seed), or ("tokens", [(x, y), ...]) with
    1-based token arrays.  machine_from: name of another fixture holding the same machine (stored once)."""
    margs = machine_args(specs, params)
    mach = run(margs + ["--emit-machine"])
    in_alpha, out_alpha = mach["inAlphabet"], mach["outAlphabet"]
    sym_pairs = []
    synth = None
    if pairs and pairs[0] == "synth":
        _, n, li, lo, seed = pairs
        synth = {"n": n, "li": li, "lo": lo, "seed": seed}
        for k in range(n):
            x = synth_tokens(seed, k, 0, li if in_alpha else 0, max(1, len(in_alpha)))
            y = synth_tokens(seed, k, 1, lo if out_alpha else 0, max(1, len(out_alpha)))
            sym_pairs.append(([in_alpha[t - 1] for t in x], [out_alpha[t - 1] for t in y]))
    elif pairs and pairs[0] == "tokens":
        sym_pairs = [([in_alpha[t - 1] for t in x], [out_alpha[t - 1] for t in y]) for x, y in pairs[1]]
    elif pairs and isinstance(pairs[0], dict):      # raw SeqPair JSON carrying an "alignment" (=> path envelope)
        raw_pairs = pairs
        sym_pairs = [([c[0] for c in p["alignment"] if c[0]], [c[1] for c in p["alignment"] if c[1]]) for p in pairs]
    else:
        sym_pairs = [(list(a), list(b)) for a, b in pairs]
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    if pairs and isinstance(pairs[0], dict):
        json.dump(raw_pairs, f)
    else:
        json.dump([{"input": {"name": "x%d" % k, "sequence": a}, "output": {"name": "y%d" % k, "sequence": b}}
                   for k, (a, b) in enumerate(sym_pairs)], f)
    f.close()
    if matrices:
        do = do + ",matrices"
    res = run(margs + ["--pairs", f.name, "--do", do])
    os.unlink(f.name)
    in_tok = {s: i + 1 for i, s in enumerate(in_alpha)}
    out_tok = {s: i + 1 for i, s in enumerate(out_alpha)}
    out_pairs = []
    for k, ((a, b), r) in enumerate(zip(sym_pairs, res["pairs"])):
        p = {"x": [in_tok.get(s, 0) for s in a], "y": [out_tok.get(s, 0) for s in b]}
        p.update(r)
        if pairs and isinstance(pairs[0], dict):
            p["alignment"] = pairs[k]["alignment"]
        out_pairs.append(p)
    if machine_from:
        with_machine = json.load(__import__("gzip").open(os.path.join(OUT, machine_from + ".json.gz"), "rt")) if os.path.exists(os.path.join(OUT, machine_from + ".json.gz")) \
            else json.load(open(os.path.join(OUT, machine_from + ".json")))
        assert with_machine["machine"] == mach, "machine_from does not hold the same machine"
    j = {"name": name, "note": note, "specs": specs, "params": params, "synth": synth,
         "pairs": out_pairs, "loglike": res.get("loglike"), "counts": res.get("counts"),
         "ref_expect": ref_expect or {}}
    if machine_from:
        j["machine_from"] = machine_from
    else:
        j["machine"] = mach
    if gz:
        import gzip
        with gzip.open(os.path.join(OUT, name + ".json.gz"), "wt", compresslevel=9) as fo:
            json.dump(j, fo, separators=(",", ":"))
    else:
        with open(os.path.join(OUT, name + ".json"), "w") as fo:
            json.dump(j, fo, separators=(",", ":"))
    print("%-28s S=%-5d T=%-6d pairs=%d" % (name, mach["nStates"], len(mach["trans"]), len(out_pairs)))


def T(rel):
    return "file:" + os.path.join(REF, rel)


def expect_matrix(rel):
    """t/expect/{fwd,back}-bitnoise-params-tiny.json: cells listed as {inPos,outPos,state,logLike}."""
    with open(os.path.join(REF, rel)) as f:
        txt = f.read()
    # the reference prints non-JSON "-inf" for log(0) (dpmatrix.defs.h:52)
    j = json.loads(txt.replace("-inf", '"-Infinity"'))
    return [[c["inPos"], c["outPos"], c["logLike"]] for c in j["cell"]]


def main():
    os.makedirs(OUT, exist_ok=True)
    pq99 = {"p": 0.99, "q": 0.01}

    # --- the reference's own DP goldens (Makefile:492-500, 515-516, 567-572) ---
    with open(os.path.join(REF, "t/expect/fwdback-bitnoise-params-tiny.json")) as f:
        fb = json.load(f)
    case("bitnoise_tiny", [T("t/machine/bitnoise.json")], [("001", "101")], params=pq99, matrices=True,
         ref_expect={"forward_matrix_5dp": expect_matrix("t/expect/fwd-bitnoise-params-tiny.json"),
                     "backward_matrix_5dp": expect_matrix("t/expect/back-bitnoise-params-tiny.json"),
                     "counts": fb[0], "loglike_4sf": -4.625},
         note="Makefile:493-500,567 test-fwd/back/fb-bitnoise-params-tiny, test-101-bitnoise-001")
    case("bitnoise_counts", [T("t/machine/bitnoise.json")], [("101", "001")], params=pq99,
         ref_expect={"param_counts": {"p": 2, "q": 1}}, note="Makefile:518-522 test-counts / test-counts2")
    with open(os.path.join(REF, "t/expect/align-stutter-noise-difflen.json")) as f:
        al = json.load(f)
    case("stutter_noise_difflen", [T("t/machine/bitstutter.json"), T("t/machine/bitnoise.json")], [("01", "101")],
         params=pq99, matrices=True,
         ref_expect={"align_output": open(os.path.join(REF, "t/expect/align-stutter-noise-difflen.json")).read(),
                     "path_to": [t["to"] for t in al[0]["meta"]["path"]["trans"]],
                     "path_in": [t.get("in", "") for t in al[0]["meta"]["path"]["trans"]],
                     "path_out": [t.get("out", "") for t in al[0]["meta"]["path"]["trans"]]},
         note="Makefile:515-516 test-align-stutter-noise (boss a b -P params -D difflen -A)")
    case("bitstutternoise_0011", [T("t/machine/bitstutter-noise.json")], [("101", "0011")], params=pq99,
         ref_expect={"forward_3sf": -9.26, "viterbi_3sf": -9.27}, note="Makefile:570-572")
    case("counter_xxx", [T("t/machine/counter.json")], [("", "xxx")], ref_expect={"param_counts": {"p": 3}},
         note="Makefile:524-526 test-counts3")
    # --- known answers quoted in js/webgpu/test/test-cpu.mjs:129-197 ("boss gives ...") ---
    case("bitnoise_p09", [T("t/machine/bitnoise.json")], [("001", "101")], params={"p": 0.9, "q": 0.1},
         ref_expect={"forward_1e-4": -2.51331, "viterbi_1e-4": -2.51331}, note="test-cpu.mjs:125-131,178-184")
    case("bitnoise_p001", [T("t/machine/bitnoise.json")], [("001", "101")], params={"p": 0.01, "q": 0.99},
         ref_expect={"forward_1e-4": -9.22039}, note="test-cpu.mjs:134-140")
    case("bitecho", [T("t/machine/bitecho.json")], [("101", "101"), ("101", "001")],
         ref_expect={"forward_exact": [0.0, "-Infinity"]}, note="test-cpu.mjs:145-162 (second pair is impossible: -inf)")
    case("unitindel", [T("t/machine/unitindel.json")], [("xx", "xxx"), ("", ""), ("x", ""), ("", "x"), ("xxxx", "x")],
         params={"ins": 0.1, "no_ins": 0.9, "del": 0.1, "no_del": 0.9}, matrices=True,
         ref_expect={"forward_1e-3": -1.6869, "viterbi_1e-3": -2.82939}, note="test-cpu.mjs:167-197; plus empty / ragged pairs")

    # --- presets at the BASELINE configs' machines, pinned by running the reference itself ---
    case("dnapsw_small", ["preset:dnapsw"], [("ACGT", "ACGT"), ("", ""), ("A", ""), ("", "G"), ("ACGTACGTAC", "AGT"),
                                             ("AAAAAAAA", "AAAAAAAA"), ("ACGT", "TGCA")], matrices=True,
         note="tiny dnapsw incl. empty and ragged pairs and tie-heavy homopolymers; -U default params")
    case("dnapsw_synth64", ["preset:dnapsw"], ("synth", 4, 64, 61, 101))
    case("dnapsw_synth300", ["preset:dnapsw"], ("synth", 2, 300, 257, 102))
    peaked = {"gapOpen": 0.05, "gapExtend": 0.5, "eqmA": 0.25, "eqmC": 0.25, "eqmG": 0.25, "eqmT": 0.25}
    for a in "ACGT":
        for b in "ACGT":
            peaked["sub%s%s" % (a, b)] = 0.91 if a == b else 0.03
    case("dnapsw_peaked", ["preset:dnapsw"], ("synth", 3, 120, 130, 103), params=peaked,
         note="SURVEY 8(d) parameter set B (peaked), stresses the log-sum-exp cutoff")
    case("protpsw_synth", ["preset:protpsw"], ("synth", 3, 50, 47, 104))
    case("protpsw_300", ["preset:protpsw"], ("synth", 1, 300, 300, 105), do="forward,rolling,viterbi,path,backward,counts")
    case("prot2dna_dnapsw", ["preset:prot2dna", "preset:dnapsw"], ("synth", 2, 7, 24, 106),
         note="config-4 style composite (S=308)")
    case("translate", ["preset:translate"], ("synth", 2, 9, 27, 107))
    case("dnapsw_dnapsw", ["preset:dnapsw", "preset:dnapsw"], ("synth", 4, 40, 37, 111),
         note="dnapsw => dnapsw: a mid-size composite WITH match transitions (the big engine's diagonal path)")
    envelope_case()
    env_expect = {}
    for nm in ("tinypath_full", "tinypath_path", "smallpath_path", "smallpath_area0", "smallpath_area1", "smallpath_area2",
               "smallpath_area3", "smallpath_area4", "asympath_area0", "asympath_area1"):
        env_expect[nm] = json.load(open(os.path.join(REF, "t/expect/%s_env.json" % nm)))
    with open(os.path.join(OUT, "aux_envelopes.json"), "w") as fo:
        json.dump({"note": "Makefile:450-462 test-env: inputs t/io/{tinypath,smallpath,asympath}.json and the expected envelopes",
                   "tinypath": json.load(open(os.path.join(REF, "t/io/tinypath.json"))),
                   "smallpath": json.load(open(os.path.join(REF, "t/io/smallpath.json"))),
                   "asympath": json.load(open(os.path.join(REF, "t/io/asympath.json"))), "expect": env_expect}, fo)

    # --- EM / M-step inputs of the reference's own tests (Makefile:502-504) and the dnapsw preset, as text,
    #     for the host mirror's symbolic layer; the expected fit is the reference's golden ---
    def rd(rel):
        return open(os.path.join(REF, rel)).read()
    with open(os.path.join(OUT, "aux_fit_inputs.json"), "w") as fo:
        json.dump({"note": "inputs of test-fit-bitnoise-seqpairlist (Makefile:502-504), test-counts (518-522) and the dnapsw preset",
                   "bitnoise_machine": rd("t/machine/bitnoise.json"), "pqcons": rd("t/io/pqcons.json"),
                   "seqpairlist": rd("t/io/seqpairlist.json"), "params": rd("t/io/params.json"),
                   "expect_fit_bitnoise_seqpairlist": rd("t/expect/fit-bitnoise-seqpairlist.json"),
                   "expect_counts": rd("t/expect/counts.json"),
                   "dnapsw_machine": rd("preset/dnapsw.json")}, fo)

    # --- config 5 style: an HMMER3 profile (examples/PF00516.hmm, local core machine of src/hmmer.cpp),
    #     alone and composed with an error model; large state spaces, generator machines (no input) ---
    hmm = "hmmer:" + os.path.join(REF, "examples/PF00516.hmm")
    case("hmmer_pf00516", [hmm], ("synth", 3, 0, 37, 108), gz=True, note="PF00516 local core, S=2439, T=24367")
    case("hmmer_pf00516_protpsw", [hmm, "preset:protpsw"], ("synth", 1, 0, 21, 109), gz=True,
         note="BASELINE config 5 machine: PF00516 => protpsw, S=12176, T=74012")
    round2_cases()


def peaked_dna():
    p = {"gapOpen": 0.05, "gapExtend": 0.5, "eqmA": 0.25, "eqmC": 0.25, "eqmG": 0.25, "eqmT": 0.25}
    for a in "ACGT":
        for b in "ACGT":
            p["sub%s%s" % (a, b)] = 0.91 if a == b else 0.03
    return p


def peaked_protein():
    aa = "ACDEFGHIKLMNPQRSTVWY"
    p = {"gapOpen": 0.05, "gapExtend": 0.5}
    for a in aa:
        p["eqm" + a] = 0.05
        for b in aa:
            p["sub%s%s" % (a, b)] = 0.62 if a == b else 0.02
    return p


def round2_cases():
    """Fixtures at the BASELINE configs' own sizes and on hard inputs ('next round' item 1)."""
    hmm = "hmmer:" + os.path.join(REF, "examples/PF00516.hmm")
    # (i) counts at config sizes
    case("dnapsw_1k", ["preset:dnapsw"], ("synth", 1, 1000, 1000, 12345), do="forward,rolling,viterbi,path,backward,counts",
         note="BASELINE configs 1 and 2: one 1 kb x 1 kb pair, every pass incl. posterior counts")
    case("protpsw_300_batch", ["preset:protpsw"], ("synth", 4, 300, 300, 205), gz=True,
         note="BASELINE config 3 size: 300 aa pairs, every pass incl. counts")
    # (ii) parameter set B: peaked weights, output = input mutated 10 % / 2 %
    xs = [synth_tokens(301, k, 0, 1000, 4) for k in range(2)]
    case("dnapsw_peaked_1k", ["preset:dnapsw"], ("tokens", [(x, mutate_tokens(301, k, x, 4)) for k, x in enumerate(xs)]),
         params=peaked_dna(), gz=True, note="set B at config size: 1 kb pairs, output = input with 10 % substitutions, 2 % indels")
    xs = [synth_tokens(302, k, 0, 300, 20) for k in range(3)]
    case("protpsw_peaked_300", ["preset:protpsw"], ("tokens", [(x, mutate_tokens(302, k, x, 20)) for k, x in enumerate(xs)]),
         params=peaked_protein(), gz=True, note="set B, protpsw: 300 aa pairs, mutated copies")
    # (iii) config 4 at a config-like shape
    case("prot2dna_dnapsw_2k", ["preset:prot2dna", "preset:dnapsw"], ("synth", 1, 300, 2000, 206), do="rolling,viterbi,path", gz=True,
         machine_from="prot2dna_dnapsw", note="BASELINE config 4 machine, one pair of 300 aa x 2 kb")
    # (iv) config 5 at its read lengths
    for nm, specs, seed in (("hmmer_pf00516_reads", [hmm], 207), ("hmmer_pf00516_protpsw_reads", [hmm, "preset:protpsw"], 208)):
        reads = [(synth_tokens(seed, k, 0, 0, 1), synth_tokens(seed, k, 1, lo, 20)) for k, lo in enumerate((50, 275, 500))]
        case(nm, specs, ("tokens", reads), do="rolling,viterbi,path", gz=True, machine_from=nm[:-6],
             note="BASELINE config 5: reads of 50, 275 and 500 residues")
    # (v) an E-step pair beyond the linear count kernels' range (log-domain FP32-accumulator path)
    case("dnapsw_3k", ["preset:dnapsw"], ("synth", 1, 3300, 3100, 209), do="forward,backward,counts,viterbi", gz=True,
         note="a pair longer than 3 kb: Forward, Backward, posterior counts, Viterbi score")


def envelope_case():
    peaked = {"gapOpen": 0.05, "gapExtend": 0.5, "eqmA": 0.25, "eqmC": 0.25, "eqmG": 0.25, "eqmT": 0.25}
    for a in "ACGT":
        for b in "ACGT":
            peaked["sub%s%s" % (a, b)] = 0.91 if a == b else 0.03
    # --- path envelopes: pairs that carry an alignment get Envelope::initPath (seqpair.cpp:104-110,134-152) ---
    import random
    random.seed(5)

    def aligned_pair(n, name):
        cols = []
        for _ in range(n):
            r = random.random()
            a, b = random.choice("ACGT"), random.choice("ACGT")
            cols.append([a, a] if r < 0.6 else [a, b] if r < 0.8 else [a, ""] if r < 0.9 else ["", b])
        return {"input": {"name": name + "x"}, "output": {"name": name + "y"}, "alignment": cols}
    case("dnapsw_path_envelope", ["preset:dnapsw"], [aligned_pair(12, "a"), aligned_pair(40, "b"), aligned_pair(1, "c"),
                                                     {"input": {"name": "dx"}, "output": {"name": "dy"}, "alignment": [["A", ""], ["", "C"], ["G", "G"]]}],
         params=peaked, matrices=True, note="pairs with an alignment: every matrix is restricted to the path envelope")


def sample_case():
    """ForwardMatrix::samplePath (forward.cpp:17-19): paths drawn by the reference with mt19937 (seed + pair index)."""
    peaked = {"gapOpen": 0.05, "gapExtend": 0.5, "eqmA": 0.25, "eqmC": 0.25, "eqmG": 0.25, "eqmT": 0.25}
    for a in "ACGT":
        for b in "ACGT":
            peaked["sub%s%s" % (a, b)] = 0.91 if a == b else 0.03
    margs = machine_args(["preset:dnapsw"], peaked)
    mach = run(margs + ["--emit-machine"])
    alpha = mach["inAlphabet"]
    shapes = [(12, 14), (30, 26), (0, 3), (5, 0), (40, 41)]
    sym_pairs = [([alpha[t - 1] for t in synth_tokens(110, k, 0, li, 4)], [alpha[t - 1] for t in synth_tokens(110, k, 1, lo, 4)])
                 for k, (li, lo) in enumerate(shapes)]
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k, "sequence": a}, "output": {"name": "y%d" % k, "sequence": b}} for k, (a, b) in enumerate(sym_pairs)], f)
    f.close()
    seed = 20261017
    res = run(margs + ["--pairs", f.name, "--do", "sample", "--sample-seed", str(seed)])
    os.unlink(f.name)
    with open(os.path.join(OUT, "aux_sample_paths.json"), "w") as fo:
        json.dump({"note": "ForwardMatrix::samplePath of the reference: pair k drawn with mt19937(seed + k); global transition ids start -> end",
                   "machine": mach, "seed": seed,
                   "pairs": [{"input": a, "output": b, "forward": r["forward"], "sample": r["sample"]} for (a, b), r in zip(sym_pairs, res["pairs"])]},
                  fo, separators=(",", ":"))
    print("aux_sample_paths             pairs=%d" % len(sym_pairs))


def post_trans_case():
    """BackwardMatrix::postTransQueue (backward.cpp:52-56) and traceFrom with a TraceTerminator (backward.cpp:98-108): the twelve
    largest (cell, transition) posteriors of a pair and the transitions visited tracing from the first one's source cell."""
    margs = machine_args(["preset:dnapsw"], peaked_dna())
    mach = run(margs + ["--emit-machine"])
    alpha = mach["inAlphabet"]
    shapes = [(9, 11), (14, 12), (0, 3), (6, 6)]
    sym_pairs = [([alpha[t - 1] for t in synth_tokens(112, k, 0, li, 4)], [alpha[t - 1] for t in synth_tokens(112, k, 1, lo, 4)])
                 for k, (li, lo) in enumerate(shapes)]
    sym_pairs[3] = (sym_pairs[3][0], sym_pairs[3][0])      # identical sequences: a sharply peaked posterior
    f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([{"input": {"name": "x%d" % k, "sequence": a}, "output": {"name": "y%d" % k, "sequence": b}} for k, (a, b) in enumerate(sym_pairs)], f)
    f.close()
    res = run(margs + ["--pairs", f.name, "--do", "posttrans"])
    os.unlink(f.name)
    with open(os.path.join(OUT, "aux_post_trans.json"), "w") as fo:
        json.dump({"note": "BackwardMatrix::postTransQueue / traceFrom of the reference: [destination inPos, outPos, global transition id, posterior]; traceFrom = ids in visit order",
                   "machine": mach, "pairs": [dict(input=a, output=b, **r) for (a, b), r in zip(sym_pairs, res["pairs"])]}, fo, separators=(",", ":"))
    print("aux_post_trans               pairs=%d" % len(sym_pairs))


def downsample_case():
    """Machine::downsample (machine.cpp:2036-2082) on alignment lattices -- a generator and a recognizer of fixed sequences composed
    around a transducer, toposorted, as `boss ... --downsample-size` sees them (boss.cpp:487-490): which transitions the selection
    loop keeps (refdrv replays the loop with the reference's own primitives, the function only returns the pruned machine)."""
    cases = []
    for specs, params, size, prob in (
            (["generate:10110", "file:" + os.path.join(REF, "t/machine/bitstutter-noise.json"), "recognize:1001110"], os.path.join(REF, "t/io/params.json"), 0.3, 0.0),
            (["generate:10110", "file:" + os.path.join(REF, "t/machine/bitstutter-noise.json"), "recognize:1001110"], os.path.join(REF, "t/io/params.json"), 1.0, 0.02),
            (["generate:ACGTTGCA", "preset:dnapsw", "recognize:ACTTGGCA"], None, 0.3, 0.0),
            (["generate:GATTACAT", "preset:dnapsw", "recognize:GCTACATT"], None, 1.0, 1e-3)):
        a = []
        for sp in specs:
            a += ["--machine", sp]
        if params:
            a += ["--params", params]
        res = run(a + ["--downsample", "%g,%g" % (size, prob)])
        cases.append({"what": " => ".join(sp.split("/")[-1] for sp in specs), "size": size, "prob": prob, "machine": res["machine"],
                      "nTransitions": res["nTransitions"], "kept": res["kept"], "allowed": res["allowed"]})
        print("aux_downsample               %-60s %d of %d transitions kept" % (cases[-1]["what"][:60], res["kept"], res["nTransitions"]))
    # Machine::stochasticDownsample (machine.cpp:2084-2128): paths sampled with mt19937 (seed) -- by number of paths (boss
    # --downsample-path) and until a fraction of the transitions is covered, at most nStates paths (--downsample-frac)
    for specs, params, frac, paths, seed in (
            (["generate:ACGTTGCA", "preset:dnapsw", "recognize:ACTTGGCA"], None, 1.0, 5, 7),
            (["generate:10110", "file:" + os.path.join(REF, "t/machine/bitstutter-noise.json"), "recognize:1001110"], os.path.join(REF, "t/io/params.json"), 0.5, -1, 3)):
        a = []
        for sp in specs:
            a += ["--machine", sp]
        if params:
            a += ["--params", params]
        n_states = len(run(a + ["--downsample", "1,0"])["machine"]["state"])
        res = run(a + ["--downsample-path", "%g,%d,%d" % (frac, paths if paths > 0 else n_states, seed)])
        cases.append({"what": " => ".join(sp.split("/")[-1] for sp in specs), "stochastic": True, "size": frac, "paths": paths, "seed": seed, "machine": res["machine"],
                      "nTransitions": res["nTransitions"], "kept": res["kept"], "allowed": res["allowed"]})
        print("aux_downsample (sampled)     %-60s %d of %d transitions on %d paths" % (cases[-1]["what"][:60], res["kept"], res["nTransitions"], res["paths"]))
    with open(os.path.join(OUT, "aux_downsample.json"), "w") as fo:
        json.dump({"note": "Machine::downsample of the reference: allowed[state][transIndex] after the selection loop, on the toposorted machine given here", "cases": cases},
                  fo, separators=(",", ":"))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "downsample":
        downsample_case()
    elif len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "posttrans":
        post_trans_case()
    elif len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "sample":
        sample_case()
    elif len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "envelope":
        envelope_case()
    elif len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "round2":
        round2_cases()
    else:
        main()
        sample_case()
        post_trans_case()
        downsample_case()
