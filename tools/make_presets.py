#!/usr/bin/env python3
"""Writes machineboss_b200/presets/*.eval.json[.gz]: the machines of the five BASELINE configs, already evaluated
(`boss -U` default parameters), in the flat form EvaluatedMachine::fromJson reads.  They are the "machine" blocks of
the committed fixtures (tests/golden, produced by oracle/gen_golden.py from the reference), so that
`boss_b200 --preset dnapsw ... -L` runs where there is no reference to build the machine."""
import gzip
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import load_golden  # noqa: E402

OUT = os.path.join(REPO, "machineboss_b200", "presets")
os.makedirs(OUT, exist_ok=True)
for preset, fixture, gz in (("dnapsw", "dnapsw_synth64", False), ("protpsw", "protpsw_synth", False),
                            ("prot2dna_dnapsw", "prot2dna_dnapsw", True), ("PF00516", "hmmer_pf00516", True),
                            ("PF00516_protpsw", "hmmer_pf00516_protpsw", True)):
    m = load_golden(fixture)["machine"]
    m = dict(m, note="preset %s, evaluated with boss -U default parameters by the reference (fixture %s)" % (preset, fixture))
    path = os.path.join(OUT, preset + ".eval.json" + (".gz" if gz else ""))
    with (gzip.open(path, "wt", compresslevel=9) if gz else open(path, "w")) as f:
        json.dump(m, f, separators=(",", ":"))
    print(path, os.path.getsize(path))
