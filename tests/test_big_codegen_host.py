"""CPU test of the big engine's code generator (mb_big.cu): the generated cell functions -- one multiply-add (or
add + compare) per transition group, straight-line -- are compiled for the HOST and driven by a plain row-major
loop, then compared with the oracle.  Every transition gets weight 1 (log-weight 0), so the Forward value is the
log of the number of paths and the Viterbi score is 0 wherever a path exists: what is checked is the structure
the generator emits (groups, table offsets, silent order, origin, the live-up / left-going state maps), with no
device involved.  Parity with real weights is the job of the -m gpu tests."""
import math
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import LSE_EXACT, FlatMachine, Oracle, load_golden, synth_tokens

HARNESS = r"""
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __constant__ static
static inline double __longlong_as_double (long long v) { double d; memcpy (&d, &v, 8); return d; }
%(cells)s
int main (int argc, char** argv) {
  // argv: Li Lo x-tokens (0-based)... y-tokens...
  const int Li = atoi (argv[1]), Lo = atoi (argv[2]);
  std::vector<int> x, y;
  for (int i = 0; i < Li; ++i) x.push_back (atoi (argv[3 + i]));
  for (int o = 0; o < Lo; ++o) y.push_back (atoi (argv[3 + Li + o]));
  static const char present[] = "%(present)s", presentLin[] = "%(present_lin)s";
  std::vector<double> Elin (MB_NEMIT_LIN), Elog (MB_NEMIT);      // the sums read their own table: one token vector per fold class
  for (int q = 0; q < MB_NEMIT; ++q) Elog[q] = present[q] == '1' ? 0.0 : -INFINITY;
  for (int q = 0; q < MB_NEMIT_LIN; ++q) Elin[q] = presentLin[q] == '1' ? 1.0 : 0.0;
  for (int q = 0; q < %(nsil)d; ++q) { mb_big_sil[q] = 1.0; mb_big_sil_log[q] = 0.0; }
  for (int q = 0; q < %(nfold)d; ++q) mb_big_fold[q] = 1.0;
  for (int mode = 0; mode < 2; ++mode) {
    const double ZERO = mode ? -INFINITY : 0.0;
    std::vector<std::vector<double> > up (Li + 1, std::vector<double> (((MB_NLU > MB_NLU_LIN ? MB_NLU : MB_NLU_LIN) + 1) * 32, ZERO));
    std::vector<std::vector<double> > loPrev (Li + 1, std::vector<double> (MB_NLL, ZERO)), loCur = loPrev;
    double res = ZERO;
    for (int o = 0; o <= Lo; ++o) {
      for (int i = 0; i <= Li; ++i) {
        double L[MB_NLL], D[MB_NLL], out[MB_NLL];
        for (int j = 0; j < MB_NLL; ++j) { L[j] = i ? loCur[i - 1][j] : ZERO; D[j] = i ? loPrev[i - 1][j] : ZERO; }
        const int a = i ? x[i - 1] : 0, b = o ? y[o - 1] : 0;
        if (mode == 0) mb_big_cell (up[i].data(), L, D, out, a, b, i == 0 && o == 0, Elin.data(), res);
        else { unsigned pw[MB_NPW]; mb_big_cell_vit (up[i].data(), L, D, out, a, b, i == 0 && o == 0, Elog.data(), res, pw); }
        for (int j = 0; j < MB_NLL; ++j) loCur[i][j] = out[j];
      }
      loPrev = loCur;
    }
    printf ("%%.17g\n", mode ? res : log (res));
  }
  return 0;
}
"""


@pytest.mark.parametrize("name,shapes", [("dnapsw_dnapsw", [(5, 6), (0, 3), (4, 0), (7, 7)]), ("translate", [(6, 2), (3, 1)]),
                                         ("prot2dna_dnapsw", [(2, 7), (3, 9)])])
def test_generated_cells_on_the_host(name, shapes, monkeypatch, tmp_path):
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    dump = str(tmp_path / "gen")
    monkeypatch.setenv("MB_JIT_DUMP", dump)
    capi.jit_compile_check(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    src = open(dump + ".big.cu").read()
    cells = src[: src.index("#define MB_FULL")]
    cells = "\n".join(l for l in cells.splitlines() if not l.startswith("typedef "))
    present = re.search(r"// MB_EMIT_PRESENT ([01]+)", src).group(1)
    nsil = int(re.search(r"// MB_NSIL (\d+)", src).group(1))
    present_lin = re.search(r"// MB_EMIT_PRESENT_LIN ([01]+)", src).group(1)
    nfold = int(re.search(r"// MB_NFOLD (\d+)", src).group(1))
    if name == "prot2dna_dnapsw":      # with equal weights every destination's insert groups are proportional: one folded value each
        assert nfold > 0 and int(re.search(r"#define MB_NLU_LIN (\d+)", src).group(1)) < int(re.search(r"#define MB_NLU (\d+)", src).group(1))
    cpp = tmp_path / "harness.cpp"
    cpp.write_text(HARNESS % {"cells": cells, "present": present, "nsil": nsil, "present_lin": present_lin, "nfold": nfold})
    exe = str(tmp_path / "harness")
    subprocess.run(["g++", "-O1", "-std=c++14", "-o", exe, str(cpp)], check=True, capture_output=True, text=True)
    ones = fm.with_weights(np.zeros_like(fm.lw))      # every transition weight 1
    orc = Oracle(ones)
    for k, (li, lo) in enumerate(shapes):
        x = synth_tokens(77, k, 0, li, max(fm.n_in, 1)) if fm.n_in else np.zeros(0, np.uint8)
        y = synth_tokens(77, k, 1, lo, max(fm.n_out, 1)) if fm.n_out else np.zeros(0, np.uint8)
        args = [str(len(x)), str(len(y))] + [str(int(t) - 1) for t in x] + [str(int(t) - 1) for t in y]
        r = subprocess.run([exe] + args, check=True, capture_output=True, text=True)
        got_f, got_v = [float(v) for v in r.stdout.split()]
        want_f = orc.forward(x, y, mode=LSE_EXACT)
        want_v, _ = orc.viterbi(x, y)
        if math.isinf(want_f):
            assert got_f == want_f, (name, k, got_f, want_f)
        else:
            assert abs(got_f - want_f) <= 1e-9 * max(1.0, abs(want_f)), (name, k, got_f, want_f)
        assert got_v == want_v, (name, k, got_v, want_v)
