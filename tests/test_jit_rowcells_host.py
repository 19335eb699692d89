"""CPU test of the score module's generated cell functions (mb_jit.cu: gen_cell_row_vit, gen_cell_row_lin) WITH REAL
WEIGHTS: the cells are compiled for the host and driven by a plain row-major loop over the tables the library
prepares for the device (mb_jit_host_tables: emission weights in row layout, the normalised linear weights and their
scales), then compared with the oracle.  What it pins without a GPU:
  * the row layout of the emission tables (one add per cell instead of an index computation per weight);
  * the linear-domain normalisation (a state's first silent group becomes a plain copy; every other weight is
    scaled by sigma_src / sigma_self; the result carries log sigma of the end state): Forward and Backward values
    must be the exact sums;
  * the Viterbi cell, score-only and with the pointer fields recorded: bit-identical scores.
The skewed sweep around the cells is the job of the -m gpu tests."""
import os
import struct
import subprocess

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
#define MB_HOST_HARNESS 1
static const char* hostE = nullptr;      // the table the cell reads: addresses are byte offsets into it
#define MB_LDS(addr, off) (*(const double*) (hostE + (addr) + (off)))
static inline double __longlong_as_double (long long v) { double d; memcpy (&d, &v, 8); return d; }
static inline long long __double_as_longlong (double d) { long long v; memcpy (&v, &d, 8); return v; }
template<bool PTR> static inline void mb_vmax (double& n, const double t, unsigned& word, const unsigned field, const unsigned val) {
  if (n < t) { n = t; if (PTR) word = (word & ~field) | val; }
}
static inline double mb_neg_inf() { return -INFINITY; }
%(cells)s

static std::vector<double> readDoubles (FILE* f) { long long n; if (fread (&n, 8, 1, f) != 1) exit (2); std::vector<double> v ((size_t) n); if (n && fread (v.data(), 8, (size_t) n, f) != (size_t) n) exit (2); return v; }

// one full matrix, row-major, with the cell function `cell`; reversed = the Backward sweep's coordinates
template<class Cell>
static void sweep (const std::vector<int>& x, const std::vector<int>& y, bool reversed, double ZERO, int WA, int WB, int originState, Cell cell, double* out) {
  const int Li = (int) x.size(), Lo = (int) y.size();
  std::vector<std::vector<double> > prev (Li + 1, std::vector<double> (MB_S, ZERO)), cur = prev;
  for (int o = 0; o <= Lo; ++o) {
    for (int i = 0; i <= Li; ++i) {
      double D[MB_S], L[MB_S], U[MB_S], N[MB_S];
      for (int s = 0; s < MB_S; ++s) { D[s] = (i && o) ? prev[i - 1][s] : ZERO; L[s] = i ? cur[i - 1][s] : ZERO; U[s] = o ? prev[i][s] : ZERO; }
      const int a = i ? (reversed ? x[Li - i] : x[i - 1]) : 0, b = o ? (reversed ? y[Lo - o] : y[o - 1]) : 0;
      const unsigned ea = (unsigned) (a * WA * 8), em = ea + (unsigned) (b * 8), ebr = (unsigned) ((MB_NIN * WA + b * WB) * 8);
      cell (D, L, U, N, ea, em, ebr, i == 0 && o == 0, i == Li && o == Lo);
      for (int s = 0; s < MB_S; ++s) cur[i][s] = N[s];
    }
    prev = cur;
  }
  for (int s = 0; s < MB_S; ++s) out[s] = cur[Li][s];
}

int main (int argc, char** argv) {
  FILE* f = fopen (argv[1], "rb");
  if (!f) return 2;
  const std::vector<double> rowLog = readDoubles (f), rowF = readDoubles (f), rowB = readDoubles (f), silLog = readDoubles (f), silN = readDoubles (f), xs = readDoubles (f), ys = readDoubles (f);
  fclose (f);
  std::vector<int> x, y;
  for (double v: xs) x.push_back ((int) v - 1);
  for (double v: ys) y.push_back ((int) v - 1);
  MBSil P;
  memcpy (&P, silLog.data(), sizeof P);
  MBSilN PN;
  memcpy (&PN, silN.data(), sizeof PN);
  double out[MB_S];
  // Viterbi: score only, and with the pointer fields recorded at the first and at the last column's position of the packed words
  double vit[3];
  int q = 0;
  hostE = (const char*) rowLog.data();
  sweep (x, y, false, -INFINITY, MB_WA_F, MB_WB_F, 0, [&] (const double (&D)[MB_S], const double (&L)[MB_S], const double (&U)[MB_S], double (&N)[MB_S], unsigned ea, unsigned em, unsigned ebr, bool origin, bool sink) {
    unsigned pk[MB_PKW] = { 0 }; mb_cell_vitr<false> (D, L, U, N, ea, em, ebr, origin, sink, P, pk, 0); }, out);
  vit[q++] = out[MB_S - 1];
  sweep (x, y, false, -INFINITY, MB_WA_F, MB_WB_F, 0, [&] (const double (&D)[MB_S], const double (&L)[MB_S], const double (&U)[MB_S], double (&N)[MB_S], unsigned ea, unsigned em, unsigned ebr, bool origin, bool sink) {
    unsigned pk[MB_PKW] = { 0 }; mb_cell_vitr<true> (D, L, U, N, ea, em, ebr, origin, sink, P, pk, 0); }, out);
  vit[q++] = out[MB_S - 1];
  sweep (x, y, false, -INFINITY, MB_WA_F, MB_WB_F, 0, [&] (const double (&D)[MB_S], const double (&L)[MB_S], const double (&U)[MB_S], double (&N)[MB_S], unsigned ea, unsigned em, unsigned ebr, bool origin, bool sink) {
    unsigned pk[MB_PKW] = { 0 }; mb_cell_vitr<true> (D, L, U, N, ea, em, ebr, origin, sink, P, pk, 8 * MB_TBBYTES * (MB_C - 1)); }, out);
  vit[q++] = out[MB_S - 1];
  // Forward and Backward, normalised linear domain
  hostE = (const char*) rowF.data();
  sweep (x, y, false, 0.0, MB_WA_F, MB_WB_F, 0, [&] (const double (&D)[MB_S], const double (&L)[MB_S], const double (&U)[MB_S], double (&N)[MB_S], unsigned ea, unsigned em, unsigned ebr, bool origin, bool) {
    mb_cell_fwd_linr (D, L, U, N, ea, em, ebr, origin, PN); }, out);
  const double fwd = out[MB_S - 1] > 0 ? log (out[MB_S - 1]) + PN.resLogF : -INFINITY;
  hostE = (const char*) rowB.data();
  sweep (x, y, true, 0.0, MB_WA_B, MB_WB_B, MB_S - 1, [&] (const double (&D)[MB_S], const double (&L)[MB_S], const double (&U)[MB_S], double (&N)[MB_S], unsigned ea, unsigned em, unsigned ebr, bool origin, bool) {
    mb_cell_bwd_linr (D, L, U, N, ea, em, ebr, origin, PN); }, out);
  const double bwd = out[0] > 0 ? log (out[0]) + PN.resLogB : -INFINITY;
  printf ("%%.17g %%.17g %%.17g %%.17g %%.17g\n", vit[0], vit[1], vit[2], fwd, bwd);
  return 0;
}
"""


def _write_doubles(f, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    f.write(struct.pack("<q", a.size))
    f.write(a.tobytes())


@pytest.mark.parametrize("name,shapes", [("dnapsw_peaked", [(9, 11), (0, 4), (5, 0), (0, 0), (30, 26)]), ("protpsw_synth", [(7, 9), (12, 3)]),
                                         ("unitindel", [(3, 4), (6, 2)]), ("bitnoise_tiny", [(3, 3), (5, 5)]), ("stutter_noise_difflen", [(2, 5), (4, 7)])])
def test_generated_row_cells_on_the_host(name, shapes, monkeypatch, tmp_path):
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    dump = str(tmp_path / "gen")
    monkeypatch.setenv("MB_JIT_DUMP", dump)
    capi.jit_compile_check(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    src = open(dump + ".viterbi.cu").read()
    head = src[: src.index("// MB_ROWCELLS_BEGIN")]
    keep = [l for l in head.splitlines() if l.startswith("#define MB_") or l.startswith("struct MBSil ") or (l.startswith("typedef ") and "mb_tbword" in l)]
    cells = "\n".join(keep) + "\n" + src[src.index("// MB_ROWCELLS_BEGIN"): src.index("// MB_ROWCELLS_END")]
    cpp = tmp_path / "harness.cpp"
    cpp.write_text(HARNESS % {"cells": cells})
    exe = str(tmp_path / "harness")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-o", exe, str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout, fm.lw)
    tables = [capi.jit_host_tables(*args, which=w) for w in range(5)]
    flags = capi.jit_host_tables(*args, which=5)
    assert flags[0] == 1.0, "the normalised kernels should be usable for %s" % name
    orc = Oracle(fm)
    for k, (li, lo) in enumerate(shapes):
        x = synth_tokens(61, k, 0, li, max(fm.n_in, 1)) if fm.n_in else np.zeros(0, np.uint8)
        y = synth_tokens(61, k, 1, lo, max(fm.n_out, 1)) if fm.n_out else np.zeros(0, np.uint8)
        data = tmp_path / ("case%d.bin" % k)
        with open(data, "wb") as f:
            for t in tables:
                _write_doubles(f, t)
            _write_doubles(f, x)
            _write_doubles(f, y)
        out = subprocess.run([exe, str(data)], check=True, capture_output=True, text=True).stdout.split()
        v0, v2, v4, fwd, bwd = [float(v) for v in out]
        want_v, _ = orc.viterbi(x, y)
        assert v0 == want_v and v2 == want_v and v4 == want_v, (name, k, v0, v2, v4, want_v)                      # bit-exact
        want_f = orc.forward(x, y, mode=LSE_EXACT)
        for got in (fwd, bwd):
            if np.isinf(want_f):
                assert got == want_f, (name, k, got, want_f)
            else:
                assert abs(got - want_f) <= 1e-11 * max(1.0, abs(want_f)), (name, k, got, want_f)


def test_normalisation_is_refused_when_a_unit_weight_is_zero():
    """A silent transition of weight 0 that the generated code treats as a state's unit group cannot be divided out:
    the host must report the normalised kernels unusable (the library then runs the first module's plain ones)."""
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden("dnapsw_peaked")["machine"])
    silent = np.where((fm.tin == 0) & (fm.tout == 0))[0]
    lw = fm.lw.copy()
    lw[silent[0]] = -np.inf
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    assert capi.jit_host_tables(*args, fm.lw, which=5)[0] == 1.0
    flags = [capi.jit_host_tables(*args, np.where(np.arange(len(lw)) == t, -np.inf, fm.lw), which=5)[0] for t in silent]
    assert 0.0 in flags      # at least one of the silent groups is a unit group
