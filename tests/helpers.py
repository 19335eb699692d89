"""Shared test helpers: flat machines, the numpy synthetic generator, ctypes bindings of the C oracle.

Only tests (and smoke / the CPU-baseline legs of bench.py) may touch ``oracle/``.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
from dataclasses import dataclass

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmb_oracle.so")
REFDRV = os.path.join(ORACLE_DIR, "_ref", "refdrv")

NEG_INF = float("-inf")


def _num(v):
    if isinstance(v, str):
        return {"-Infinity": NEG_INF, "Infinity": float("inf"), "NaN": float("nan")}[v]
    return float(v)


@dataclass
class FlatMachine:
    """Flat evaluated machine in the reference's enumeration order (src/eval.cpp:49-69)."""

    n_states: int
    n_in: int
    n_out: int
    src: np.ndarray
    dst: np.ndarray
    tin: np.ndarray
    tout: np.ndarray
    lw: np.ndarray
    in_alphabet: list
    out_alphabet: list

    @property
    def n_trans(self) -> int:
        return int(self.src.shape[0])

    @staticmethod
    def from_json(j: dict) -> "FlatMachine":
        t = j["trans"]
        return FlatMachine(
            n_states=int(j["nStates"]),
            n_in=len(j["inAlphabet"]),
            n_out=len(j["outAlphabet"]),
            src=np.array([r[0] for r in t], dtype=np.int32),
            dst=np.array([r[1] for r in t], dtype=np.int32),
            tin=np.array([r[2] for r in t], dtype=np.int32),
            tout=np.array([r[3] for r in t], dtype=np.int32),
            lw=np.array([_num(r[4]) for r in t], dtype=np.float64),
            in_alphabet=list(j["inAlphabet"]),
            out_alphabet=list(j["outAlphabet"]),
        )

    def with_weights(self, lw: np.ndarray) -> "FlatMachine":
        return FlatMachine(self.n_states, self.n_in, self.n_out, self.src, self.dst, self.tin, self.tout,
                           np.ascontiguousarray(lw, dtype=np.float64), self.in_alphabet, self.out_alphabet)


def synthetic_profile(n_nodes: int = 80, n_out: int = 4, seed: int = 7, skip: bool = False) -> FlatMachine:
    """A generator with a period of 3 states per node (M, I, D) whose transitions consume the token when they LEAVE a state --
    unlike the reference's HMMER import (Mx / M, Ix / I), so that the column engine meets diagonal groups (node k -> k+1
    consuming a token), token-consuming self-loops, a begin hub entered with a token, and flanking states with their own
    dynamics (N before, C after the nodes).  skip: add node k -> k+2 transitions, which no column sweep can take."""
    rng = np.random.default_rng(seed)
    S0, N, B = 0, 1, 2
    M = lambda k: 3 + 3 * (k - 1)
    I = lambda k: M(k) + 1
    D = lambda k: M(k) + 2
    E = 3 + 3 * n_nodes
    C, T = E + 1, E + 2
    rows = []

    def silent(a, b):
        rows.append((a, b, 0, 0, float(np.log(rng.uniform(0.05, 0.6)))))

    def emit(a, b):
        for c in range(1, n_out + 1):
            rows.append((a, b, 0, c, float(np.log(rng.uniform(0.02, 0.5)))))

    silent(S0, N); silent(S0, B); emit(N, N); silent(N, B)
    for k in range(1, n_nodes + 1):
        emit(B, M(k))
    for k in range(1, n_nodes + 1):
        emit(M(k), I(k)); emit(I(k), I(k)); silent(M(k), E)
        if k < n_nodes:
            emit(M(k), M(k + 1)); emit(I(k), M(k + 1)); silent(M(k), D(k + 1)); silent(D(k), D(k + 1)); emit(D(k), M(k + 1))
        else:
            silent(D(k), E); silent(I(k), E)
        if skip and k + 2 <= n_nodes:
            silent(M(k), D(k + 2))
    silent(E, C); emit(C, C); silent(C, T); silent(E, T)
    rows.sort(key=lambda r: r[0])      # the reference enumerates transitions state by state (eval.cpp:49-69)
    a = np.array(rows, dtype=object)
    return FlatMachine(T + 1, 0, n_out, np.array([r[0] for r in rows], np.int32), np.array([r[1] for r in rows], np.int32),
                       np.zeros(len(rows), np.int32), np.array([r[3] for r in rows], np.int32), np.array([r[4] for r in rows], np.float64),
                       [], [chr(65 + c) for c in range(n_out)])


def load_golden(name: str) -> dict:
    path = os.path.join(GOLDEN, name + ".json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
    else:
        import gzip
        with gzip.open(path + ".gz", "rt") as f:
            j = json.load(f)
    if isinstance(j, dict) and "machine_from" in j:      # large machines are stored once
        j["machine"] = load_golden(j["machine_from"])["machine"]
    return j


def golden_names() -> list:
    # aux_* files hold auxiliary inputs (not DP cases)
    return sorted(f[:-5] if f.endswith(".json") else f[:-8] for f in os.listdir(GOLDEN)
                  if f.endswith((".json", ".json.gz")) and not f.startswith("aux_"))


# ---------------------------------------------------------------------------------------------
# synthetic tokens: numpy restatement of oracle/synth.h (kept in the tests AND in bench.py)
# ---------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth_tokens(seed: int, pair_index: int, which: int, length: int, n_sym: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
                + np.uint64(2 * pair_index + which) * np.uint64(0xD1B54A32D192ED03))
        p = np.arange(length, dtype=np.uint64) + base
    return (1 + (_splitmix64(p) % np.uint64(n_sym))).astype(np.uint8)


def mutate_tokens(seed: int, pair_index: int, x: np.ndarray, n_sym: int, sub: float = 0.10, indel: float = 0.02) -> np.ndarray:
    """SURVEY 8(d) parameter set B's data: the output is the input with `sub` substitutions and `indel` insertions
    + deletions per position (half each), drawn from the splitmix64 stream (pair, which = 2)."""
    n = len(x)
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
                + np.uint64(2 * pair_index) * np.uint64(0xD1B54A32D192ED03) + np.uint64(0x5851F42D4C957F2D))
        r = _splitmix64(np.arange(3 * n, dtype=np.uint64) + base)
    u = (r[:n] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    alt = (r[n:2 * n] % np.uint64(n_sym - 1)).astype(np.int64)
    ins = (1 + r[2 * n:] % np.uint64(n_sym)).astype(np.uint8)
    xi = x.astype(np.int64)
    subst = (1 + (xi - 1 + 1 + alt) % n_sym).astype(np.uint8)      # a different symbol
    out = []
    for p in range(n):
        if u[p] < indel / 2:
            continue                                 # deletion
        out.append(subst[p] if u[p] < indel / 2 + sub else x[p])
        if u[p] > 1.0 - indel / 2:
            out.append(ins[p])                       # insertion after this position
    return np.array(out, dtype=np.uint8)


# ---------------------------------------------------------------------------------------------
# C oracle (oracle/libmb_oracle.so) over ctypes
# ---------------------------------------------------------------------------------------------
class _MboMachine(ctypes.Structure):
    _fields_ = [("nStates", ctypes.c_int32), ("nInTok", ctypes.c_int32), ("nOutTok", ctypes.c_int32),
                ("nTrans", ctypes.c_int64),
                ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("tin", ctypes.c_void_p),
                ("tout", ctypes.c_void_p), ("lw", ctypes.c_void_p)]


_oracle = None


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = ctypes.CDLL(ORACLE_SO)
        P, D, I64, U8 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_void_p
        lib.mbo_forward.restype = D
        lib.mbo_forward.argtypes = [P, U8, I64, U8, I64, ctypes.c_int, P]
        lib.mbo_backward.restype = D
        lib.mbo_backward.argtypes = [P, U8, I64, U8, I64, ctypes.c_int, P]
        lib.mbo_viterbi.restype = D
        lib.mbo_viterbi.argtypes = [P, U8, I64, U8, I64, P, P, I64, P]
        lib.mbo_counts.restype = D
        lib.mbo_counts.argtypes = [P, U8, I64, U8, I64, ctypes.c_int, P, P]
        lib.mbo_synth.restype = None
        lib.mbo_synth.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, I64, ctypes.c_int, P]
        lib.mbo_set_envelope.restype = None
        lib.mbo_set_envelope.argtypes = [P, P]
        lib.mbo_log_sum_exp.restype = D
        lib.mbo_log_sum_exp.argtypes = [D, D, ctypes.c_int]
        _oracle = lib
    return _oracle


LSE_TABLE, LSE_EXACT = 0, 1


class Oracle:
    """The C restatement bound to one flat machine."""

    def __init__(self, m: FlatMachine):
        self.m = m
        self.lib = oracle_lib()
        self._keep = [np.ascontiguousarray(a) for a in (m.src, m.dst, m.tin, m.tout, m.lw)]
        self.c = _MboMachine(m.n_states, m.n_in, m.n_out, m.n_trans, *[a.ctypes.data for a in self._keep])

    def set_envelope(self, env=None):
        """env: list of [inStart, inEnd) per output row (Lo+1 rows), or None for the full matrix."""
        if env is None:
            self._env = None
            self.lib.mbo_set_envelope(None, None)
        else:
            a = np.ascontiguousarray(np.array(env, dtype=np.int64))
            self._env = (np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]))
            self.lib.mbo_set_envelope(self._env[0].ctypes.data, self._env[1].ctypes.data)

    @staticmethod
    def _tok(a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return a, a.ctypes.data if a.size else None

    def forward(self, x, y, mode=LSE_TABLE, matrix=False):
        x, xp = self._tok(x)
        y, yp = self._tok(y)
        mat = np.empty((len(y) + 1, len(x) + 1, self.m.n_states)) if matrix else None
        ll = self.lib.mbo_forward(ctypes.byref(self.c), xp, len(x), yp, len(y), mode, mat.ctypes.data if matrix else None)
        return (ll, mat) if matrix else ll

    def backward(self, x, y, mode=LSE_TABLE, matrix=False):
        x, xp = self._tok(x)
        y, yp = self._tok(y)
        mat = np.empty((len(y) + 1, len(x) + 1, self.m.n_states)) if matrix else None
        ll = self.lib.mbo_backward(ctypes.byref(self.c), xp, len(x), yp, len(y), mode, mat.ctypes.data if matrix else None)
        return (ll, mat) if matrix else ll

    def viterbi(self, x, y, path=True, matrix=False):
        x, xp = self._tok(x)
        y, yp = self._tok(y)
        mat = np.empty((len(y) + 1, len(x) + 1, self.m.n_states)) if matrix else None
        cap = (len(x) + len(y) + 1) * max(1, self.m.n_states) + 1
        buf = np.empty(cap, dtype=np.int32) if path else None
        n = ctypes.c_int64(0)
        sc = self.lib.mbo_viterbi(ctypes.byref(self.c), xp, len(x), yp, len(y), mat.ctypes.data if matrix else None,
                                  buf.ctypes.data if path else None, cap, ctypes.byref(n))
        out = [sc]
        if path:
            out.append(buf[: n.value].copy())
        if matrix:
            out.append(mat)
        return out[0] if len(out) == 1 else tuple(out)

    def counts(self, x, y, mode=LSE_TABLE, counts=None):
        x, xp = self._tok(x)
        y, yp = self._tok(y)
        if counts is None:
            counts = np.zeros(self.m.n_trans)
        bll = ctypes.c_double(0)
        fll = self.lib.mbo_counts(ctypes.byref(self.c), xp, len(x), yp, len(y), mode, counts.ctypes.data, ctypes.byref(bll))
        return fll, bll.value, counts


def pairs_from_golden(case: dict):
    return [(np.array(p["x"], dtype=np.uint8), np.array(p["y"], dtype=np.uint8)) for p in case["pairs"]]


def gnum(v):
    return _num(v)
