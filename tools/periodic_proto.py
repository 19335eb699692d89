"""Prototype of the periodic-structure analysis for the column engine (mb_prof.cu)."""
import json, gzip, sys, collections
import numpy as np

def load(name):
    d = json.load(gzip.open('machineboss_b200/presets/%s.eval.json.gz' % name))
    return d['nStates'], d['trans']

def analyse(S, trans):
    out = collections.defaultdict(list)
    for tr in trans:
        out[tr[0]].append(tr)
    sig = [(len(out[s]), sum(1 for t in out[s] if t[3])) for s in range(S)]
    lo, hi = S // 4, 3 * S // 4
    P = None
    for p in range(1, 257):
        if all(sig[s] == sig[s + p] for s in range(lo, hi)):
            P = p; break
    if P is None:
        return None
    # maximal range of periodic signature
    a = lo
    while a - 1 >= 0 and sig[a - 1] == sig[a - 1 + P]: a -= 1
    b = hi
    while b + P < S and sig[b] == sig[b + P]: b += 1
    b += P   # [a, b) has periodic signature
    best = None
    for phase in range(P):
        a0 = a + phase
        K = (b - a0) // P
        # try extending K by allowing irregular periods at both ends (missing transitions)
        while a0 - P >= 1: a0 -= P; K += 1
        while a0 + (K + 1) * P <= S - 1: K += 1
        res = classify(S, trans, a0, P, K)
        # shrink from the ends while classification fails
        tries = 0
        a1, K1 = a0, K
        while res is None and K1 > 4 and tries < 8:
            # drop first and last period alternately
            if tries % 2 == 0: K1 -= 1
            else: a1 += P; K1 -= 1
            res = classify(S, trans, a1, P, K1); tries += 1
        if res is None: continue
        cand = (res['nGroups'] + 4 * (res['nCarried'] + res['nAcc']) + res['nPre'] + res['nSuf'], phase, a1, K1, res)
        if best is None or cand[0] < best[0]: best = cand
    return P, best

def classify(S, trans, a0, P, K):
    end = a0 + K * P
    def where(s):
        if s < a0: return ('pre', s)
        if s >= end: return ('suf', s - end)
        return ('per', (s - a0) // P, (s - a0) % P)
    groups = set(); carried = set(); acc = set()
    for (src, dst, i, o, w, *_rest) in trans:
        ws, wd = where(src), where(dst)
        emit = o != 0
        if ws[0] == 'per' and wd[0] == 'per':
            dk = wd[1] - ws[1]
            if dk not in (0, 1): return None
            if not emit and dk == 0 and wd[2] <= ws[2]: return None
            groups.add((ws[2], dk, wd[2], emit))
        elif ws[0] == 'pre' and wd[0] == 'per':
            carried.add(src); groups.add(('c', src, wd[2], emit))
        elif ws[0] == 'per' and wd[0] == 'suf':
            acc.add((dst, emit)); groups.add((ws[2], 'a', dst, emit))
        elif ws[0] == 'pre' and wd[0] == 'pre': pass
        elif ws[0] == 'suf' and wd[0] == 'suf': pass
        elif ws[0] == 'pre' and wd[0] == 'suf': pass
        else: return None
    return dict(nGroups=len(groups), nCarried=len(carried), nAcc=len(acc), nPre=a0, nSuf=S - end, groups=groups)

for name in sys.argv[1:]:
    S, trans = load(name)
    r = analyse(S, trans)
    if r is None or r[1] is None: print(name, 'no period', r); continue
    P, (score, phase, a0, K, res) = r
    print(name, 'S', S, 'T', len(trans), 'P', P, 'phase', phase, 'a0', a0, 'K', K, {k: v for k, v in res.items() if k != 'groups'})
    em = sum(1 for g in res['groups'] if g[3]); print('  emitting groups', em, 'silent', len(res['groups']) - em)
