"""What can be pinned about the aligner without parasail: the semi-global
alignment of oracle/sg_align.c -- the stand-in for parasail.sg_trace_scan_16 behind the reference's
block statistic (modules/cluster.py:130-169) -- is checked against an independent, differently
structured DP (three matrices swept by anti-diagonals in NumPy): the score is optimal, the CIGAR is
a valid path of exactly that score, and every tie-break variant is co-optimal. What cannot be pinned
is WHICH co-optimal path parasail reports; scripts/tiebreak_sensitivity.py measures how much the
clustering depends on that choice."""
import ctypes

import numpy as np
import pytest

from oracle import cluster_oracle as oc
from oracle import consensus_oracle as co

NEG = -10 ** 9


def semiglobal_score(s1, s2, open_pen, ext=1, match=2, mismatch=-2):
    """max over the last row / last column of the Gotoh matrices with free leading gaps."""
    a = np.frombuffer(s1.encode(), dtype=np.uint8)
    b = np.frombuffer(s2.encode(), dtype=np.uint8)
    n1, n2 = len(a), len(b)
    idx = np.arange(n1 + 1)
    h2 = np.full(n1 + 1, NEG); h1 = np.full(n1 + 1, NEG)        # diagonals d-2, d-1 (indexed by row i)
    e1 = np.full(n1 + 1, NEG); f1 = np.full(n1 + 1, NEG)
    h2[0] = 0                                                    # d = 0: cell (0, 0)
    h1[0] = 0; h1[1 if n1 >= 1 else 0] = 0                       # d = 1: cells (0, 1) and (1, 0)
    best = NEG
    last_col = np.full(n1 + 1, NEG)
    last_row = np.full(n2 + 1, NEG)
    for d in range(2, n1 + n2 + 1):
        lo, hi = max(0, d - n2), min(n1, d)                      # rows on this diagonal
        h0 = np.full(n1 + 1, NEG); e0 = np.full(n1 + 1, NEG); f0 = np.full(n1 + 1, NEG)
        i = idx[max(lo, 1):hi + 1]
        i = i[(d - i) >= 1]                                      # interior cells (i >= 1, j >= 1)
        if len(i):
            j = d - i
            e0[i] = np.maximum(h1[i] - open_pen, e1[i] - ext)            # from (i, j-1)
            f0[i] = np.maximum(h1[i - 1] - open_pen, f1[i - 1] - ext)    # from (i-1, j)
            sub = np.where(a[i - 1] == b[j - 1], match, mismatch)
            h0[i] = np.maximum(h2[i - 1] + sub, np.maximum(e0[i], f0[i]))
        if lo == 0:
            h0[0] = 0                                            # (0, d)
        if d <= n1:
            h0[d] = 0                                            # (d, 0)
        for ii in (i if len(i) else []):
            jj = d - ii
            if jj == n2:
                last_col[ii] = h0[ii]
            if ii == n1:
                last_row[jj] = h0[ii]
        h2, h1, e1, f1 = h1, h0, e0, f0
    best = max(int(last_col.max()), int(last_row.max()))
    return best


def rescore(ops, s1, s2, open_pen, ext=1, match=2, mismatch=-2):
    """Score of the path `ops` (=, X, I, D) with free end gaps; also checks that it spells both sequences."""
    assert ops.count("=") + ops.count("X") + ops.count("I") == len(s1)
    assert ops.count("=") + ops.count("X") + ops.count("D") == len(s2)
    lead = len(ops) - len(ops.lstrip(ops[0])) if ops[0] in "ID" else 0
    core = ops[lead:]
    trail = len(core) - len(core.rstrip(core[-1])) if core and core[-1] in "ID" else 0
    core = core[:len(core) - trail] if trail else core
    i = ops[:lead].count("I")
    j = ops[:lead].count("D")
    score, prev = 0, ""
    for op in core:
        if op in "=X":
            assert (s1[i] == s2[j]) == (op == "=")
            score += match if op == "=" else mismatch
            i += 1; j += 1
        else:
            score -= ext if op == prev else open_pen
            if op == "I":
                i += 1
            else:
                j += 1
        prev = op
    return score


def _pairs(rng, n):
    out = []
    for _ in range(n):
        L = int(rng.integers(20, 120))
        s = "".join(rng.choice(list("ACGT"), size=L))
        kind = rng.integers(0, 4)
        if kind == 0:
            t = "".join(rng.choice(list("ACGT"), size=int(rng.integers(20, 120))))
        else:
            t = list(s)
            for _e in range(int(rng.integers(0, 12))):
                p = int(rng.integers(0, max(1, len(t))))
                r = rng.random()
                if r < 0.4 and len(t) > 5:
                    del t[p]
                elif r < 0.7:
                    t.insert(p, "ACGT"[int(rng.integers(4))])
                else:
                    t[p] = "ACGT"[int(rng.integers(4))]
            t = "".join(t)
            if kind == 2:
                t = t[int(rng.integers(0, 10)):]                 # overhangs: free end gaps matter
            if kind == 3:
                t = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 15)))) + t
        out.append((s, t))
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_score_is_optimal_and_cigar_is_a_path_of_that_score(seed):
    rng = np.random.default_rng(seed)
    for s, t in _pairs(rng, 1000):
        o = int(rng.integers(2, 6))
        ops, score = co.align_ops(s, t, o)
        assert score == semiglobal_score(s, t, o), (s, t, o)
        assert rescore(ops, s, t, o) == score, (s, t, o, ops)


def test_tiebreak_variants_are_co_optimal():
    """Each alternative tie-break (H order, open-vs-extend ties, end-cell scan) yields a valid path of the
    same optimal score; only the path among co-optimal ones differs."""
    lib = oc._lib()
    lib.oracle_sg_set_tiebreak.argtypes = [ctypes.c_int]
    rng = np.random.default_rng(9)
    pairs = _pairs(rng, 400)
    differ = {v: 0 for v in (1, 2, 4, 8, 15)}
    try:
        for s, t in pairs:
            lib.oracle_sg_set_tiebreak(0)
            ops0, sc0 = co.align_ops(s, t, 3)
            for v in differ:
                lib.oracle_sg_set_tiebreak(v)
                ops, sc = co.align_ops(s, t, 3)
                assert sc == sc0
                assert rescore(ops, s, t, 3) == sc
                differ[v] += ops != ops0
    finally:
        lib.oracle_sg_set_tiebreak(0)
    assert sum(differ.values()) > 0          # the variants really take other paths sometimes
