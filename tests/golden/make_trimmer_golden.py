#!/usr/bin/env python
"""Golden vectors for the primer-trimming driver (SURVEY.md 8 f rank 4): the REFERENCE's
modules/barcode_trimmer.py (imported from /root/reference, build container only) driven by a stub
`edlib` whose align() is a brute-force statement of edlib's HW / task="locations" contract (minimum
edit distance over ALL substrings, every end that reaches it, earliest start per end). The vectors
pin the driver (read_barcodes, get_universal_tails, reverse_complement, find_barcode_locations,
remove_barcodes); the search itself stays unpinned against real edlib (absent).
    python tests/golden/make_trimmer_golden.py  ->  tests/golden/trimmer.json.gz
"""
import gzip
import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def brute_align(query, target, mode="HW", task="locations", k=-1, additionalEqualities=None):
    """Every substring target[s:e] against the query: one start-anchored Levenshtein table per
    start s (vectorised over s), so dist[s, j] = lev(query, target[s:s+j])."""
    assert mode == "HW" and task == "locations"
    pairs = set(additionalEqualities or [])
    pairs |= {(b, a) for a, b in pairs}
    m, n = len(query), len(target)
    if n == 0:
        return {"editDistance": -1, "locations": []}
    big = 10 ** 6
    eqm = np.array([[1 if (q == t or (q, t) in pairs) else 0 for t in target] for q in query], dtype=np.int64)   # m x n
    starts = np.arange(n)
    prev = np.tile(np.arange(n + 1, dtype=np.int64), (n, 1))          # row 0: j insertions
    for i in range(1, m + 1):
        cur = np.empty_like(prev)
        cur[:, 0] = i
        for j in range(1, n + 1):
            col = starts + j - 1                                         # target index of the j-th base
            ok = col < n
            cost = np.where(ok, 1 - eqm[i - 1, np.minimum(col, n - 1)], big)
            cur[:, j] = np.minimum(np.minimum(prev[:, j - 1] + cost, prev[:, j] + 1), cur[:, j - 1] + 1)
        prev = cur
    dist = prev.copy()
    for sidx in range(n):
        dist[sidx, n - sidx + 1:] = big                                 # beyond the end of the target
    dist[:, 0] = big                                                     # non-empty substrings only
    best = int(dist.min())
    if k >= 0 and best > k:
        return {"editDistance": -1, "locations": []}
    by_end = {}
    for sidx in range(n):                                                # s ascends: earliest start per end
        for j in np.nonzero(dist[sidx] == best)[0]:
            by_end.setdefault(sidx + int(j) - 1, sidx)
    return {"editDistance": best, "locations": [(by_end[e], e) for e in sorted(by_end)]}


stub = types.ModuleType("edlib")
stub.align = brute_align
sys.modules["edlib"] = stub
sys.path.insert(0, "/root/reference")
from modules import barcode_trimmer as ref  # noqa: E402

PRIMERS = ">P1-F some text\nACAAATCAYAARGAYATYGG\n>P2-R\nTTCAGGRTGNCCRAARAAYCA\n>short\nACGTTGCA\n"


def mutate(rng, s, n):
    s = list(s)
    for _ in range(n):
        p = int(rng.integers(len(s)))
        r = rng.random()
        if r < 0.34:
            s[p] = "ACGT"[int(rng.integers(4))]
        elif r < 0.67:
            del s[p]
        else:
            s.insert(p, "ACGT"[int(rng.integers(4))])
    return "".join(s)


def concrete(rng, primer):
    iupac = {"M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT",
             "B": "CGT", "N": "ACGT", "X": "ACGT"}
    return "".join(c if c in "ACGT" else iupac[c][int(rng.integers(len(iupac[c])))] for c in primer)


def main():
    rng = np.random.default_rng(12)
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(PRIMERS)
        pf = f.name
    barcodes = ref.read_barcodes(pf)
    os.unlink(pf)
    tails = ref.get_universal_tails()
    out = {"primers": PRIMERS, "barcodes": barcodes, "tails": tails,
           "revcomp": [[s, ref.reverse_complement(s)] for s in ["ACGT", "AcgTNnXYRKMSWBVHDyrkmswbvhd", ""]],
           "locations": [], "centers": []}
    names = list(barcodes)
    for case in range(36):
        L = int(rng.integers(20, 420))
        body = "".join(rng.choice(list("ACGT"), size=L))
        center = body
        if rng.random() < 0.8:
            b = barcodes[names[int(rng.integers(len(names)))]]
            center = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 30)))) + mutate(rng, concrete(rng, b), int(rng.integers(0, 4))) + center
        if rng.random() < 0.8:
            b = barcodes[names[int(rng.integers(len(names)))]]
            center = center + mutate(rng, concrete(rng, b), int(rng.integers(0, 4))) + "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 30))))
        for bc, tw, k in ((barcodes, 150, 2), (tails, 60, 3), (barcodes, 150, 0)):
            args = SimpleNamespace(trim_window=tw, primer_max_ed=k)
            centers = [[10, case, center, "path"]]
            upd = ref.remove_barcodes(centers, bc, args)
            out["centers"].append({"center": center, "which": "barcodes" if bc is barcodes else "tails",
                                   "trim_window": tw, "k": k, "updated": upd, "result": centers[0][2]})
        w = center[:150]
        out["locations"].append({"window": w, "k": 2, "hits": [list(x) for x in ref.find_barcode_locations(w, barcodes, 2)]})
    with gzip.open(os.path.join(HERE, "trimmer.json.gz"), "wt") as f:
        json.dump(out, f)
    print(len(out["centers"]), "center cases,", sum(1 for c in out["centers"] if c["updated"]), "trimmed")


if __name__ == "__main__":
    main()
