"""Primer trimming (SURVEY.md 8 f rank 4), CPU only. The driver is pinned by golden vectors made
with the REFERENCE's barcode_trimmer.py driven by a brute-force edlib stub
(tests/golden/make_trimmer_golden.py); the infix search itself is checked against edlib's
documented HW / locations contract by brute force over all substrings (real edlib is absent:
parity of the search is unpinned)."""
import os
from types import SimpleNamespace

import numpy as np

from conftest import load_golden
from ngspeciesid_b200.modules import barcode_trimmer as bt

G = load_golden("trimmer.json.gz")


def test_barcode_tables_and_reverse_complement(tmp_path):
    p = os.path.join(str(tmp_path), "primers.fa")
    with open(p, "w") as f:
        f.write(G["primers"])
    assert bt.read_barcodes(p) == G["barcodes"]
    assert list(bt.read_barcodes(p)) == list(G["barcodes"])          # same insertion order (first hit wins ties)
    assert bt.get_universal_tails() == G["tails"]
    for s, rc in G["revcomp"]:
        assert bt.reverse_complement(s) == rc


def test_find_barcode_locations_match_reference_driver():
    for c in G["locations"]:
        got = [list(x) for x in bt.find_barcode_locations(c["window"], G["barcodes"], c["k"])]
        assert got == c["hits"]


def test_remove_barcodes_matches_reference_driver():
    n_trimmed = 0
    for c in G["centers"]:
        bc = G["barcodes"] if c["which"] == "barcodes" else G["tails"]
        centers = [[10, 0, c["center"], "path"]]
        upd = bt.remove_barcodes(centers, bc, SimpleNamespace(trim_window=c["trim_window"], primer_max_ed=c["k"]))
        assert upd == c["updated"] and centers[0][2] == c["result"]
        n_trimmed += upd
    assert n_trimmed > 20


def _lev(a, b):
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j - 1] + (0 if bt._same(a[i - 1], b[j - 1]) else 1), prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[len(b)]


def test_infix_search_contract_by_brute_force():
    """Minimum over all non-empty substrings, all ends that reach it (ascending), earliest start
    per end, nothing above k -- on short random strings with IUPAC codes in the query."""
    rng = np.random.default_rng(5)
    for _ in range(150):
        q = "".join(rng.choice(list("ACGTRYN"), size=int(rng.integers(1, 7))))
        t = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 14))))
        best, by_end = None, {}
        for s in range(len(t)):
            for e in range(s + 1, len(t) + 1):
                d = _lev(q, t[s:e])
                if best is None or d < best:
                    best, by_end = d, {}
                if d == best:
                    by_end.setdefault(e - 1, s)
        for k in (-1, 0, 1, 2):
            ed, locs = bt.find_locations(q, t, k)
            if k >= 0 and best > k:
                assert (ed, locs) == (-1, [])
            else:
                assert ed == best and locs == [(by_end[e], e) for e in sorted(by_end)], (q, t, k)
