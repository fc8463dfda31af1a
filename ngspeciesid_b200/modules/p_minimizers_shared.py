"""Empirical probabilities that a minimizer is shared between two reads of given error rates.
Same accessor as the reference (modules/p_minimizers_shared.py:2); the 41 880 rows live in
ngspeciesid_b200/data/p_shared_table.npz (data extracted from the reference's table by
tests/golden/make_golden.py) instead of a 1.8 MB Python literal."""
import os

import numpy as np

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "p_shared_table.npz")


def read_empirical_p():
    """-> list of (k, w, p_shared, e1, e2) tuples, in the reference's row order."""
    z = np.load(_TABLE)
    return [(int(k), int(w), float(p), round(e1 / 100.0, 2), round(e2 / 100.0, 2))
            for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]


def p_emp_for(k, w):
    """The 225-key dict NGSpeciesID:72-77 builds for (k, w)."""
    out = {}
    for kk, ww, p, e1, e2 in read_empirical_p():
        if kk == k and abs(ww - w) <= 2:
            out[(float(e1), float(e2))] = float(p)
            out[(float(e2), float(e1))] = float(p)
    return out
