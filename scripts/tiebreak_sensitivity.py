"""How much do cluster assignments depend on WHICH co-optimal alignment the aligner reports?
(VERDICT r1 item 8 ii.) parasail 1.2.4 is not available, so its tie-breaks are unverified; this runs the
oracle clustering (pinned to the reference's golden vectors) on the golden scenario inputs under every
alternative tie-break of oracle/sg_align.c and counts the decisions that change.
    python scripts/tiebreak_sensitivity.py            -> table for DESIGN.md section 2"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from conftest import scenario_reads
from oracle import cluster_oracle as oc

VARIANTS = [(0, "documented choice (H: diag > D > I; ties extend; last column, then last row)"),
            (1, "H prefers I over D"), (2, "gap ties open instead of extend"),
            (4, "end cell: last row before last column"), (8, "H prefers gaps over the diagonal"),
            (15, "all four together")]


def main():
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    z = np.load(os.path.join(root, "ngspeciesid_b200", "data", "p_shared_table.npz"))
    p_table = [(int(k), int(w), float(p), e1 / 100.0, e2 / 100.0) for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]
    lib = oc._lib()
    lib.oracle_sg_set_tiebreak.argtypes = [ctypes.c_int]
    print("| scenario | reads | decided by alignment | " + " | ".join("variant %d" % v for v, _ in VARIANTS[1:]) + " |")
    print("|---|---|---|" + "---|" * (len(VARIANTS) - 1))
    for tag in ("h1", "supp1k", "synth2k"):
        args = oc.default_args()
        ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads(tag), args.k))
        p_emp = oc.load_p_emp(p_table, args.k, args.w)
        base = None
        cells = []
        for v, _name in VARIANTS:
            lib.oracle_sg_set_tiebreak(v)
            st = oc.Stats()
            oc.single_clustering(list(ra), p_emp, args, st)
            dec = [(w, h) for _r, w, h in st.trace]
            if base is None:
                base = dec
                n_aln = sum(1 for _w, h in dec if h == "align") + sum(1 for w, h in dec if h == "new")
                n_aln = st.aln_called
            else:
                cells.append("%d changed" % sum(1 for a, b in zip(base, dec) if a != b))
        lib.oracle_sg_set_tiebreak(0)
        print("| %s | %d | %d | %s |" % (tag, len(ra), n_aln, " | ".join(cells)))
    for v, name in VARIANTS:
        print("variant %d: %s" % (v, name))


if __name__ == "__main__":
    main()
