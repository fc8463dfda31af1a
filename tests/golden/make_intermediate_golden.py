#!/usr/bin/env python
"""Golden vectors for the per-round snapshot files of the --t N path (O/<it>/pre_clusters.csv,
O/<it>/cluster_origins.csv): the REFERENCE's parallelize.print_intermediate_results (imported from
/root/reference with the parasail shim, build container only) on fixed inputs.
    python tests/golden/make_intermediate_golden.py  ->  tests/golden/intermediate.json.gz
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "parasail_shim"))
sys.path.insert(0, "/root/reference")
from modules import parallelize as ref  # noqa: E402

CASES = [
    {"clusters": {"7": ["r7 x=1_10.5", "r9_a_b_3.25", "r1_99.0"], "2": ["r2_50.0"], "11": ["r11 y_7.125", "r12_6.0"]},
     "reps": {"7": [7, 1, "r7 x=1_10.5", "ACGT", "IIII", 10.5, 0.0123456789, "ACGT"],
              "2": [2, 2, "r2_50.0", "GGCC", "!!!!", 50.0, 0.1, "GC"],
              "11": [11, 1, "r11 y_7.125", "TTTT", "5555", 7.125, 1e-05, "T"]}, "it": 1},
    {"clusters": {"0": ["a_1.0"]}, "reps": {"0": [0, 1, "a_1.0", "A", "I", 1.0, 0.25, "A"]}, "it": 3},
]


def main():
    out = []
    for c in CASES:
        clusters = {int(k): v for k, v in c["clusters"].items()}
        reps = {int(k): tuple(v) for k, v in c["reps"].items()}
        with tempfile.TemporaryDirectory() as d:
            ref.print_intermediate_results(clusters, reps, SimpleNamespace(outfolder=d), c["it"])
            files = {}
            for name in ("pre_clusters.csv", "cluster_origins.csv"):
                with open(os.path.join(d, str(c["it"]), name), "rb") as f:
                    files[name] = f.read().decode()
        out.append(dict(c, files=files))
    with gzip.open(os.path.join(HERE, "intermediate.json.gz"), "wt") as f:
        json.dump(out, f)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
