import gzip
import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        return json.load(f)


def read_fastq_gz(name):
    """Minimal 4-line FASTQ reader for the committed fixtures -> [(acc, seq, qual)]."""
    out = []
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        lines = f.read().split("\n")
    for i in range(0, len(lines) - 3, 4):
        out.append((lines[i][1:], lines[i + 1], lines[i + 3]))
    return out


def scenario_reads(tag):
    """Input records (acc, seq, qual) of a golden clustering scenario."""
    from ngspeciesid_b200.synth import simulate_reads
    if tag.startswith("h1"):
        return read_fastq_gz("sample_h1.fastq.gz")
    if tag.startswith("supp1k"):
        return read_fastq_gz("supp1_1000.fastq.gz")
    if tag.startswith("synth2k"):
        return list(simulate_reads(2000, n_species=5, seed=7).records())
    if tag.startswith("synthpb"):
        return list(simulate_reads(600, n_species=4, len_lo=1900, len_hi=2000, seed=11,
                                   profile="pacbio", per_read_len=(500, 2000)).records())
    raise KeyError(tag)


def scenario_args(golden):
    """Translate the stored reference CLI flags into the hot-path knobs."""
    from oracle.cluster_oracle import default_args
    a = golden["args"]
    kw = {}
    if "--isoseq" in a:
        kw.update(k=15, w=50)
    if "--symmetric_map_align_thresholds" in a:
        kw.update(symmetric_map_align_thresholds=True)
    kw["nr_cores"] = int(a[a.index("--t") + 1])
    return default_args(**kw)


@pytest.fixture(scope="session")
def p_table():
    import numpy as np
    z = np.load(os.path.join(ROOT, "ngspeciesid_b200", "data", "p_shared_table.npz"))
    return [(int(k), int(w), float(p), e1 / 100.0, e2 / 100.0)
            for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]
