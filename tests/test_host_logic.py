"""Host-side helpers of the product path (no GPU): bucket thresholds, gap table, ranks, batching."""
import math

import numpy as np

from ngspeciesid_b200 import engine
from ngspeciesid_b200.modules import p_minimizers_shared, parallelize
from oracle import cluster_oracle as oc


def test_bucket_thresholds_match_round():
    thr = engine.bucket_thresholds()
    vals = engine.bucket_values()
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(0, 0.2, 20000)) + [t for t in thr] + [math.nextafter(t, 0) for t in thr] + [0.125, 0.005, 0.0, 0.5]
    for x in xs:
        x = float(x)        # np.float64.__round__ is not Python's correctly-rounded round()
        b = int(sum(1 for t in thr if x >= t))
        assert vals[b] == oc.error_bucket(x), x


def test_max_gap_table_matches_sequential_product():
    p_emp = p_minimizers_shared.p_emp_for(13, 20)
    assert len(p_emp) == 225
    mg = engine.max_gap_table(p_emp, 0.1)
    vals = engine.bucket_values()
    for b1 in (0, 5, 14):
        for b2 in (0, 7, 14):
            q = 1.0 - p_emp[(vals[b1], vals[b2])]
            g = mg[b1 * 15 + b2]
            for gap in range(0, 40):
                prod = 1
                for _ in range(gap):
                    prod = prod * q
                assert (not (prod < 0.1)) == (gap <= g)


def test_table_equals_reference_rows(p_table):
    rows = p_minimizers_shared.read_empirical_p()
    assert len(rows) == 41880
    assert rows[0][:2] == (10, 10) and abs(rows[0][2] - 0.15071949855943487) == 0.0
    assert oc.load_p_emp(p_table, 15, 50) == p_minimizers_shared.p_emp_for(15, 50)


def test_accession_ranks():
    accs = ["b_1.5", "a_2", "c", "a_2", "B", "a_10"]
    r = engine.accession_ranks(accs)
    for i in range(len(accs)):
        for j in range(len(accs)):
            assert (accs[i] < accs[j]) == (r[i] < r[j])


def test_decode_kmer():
    assert engine.decode_kmer(0b00011011, 4) == "ACGT"
    assert engine.decode_kmer((1 << 31) | (1 << 4) | 0b0110, 13) == "CG"
    assert engine.decode_kmer((1 << 31) | 1, 13) == ""


def test_batch_list_matches_oracle():
    rng = np.random.default_rng(1)
    reads = [(i, 0, "r%d_1.0" % i, "A" * int(rng.integers(50, 500)), "", 1.0) for i in range(300)]
    for t in (2, 3, 8):
        assert list(parallelize.batch_list(reads, t, batch_type="total_nt")) == oc.split_batches(reads, t, "total_nt")
        assert list(parallelize.batch_list(reads, t, batch_type="nr_reads")) == oc.split_batches(reads, t, "nr_reads")
    tagged = [(r[0], 1 + (i * 4) // len(reads), r[2], r[3], r[4], r[5]) for i, r in enumerate(reads)]
    assert list(parallelize.batch_list(tagged, 4, merge_consecutive=True)) == oc.pair_batches(tagged)


def test_round_snapshot_files_match_reference(tmp_path):
    """parallelize.print_intermediate_results against the files the reference's own function wrote
    for the same inputs (tests/golden/make_intermediate_golden.py)."""
    import os
    from types import SimpleNamespace
    from conftest import load_golden
    for c in load_golden("intermediate.json.gz"):
        clusters = {int(k): v for k, v in c["clusters"].items()}
        reps = {int(k): tuple(v) for k, v in c["reps"].items()}
        parallelize.print_intermediate_results(clusters, reps, SimpleNamespace(outfolder=str(tmp_path)), c["it"])
        for name, text in c["files"].items():
            assert open(os.path.join(str(tmp_path), str(c["it"]), name)).read() == text
