"""Host-side helpers of the product path (no GPU): bucket thresholds, gap table, ranks, batching."""
import math

import pytest

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


def test_hostpack_pack_fields_matches_join():
    """csrc/hostpack.c: tuples of str -> the arrays ngsid_upload_reads takes, the same bytes as join + encode."""
    import random
    from ngspeciesid_b200.build import load_hostpack
    _hostpack = load_hostpack()
    rng = random.Random(7)
    recs = []
    for i in range(500):
        n = rng.randrange(0, 40)
        s = "".join(rng.choice("ACGTN") for _ in range(n))
        rec = (i, 0, "acc_%d" % i, s, "".join(chr(33 + rng.randrange(60)) for _ in range(n)), 1.0)
        recs.append(rec if i % 3 else list(rec) + [0.1, "x"])          # tuples and lists, 6 and 8 fields
    a, q, o = _hostpack.pack_fields(recs, 3, 4)
    off = np.frombuffer(o, dtype=np.int64)
    assert a == "".join(r[3] for r in recs).encode() and q == "".join(r[4] for r in recs).encode()
    assert off[0] == 0 and off[-1] == len(a) and list(np.diff(off)) == [len(r[3]) for r in recs]
    assert _hostpack.pack_fields([], 3, 4)[0] == b""
    assert _hostpack.measure_fields(recs, 3) == len(a)
    da, dq = np.zeros(len(a) + 8, dtype=np.uint8), np.zeros(len(a) + 8, dtype=np.uint8)
    o2 = _hostpack.pack_fields_into(recs, 3, 4, da.ctypes.data, dq.ctypes.data, len(a))
    assert o2 == o and da[:len(a)].tobytes() == a and dq[:len(a)].tobytes() == q and not da[len(a):].any()
    with pytest.raises(BufferError):
        _hostpack.pack_fields_into(recs, 3, 4, da.ctypes.data, dq.ctypes.data, len(a) - 1)
    for bad, exc in (([(0, "é", "I")], ValueError), ([(0, "AC", "I")], ValueError), ([(0, b"AC", "II")], TypeError),
                     ([(0, "AC")], IndexError), ([5], TypeError)):
        with pytest.raises(exc):
            _hostpack.pack_fields(bad, 1, 2)


def test_accession_ranks_ascii_and_unicode_agree_with_python_order():
    from ngspeciesid_b200 import engine as E
    accs = ["read_10_3.5", "read_9_3.5", "a", "read_10_3.5", "Z", "é_1", "~"]
    r = E.accession_ranks(accs)
    order = sorted(set(accs), key=lambda x: x.encode("utf-8"))
    for i, a in enumerate(accs):
        for j, b_ in enumerate(accs):
            assert (r[i] < r[j]) == (order.index(a) < order.index(b_))
    r2 = E.accession_ranks(accs[:5])
    for i in range(5):
        for j in range(5):
            assert (r2[i] < r2[j]) == (accs[i] < accs[j])
