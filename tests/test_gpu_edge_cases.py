"""Edge cases of the hot path through the C ABI (B200): empty and single-read inputs, ragged
lengths from below k to several thousand bases, identical reads, reads that the reference skips,
alignment pairs at the extremes of the supported sizes, and the error codes of calls out of order
or out of range. Expected values come from the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import cluster_oracle as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ngspeciesid_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _rand_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), size=n))


def _cluster_vs_oracle(eng, recs, p_table, k=13, w=20, **kw):
    from ngspeciesid_b200 import engine as E
    p_emp = oc.load_p_emp(p_table, k, w)
    ra = [(i, 0, a, s, q, float(len(recs) - i)) for i, (a, s, q) in enumerate(recs)]
    stats = oc.Stats()
    oc.single_clustering(ra, p_emp, oc.default_args(k=k, w=w), stats)
    exp = {rid: wn for rid, wn, _how in stats.trace}
    eng.upload_records([(s, q) for _a, s, q in recs])
    eng.minimizers(k, w)
    eng.quality_stats()
    assign, via, st = eng.cluster(k, w, E.max_gap_table(p_emp, 0.1), np.arange(len(recs)),
                                  E.accession_ranks([a for a, _s, _q in recs]), **kw)
    for i in range(len(recs)):
        if i in exp:
            assert int(assign[i]) == (-1 if exp[i] < 0 else exp[i]), i
        else:
            assert int(assign[i]) == -2, i           # skipped by the reference (cluster.py:266-268)
    return assign, st


def test_empty_input(eng, p_table):
    from ngspeciesid_b200 import engine as E
    z = np.zeros(0, dtype=np.uint8)
    eng.upload(z, z, np.zeros(1, dtype=np.int64))
    eng.minimizers(13, 20)
    eng.quality_stats()
    lc, cnt, km, pos = eng.get_minimizers()
    assert len(lc) == len(cnt) == len(km) == len(pos) == 0
    sc, er = eng.sort_scores(13)
    assert len(sc) == len(er) == 0
    p_emp = oc.load_p_emp(p_table, 13, 20)
    assign, via, st = eng.cluster(13, 20, E.max_gap_table(p_emp, 0.1), np.zeros(0, dtype=np.int32),
                                  np.zeros(0, dtype=np.uint32))
    assert len(assign) == 0 and st["n_processed"] == 0 and st["n_new_reps"] == 0
    assert len(eng.sg_block_align([], [], [], [], 13)) == 0


def test_single_read_and_identical_reads(eng, p_table):
    rng = np.random.default_rng(1)
    s = _rand_seq(rng, 700)
    q = "5" * len(s)
    a, st = _cluster_vs_oracle(eng, [("only", s, q)], p_table)
    assert list(a) == [-1] and st["n_new_reps"] == 1
    a, st = _cluster_vs_oracle(eng, [("r%d" % i, s, q) for i in range(70)], p_table)
    assert list(a) == [-1] + [0] * 69


def test_ragged_lengths_and_skipped_reads(eng, p_table):
    """Lengths from 1 base to 2500 in one batch, homopolymer-only reads, reads whose compressed
    length is below k (skipped by the reference) or below w (its one-minimizer quirk)."""
    rng = np.random.default_rng(2)
    base = _rand_seq(rng, 2500)
    recs = []

    def noisy(t, e):
        out = []
        for ch in t:
            r = rng.random()
            if r < e / 3:
                continue
            out.append(ch if r >= e else "ACGT"[rng.integers(4)])
            if r > 1 - e / 3:
                out.append("ACGT"[rng.integers(4)])
        return "".join(out) or "A"
    for n in [1, 2, 5, 12, 13, 14, 19, 20, 21, 27, 28, 40, 64, 100, 255, 256, 257, 500, 750, 1000, 1500, 2000, 2500]:
        for e in (0.0, 0.06):
            s = noisy(base[:n], e)
            recs.append(("len%d_e%g" % (n, e), s, "".join(chr(33 + int(x)) for x in rng.integers(5, 40, size=len(s)))))
    recs.append(("hpolA", "A" * 300, "I" * 300))
    recs.append(("hpolAC", "A" * 150 + "C" * 150, "I" * 300))
    recs.append(("short_comp", "AAAACCCCGGGGTTTTAAAACCCC" * 3, "I" * 72))     # compressed length 18 < w
    order = rng.permutation(len(recs))
    recs = [recs[i] for i in order]
    for tile in (0, 7):
        _cluster_vs_oracle(eng, recs, p_table, tile_reads=tile)
    _cluster_vs_oracle(eng, recs, p_table, k=15, w=50)


def test_alignment_size_extremes(eng):
    """1-base sequences, very unequal lengths and pairs near the 13 k-base limit of the trace
    kernel against the C oracle (score and block statistic)."""
    lib = oc._lib()
    rng = np.random.default_rng(3)
    long1 = _rand_seq(rng, 12000)
    long2 = "".join(ch for ch in long1 if rng.random() > 0.03)
    reads = ["A", "C", "ACGT", _rand_seq(rng, 750), long1, long2, _rand_seq(rng, 3000), "ACGTACGTACGTAC"]
    eng.upload_records([(s, "5" * len(s)) for s in reads])
    A, B, O, M = [], [], [], []
    for a in range(len(reads)):
        for b in range(len(reads)):
            if len(reads[a]) * len(reads[b]) > 40e6 and not (a, b) in ((4, 5), (5, 4)):
                continue
            A.append(a); B.append(b); O.append(2 + (a + b) % 4); M.append(1 + (a * b) % 13)
    for shape in (0, 1, 2):
        eng.set_option(3, shape)
        try:
            cnt, score = eng.sg_block_align(A, B, O, M, 13, want_score=True)
        finally:
            eng.set_option(3, 0)
        for i in range(len(A)):
            nc, sc = ctypes.c_int(0), ctypes.c_int(0)
            s1, s2 = reads[A[i]], reads[B[i]]
            c = lib.oracle_sg_block_align(s1.encode(), len(s1), s2.encode(), len(s2), O[i], 1, 13, M[i],
                                          ctypes.byref(nc), ctypes.byref(sc))
            assert (int(cnt[i]), int(score[i])) == (c, sc.value), (shape, len(s1), len(s2), O[i], M[i])


def test_error_codes(eng, p_table):
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200._lib import NgsidError
    rng = np.random.default_rng(4)
    s = _rand_seq(rng, 200)
    eng.upload_records([(s, "5" * 200), (s, "5" * 200)])
    p_emp = oc.load_p_emp(p_table, 13, 20)
    mg = E.max_gap_table(p_emp, 0.1)
    with pytest.raises(NgsidError) as ei:                       # cluster before minimizers / quality stats
        eng.cluster(13, 20, mg, np.arange(2), np.arange(2, dtype=np.uint32))
    assert ei.value.code == -5
    with pytest.raises(NgsidError) as ei:
        eng.minimizers(16, 20)                                  # k > 15: outside this build
    assert ei.value.code == -4
    with pytest.raises(NgsidError) as ei:
        eng.minimizers(13, 12)                                  # w < k
    assert ei.value.code == -1
    eng.minimizers(13, 20)
    eng.quality_stats()
    with pytest.raises(NgsidError) as ei:                       # k/w differ from the extracted minimizers
        eng.cluster(12, 19, mg, np.arange(2), np.arange(2, dtype=np.uint32))
    assert ei.value.code == -5
    with pytest.raises(NgsidError) as ei:
        eng.cluster(13, 20, mg, np.array([0, 7]), np.arange(2, dtype=np.uint32))   # read index out of range
    assert ei.value.code == -1
    with pytest.raises(NgsidError) as ei:
        eng.sg_block_align([0], [5], [3], [5], 13)
    assert ei.value.code == -1
    eng.upload_records([("ACGTRYACGT", "5555555555")])              # bases outside ACGT are legal input (exception path)
    # the context is still usable after every error
    eng.upload_records([(s, "5" * 200), (s, "5" * 200)])
    eng.minimizers(13, 20)
    eng.quality_stats()
    a, _v, _st = eng.cluster(13, 20, mg, np.arange(2), E.accession_ranks(["a", "b"]))
    assert list(a) == [-1, 0]


def test_reads_with_bases_outside_acgt(eng, p_table):
    """the reference compares raw characters (modules/cluster.py:19-37, 265), so N, IUPAC
    codes and lower-case bases are legal: they order by character code in the minimizers, are table keys
    like any k-mer, never match in the alignment score (parasail's ACGT matrix) but do count as equal
    columns in the block statistic. Minimizers, alignment statistic and cluster assignments = oracle."""
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200.synth import simulate_reads
    rng = np.random.default_rng(123)
    recs = list(simulate_reads(300, n_species=3, len_lo=350, len_hi=420, seed=9).records())
    mutated = []
    for n, (acc, s, q) in enumerate(recs):
        s = list(s)
        if n % 3 == 0:                                  # a third of the reads carry a few such bases
            for _ in range(int(rng.integers(1, 6))):
                p = int(rng.integers(0, len(s)))
                s[p] = "NNNRYacgtn"[int(rng.integers(10))]
        if n % 50 == 0:                                 # and a run of N (compresses to one N)
            p = int(rng.integers(20, len(s) - 20))
            s[p:p + 6] = list("NNNNNN")
        mutated.append((acc, "".join(s), q))
    ra = oc.read_array_from_sorted(oc.sort_stage(mutated, 13))
    eng.upload_records([(r[3], r[4]) for r in ra])
    eng.minimizers(13, 20)
    eng.quality_stats()
    len_c, counts, kmer, pos = eng.get_minimizers()
    o = 0
    n_special = 0
    for i, r in enumerate(ra):
        seqc, _ = oc.hpol_compress(r[3])
        exp = oc.minimizers(seqc, 13, 20)
        got = [(eng.kmer_string(kmer[o + j], 13), int(pos[o + j])) for j in range(counts[i])]
        assert len_c[i] == len(seqc) and got == exp, i
        n_special += sum(1 for km, _p in exp if set(km) - set("ACGT"))
        o += counts[i]
    assert n_special > 20
    # alignment statistic on pairs that contain such bases on both sides
    a = [i for i in range(len(ra)) if set(ra[i][3]) - set("ACGT")][:40]
    b = a[1:] + a[:1]
    cnt, score = eng.sg_block_align(a, b, [3] * len(a), [9] * len(a), 13, want_score=True)
    for x, y, c, sc in zip(a, b, cnt, score):
        ops, esc = co_align(ra[x][3], ra[y][3], 3)
        assert sc == esc and c == oc._lib().oracle_block_count(ops.encode(), len(ops), 13, 9)
    # the whole pass
    p_emp = oc.load_p_emp(p_table, 13, 20)
    stats = oc.Stats()
    oc.single_clustering(ra, p_emp, oc.default_args(), stats)
    exp = [w for _r, w, _h in stats.trace]
    assign, _via, _st = eng.cluster(13, 20, E.max_gap_table(p_emp, 0.1), np.arange(len(ra)), E.accession_ranks([r[2] for r in ra]))
    assert list(assign) == exp


def co_align(s1, s2, open_pen):
    from oracle import consensus_oracle as co
    return co.align_ops(s1, s2, open_pen)
