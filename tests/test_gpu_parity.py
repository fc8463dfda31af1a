"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden
vectors. Everything here needs a B200 (`-m gpu`)."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden, scenario_reads, scenario_args
from oracle import cluster_oracle as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ngspeciesid_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _check_minimizers(eng, seqs, k, w):
    from ngspeciesid_b200.engine import decode_kmer
    eng.upload_records([(s, "5" * len(s)) for s in seqs])
    eng.minimizers(k, w)
    len_c, counts, kmer, pos = eng.get_minimizers()
    o = 0
    for i, s in enumerate(seqs):
        seqc, _ = oc.hpol_compress(s)
        assert len_c[i] == len(seqc)
        exp = oc.minimizers(seqc, k, w) if len(seqc) >= k else []
        got = [(decode_kmer(kmer[o + j], k), int(pos[o + j])) for j in range(counts[i])]
        assert got == exp, (i, k, w, len(s), len(seqc))
        o += counts[i]


@pytest.mark.parametrize("tag,k,w", [("h1", 13, 20), ("h1", 15, 50), ("supp1k", 13, 20), ("supp1k", 15, 50),
                                     ("synth2k", 13, 20), ("synthpb", 15, 50), ("h1", 10, 100), ("h1", 13, 13)])
def test_k1_minimizers_fixtures(eng, tag, k, w):
    seqs = [s for _a, s, _q in scenario_reads(tag)]
    _check_minimizers(eng, seqs, k, w)
    for variant in (1,):                # the generic warp-per-read kernel must agree as well
        eng.set_option(1, variant)
        try:
            _check_minimizers(eng, seqs, k, w)
        finally:
            eng.set_option(1, 0)


@pytest.mark.parametrize("k,w", [(13, 20), (12, 19), (5, 12), (2, 9)])
def test_k1_fast_kernel_window8(eng, k, w):
    """Shapes served by the thread-per-read kernel (w-k+1 == 8), incl. repeats and odd lengths."""
    rng = np.random.default_rng(k * 100 + w)
    seqs = ["ACACACACACACACACACACACACACACACACACAC" * 3, "ACGT" * 40, "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAC",
            "A" * 100 + "C" * 100 + "G" * 5 + "ACGTTGCA" * 9]
    for n in list(range(1, 140)) + [255, 256, 257, 1000, 3001]:
        p = rng.dirichlet([0.5, 0.5, 0.5, 0.5])
        seqs.append("".join(rng.choice(list("ACGT"), size=n, p=p)))
    _check_minimizers(eng, seqs, k, w)


def test_k1_minimizers_edge_cases(eng):
    cases = load_golden("minimizers.json.gz")
    by_kw = {}
    for c in cases:
        if c["k"] <= 15:
            by_kw.setdefault((c["k"], c["w"]), []).append(c)
    from ngspeciesid_b200.engine import decode_kmer
    for (k, w), cs in by_kw.items():
        eng.upload_records([(c["seq"], "5" * len(c["seq"])) for c in cs])
        eng.minimizers(k, w)
        _lc, counts, kmer, pos = eng.get_minimizers()
        o = 0
        for i, c in enumerate(cs):
            got = [[decode_kmer(kmer[o + j], k), int(pos[o + j])] for j in range(counts[i])]
            assert got == c["mins"], (k, w, len(c["seq"]))
            o += counts[i]
    # homopolymer-rich and degenerate inputs, single base reads, reads shorter than k
    rng = np.random.default_rng(3)
    seqs = ["A" * 50, "ACGT" * 5, "A", "AC", "ACGTACGTACGTAC", "AAAACCCCGGGGTTTT" * 20]
    for _ in range(200):
        n = int(rng.integers(1, 400))
        p = rng.dirichlet([0.3, 0.3, 0.3, 0.3])
        seqs.append("".join(rng.choice(list("ACGT"), size=n, p=p)))
    _check_minimizers(eng, seqs, 13, 20)
    _check_minimizers(eng, seqs, 15, 50)
    _check_minimizers(eng, seqs, 5, 9)


def test_k1_accepts_non_acgt(eng):
    """Bases outside ACGT go through the exception path (tests/test_gpu_edge_cases.py has the full check)."""
    s = "ACGTNACGTTGCANNACGATCGATCGGCTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCTAGCTAGGGCTA"
    eng.upload_records([(s, "5" * len(s))])
    eng.minimizers(13, 20)
    len_c, counts, kmer, pos = eng.get_minimizers()
    seqc, _ = oc.hpol_compress(s)
    assert [(eng.kmer_string(c, 13), int(p)) for c, p in zip(kmer, pos)] == oc.minimizers(seqc, 13, 20)


@pytest.mark.parametrize("tag", ["h1", "supp1k", "synth2k"])
def test_k0_quality_stats(eng, tag):
    recs = scenario_reads(tag)
    eng.upload_records([(s, q) for _a, s, q in recs])
    eng.quality_stats()
    ec, eu, bk = eng.get_quality_stats()
    from ngspeciesid_b200.engine import bucket_values
    vals = bucket_values()
    for i, (_a, s, q) in enumerate(recs):
        seqc, runs = oc.hpol_compress(s)
        qc = oc.compress_quality(q, runs)
        e1 = oc.poisson_mean(qc) / float(len(qc))
        e2 = oc.poisson_mean(q) / float(len(s))
        assert repr(float(ec[i])) == repr(e1), i
        assert repr(float(eu[i])) == repr(e2), i
        assert vals[bk[i]] == oc.error_bucket(e1)


def test_k0_matches_reference_golden(eng):
    g = load_golden("primitives.json.gz")[0]
    recs = [r for r in scenario_reads("h1")[:280]]
    keep = []
    for a, s, q in recs:
        seqc, _ = oc.hpol_compress(s)
        if len(seqc) >= 13 and len(s) >= 26:
            keep.append((s, q))
    eng.upload_records(keep)
    eng.quality_stats()
    ec, eu, _bk = eng.get_quality_stats()
    assert [repr(float(x)) for x in ec] == [r["err_c"] for r in g["reads"]]
    assert [repr(float(x)) for x in eu] == [r["err_u"] for r in g["reads"]]


def _oracle_align(lib, a, b, o, k, m):
    nc, sc = ctypes.c_int(0), ctypes.c_int(0)
    c = lib.oracle_sg_block_align(a.encode(), len(a), b.encode(), len(b), o, 1, k, m, ctypes.byref(nc), ctypes.byref(sc))
    return c, sc.value


@pytest.mark.parametrize("payload_kernel,shape", [(0, 0), (0, 1), (0, 2), (1, 0)])
def test_k4_block_align_random_pairs(eng, payload_kernel, shape):
    eng.set_option(2, payload_kernel)      # 0: DP + trace + traceback kernels, 1: trace-free payload kernel
    eng.set_option(3, shape)               # DP shape: 0 per launch, 1 warp per pair, 2 block per pair (pipelined strips)
    try:
        _k4_random_pairs(eng)
    finally:
        eng.set_option(2, 0)
        eng.set_option(3, 0)


@pytest.mark.parametrize("shape", [0, 1])
def test_k4_thread_per_pair_traceback(eng, shape):
    """The bulk traceback kernel (one thread per pair, option 4 = 2) against the oracle on the same
    random pairs as the warp kernel, and against the warp kernel on scores / identity columns."""
    eng.set_option(4, 2)
    eng.set_option(3, shape)
    try:
        _k4_random_pairs(eng)
        rng = np.random.default_rng(9)
        reads = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(1, 900, size=60)]
        reads += [reads[3][:200] + reads[4][50:], reads[5] * 2]
        eng.upload_records([(s, "5" * len(s)) for s in reads])
        A = [int(x) for x in rng.integers(0, len(reads), size=300)]
        B = [int(x) for x in rng.integers(0, len(reads), size=300)]
        got = eng.sg_align_paths(A, B, [3] * len(A))
        eng.set_option(4, 1)
        exp = eng.sg_align_paths(A, B, [3] * len(A))
        for g, e in zip(got, exp):
            assert (np.asarray(g) == np.asarray(e)).all()
    finally:
        eng.set_option(4, 0)
        eng.set_option(3, 0)


@pytest.mark.parametrize("shape", [1, 2])
def test_k4_paths_identity_and_windows(eng, shape):
    """Score, identity and window breaking points against the oracle's expanded CIGAR."""
    eng.set_option(3, shape)
    try:
        _k4_paths(eng)
    finally:
        eng.set_option(3, 0)


def _k4_paths(eng):
    from oracle import consensus_oracle as co
    rng = np.random.default_rng(5)

    def rnd(n):
        return "".join(rng.choice(list("ACGT"), size=n))

    def noisy(s, e):
        out = []
        for ch in s:
            r = rng.random()
            if r < e / 3:
                continue
            if r < 2 * e / 3:
                out.append("ACGT"[rng.integers(4)])
            out.append(ch if r >= e else "ACGT"[rng.integers(4)])
        return "".join(out) or "A"

    targets = [rnd(760), rnd(1203), rnd(499), rnd(500), rnd(501), rnd(30)]
    reads, A, B = [], [], []
    for ti, t in enumerate(targets):
        for e in (0.0, 0.08, 0.2):
            reads.append(noisy(t, e)); A.append(len(reads) - 1); B.append(-ti - 1)
        reads.append(noisy(t[len(t) // 3:], 0.1)); A.append(len(reads) - 1); B.append(-ti - 1)
        reads.append(rnd(200)); A.append(len(reads) - 1); B.append(-ti - 1)
    eng.upload_records([(s, "5" * len(s)) for s in reads])
    score, nmatch, ncols, win = eng.sg_align_paths(A, B, [3] * len(A), aux=targets, window=500, want_windows=True)
    for i in range(len(A)):
        t = targets[-B[i] - 1]
        ops, sc = co.align_ops(reads[A[i]], t, open_pen=3)
        assert int(score[i]) == sc
        assert int(ncols[i]) == len(ops) and int(nmatch[i]) == ops.count("=")
        exp = co.window_segments(ops, len(t), 500)
        for w in range(16):
            got = tuple(int(x) for x in win[i, w])
            if w in exp:
                assert got == exp[w], (i, w)
            else:
                assert got == (-1, -1, -1, -1), (i, w)
    # both operands from the auxiliary arena (consensus vs consensus, consensus.py:129-145)
    eng.upload_records([("ACGT", "5555")])
    s2, m2, c2 = eng.sg_align_paths([-1, -1], [-2, -1], [3, 3], aux=[targets[0], noisy(targets[0], 0.05)])
    ops, sc = co.align_ops(targets[0], targets[0])
    assert int(s2[1]) == sc and int(m2[1]) == len(targets[0]) == int(c2[1])


def _k4_random_pairs(eng):
    lib = oc._lib()
    rng = np.random.default_rng(11)

    def rnd(n):
        return "".join(rng.choice(list("ACGT"), size=n))

    def mutate(s, e):
        out = []
        for ch in s:
            r = rng.random()
            if r < e / 3:
                continue
            if r < 2 * e / 3:
                out.append("ACGT"[rng.integers(4)])
            if r < e:
                out.append("ACGT"[rng.integers(4)])
                continue
            out.append(ch)
        return "".join(out) or "A"

    reads, pairs = [], []
    for L in [1, 2, 5, 12, 13, 14, 30, 100, 255, 256, 257, 300, 511, 513, 750, 800, 1100, 2000]:
        for e in (0.0, 0.05, 0.15, 0.3):
            a = rnd(L)
            b = mutate(a, e)
            reads += [a, b]
            pairs.append((len(reads) - 2, len(reads) - 1))
        a, b = rnd(L), rnd(max(1, L // 2 + 3))
        reads += [a, b]
        pairs.append((len(reads) - 2, len(reads) - 1))
        pairs.append((len(reads) - 1, len(reads) - 2))
        # overlap (dovetail) and containment
        a = rnd(L + 40)
        reads += [a[: L + 10], a[20:]]
        pairs.append((len(reads) - 2, len(reads) - 1))
    eng.upload_records([(s, "5" * len(s)) for s in reads])
    for k in (13, 15, 5):
        A, B, O, M = [], [], [], []
        for (x, y) in pairs:
            for o in (2, 3, 5):
                for m in (1, 7, k - 2, k):
                    A.append(x); B.append(y); O.append(o); M.append(m)
        cnt, score = eng.sg_block_align(A, B, O, M, k, want_score=True)
        for i in range(len(A)):
            ec, es = _oracle_align(lib, reads[A[i]], reads[B[i]], O[i], k, M[i])
            assert (int(cnt[i]), int(score[i])) == (ec, es), (len(reads[A[i]]), len(reads[B[i]]), O[i], k, M[i])


def _run_scenario(tag, p_table):
    from ngspeciesid_b200.modules import parallelize
    g = load_golden("clusters_%s.json.gz" % tag)
    args = scenario_args(g)
    args.device = 0
    srt = oc.sort_stage(scenario_reads(tag), args.k)
    ra = oc.read_array_from_sorted(srt)
    p_emp = oc.load_p_emp(p_table, args.k, args.w)
    if args.nr_cores > 1:
        clusters, reps = parallelize.parallel_clustering(ra, p_emp, args)
    else:
        clusters, reps = parallelize.single_clustering(ra, p_emp, args)
    idx_of = {r[2]: r[0] for r in ra}
    out = oc.output_order(clusters, reps)
    got = [[idx_of[a] for a in accs] for _rep, accs in out]
    assert got == g["clusters"]
    origins = [[i, rep, repr(reps[rep][5]), repr(reps[rep][6])] for i, (rep, _a) in enumerate(out)]
    assert origins == [o[:4] for o in g["origins"]]
    for rep, _a in out:
        assert reps[rep][7] == oc.hpol_compress(reps[rep][3])[0]
    # the two output files, byte for byte what the reference wrote for this scenario
    import tempfile
    from ngspeciesid_b200.modules import cluster_output
    with tempfile.TemporaryDirectory() as folder:
        cluster_output.write_cluster_tsvs(clusters, reps, folder)
        for name, key in (("final_clusters.tsv", "final_clusters_sha1"), ("final_cluster_origins.tsv", "final_cluster_origins_sha1")):
            with open(os.path.join(folder, name), "rb") as f:
                assert hashlib.sha1(f.read()).hexdigest() == g[key], name


@pytest.mark.parametrize("tag", ["h1_t1", "h1_t4", "h1_sym_t1", "supp1k_t1", "supp1k_t8",
                                 "synth2k_t1", "synth2k_t8", "synthpb_t1"])
def test_clustering_matches_reference_golden(tag, p_table):
    _run_scenario(tag, p_table)


@pytest.mark.gpu
def test_rank_sharded_driver_on_gpu(p_table):
    """parallelize.parallel_clustering_ranks with the GPU operator (single-rank process group here;
    the world_size-2/3 exchange is covered on the CPU by tests/test_ranks_gloo.py)."""
    import torch.distributed as dist
    from ngspeciesid_b200.modules import parallelize
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29917")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        g = load_golden("clusters_h1_t4.json.gz")
        args = scenario_args(g)
        args.device = 0
        ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads("h1_t4"), args.k))
        p_emp = oc.load_p_emp(p_table, args.k, args.w)
        clusters, reps = parallelize.parallel_clustering_ranks(ra, p_emp, args)
        idx_of = {r[2]: r[0] for r in ra}
        got = [[idx_of[a] for a in accs] for _rep, accs in oc.output_order(clusters, reps)]
        assert got == g["clusters"]
    finally:
        dist.destroy_process_group()


def _merge_representatives(bench, eng2, seq, qual, offsets, acc, gathered, params):
    """gathered[b] = global read ids of the representatives batch b ended with: uploads these reads
    only and runs the merge rounds on them -> ({merged representative: winner}, final representatives)."""
    from ngspeciesid_b200 import engine as E
    ids = sorted(set(x for g in gathered for x in g))
    idx = {g: i for i, g in enumerate(ids)}
    parts = [bench.slice_reads(seq, qual, offsets, g, g + 1) for g in ids]
    m_off = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum([len(p[0]) for p in parts], out=m_off[1:])
    eng2.upload(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), m_off)
    eng2.minimizers(13, 20)
    eng2.quality_stats()
    ar = E.accession_ranks([acc[g] for g in ids])
    merges, final = bench.merge_rounds(eng2, params, ar, {b + 1: [idx[x] for x in g] for b, g in enumerate(gathered)}, len(gathered))
    return {ids[a]: ids[b] for a, b in merges.items()}, [ids[i] for i in final]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bench_batches_and_merge_rounds_match_t_n(world, p_table):
    """bench.py's N-GPU decomposition (batch_bounds -> one clustering pass per batch -> ids-only
    merge rounds on rank 0), run here batch after batch on one GPU, gives the clusters of the
    --t N path (oracle, pinned to the reference's --t 4 / --t 8 golden vectors)."""
    import bench
    from ngspeciesid_b200 import engine as E
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads("synth2k"), 13))
    p_emp = oc.load_p_emp(p_table, 13, 20)
    exp_cl, _exp_rp = oc.parallel_clustering(ra, p_emp, oc.default_args(nr_cores=world))
    idx_of = {r[2]: r[0] for r in ra}
    expected = sorted(sorted(idx_of[a] for a in accs) for accs in exp_cl.values())

    seq = np.frombuffer("".join(r[3] for r in ra).encode(), dtype=np.uint8)
    qual = np.frombuffer("".join(r[4] for r in ra).encode(), dtype=np.uint8)
    off = np.zeros(len(ra) + 1, dtype=np.int64)
    np.cumsum([len(r[3]) for r in ra], out=off[1:])
    acc = [r[2] for r in ra]
    params = {"max_gap": E.max_gap_table(p_emp, 0.1)}
    bounds = bench.batch_bounds(np.diff(off), world)
    eng, eng2 = E.Engine(0), E.Engine(0)
    try:
        rep_of, gathered = {}, []
        for b in range(world):
            lo, hi = bounds[b], bounds[b + 1]
            s_seq, s_qual, s_off = bench.slice_reads(seq, qual, off, lo, hi)
            eng.upload(s_seq, s_qual, s_off)
            eng.minimizers(13, 20)
            eng.quality_stats()
            assign, _via, _st = eng.cluster(13, 20, params["max_gap"], np.arange(hi - lo, dtype=np.int32),
                                            E.accession_ranks(acc[lo:hi]))
            for i, a in enumerate(assign):
                if a != -2:
                    rep_of[lo + i] = lo + i if a == -1 else lo + int(a)
            gathered.append([lo + i for i in np.nonzero(assign == -1)[0]])
        merges, final = _merge_representatives(bench, eng2, seq, qual, off, acc, gathered, params)
    finally:
        eng.close(); eng2.close()

    def root(r):
        while r in merges:
            r = merges[r]
        return r
    groups = {}
    for g, r in rep_of.items():
        groups.setdefault(root(r), []).append(g)
    assert sorted(groups) == sorted(final)
    assert sorted(sorted(v) for v in groups.values()) == expected


def _mixed_reads(n, seed):
    """Species reads plus a large share of unrelated junk reads, so that new representatives,
    chains of tentative representatives and re-evaluations all occur."""
    from ngspeciesid_b200.synth import simulate_reads
    rng = np.random.default_rng(seed)
    rs = simulate_reads(n, n_species=6, len_lo=300, len_hi=400, seed=seed)
    recs = list(rs.records())
    junk = []
    for i in range(n // 3):
        L = int(rng.integers(200, 420))
        s = "".join(rng.choice(list("ACGT"), size=L))
        q = "".join(chr(33 + int(x)) for x in rng.integers(8, 30, size=L))
        junk.append(("junk%d" % i, s, q))
    allr = recs + junk
    perm = rng.permutation(len(allr))
    return [allr[i] for i in perm]


@pytest.mark.parametrize("seed,tile", [(21, 0), (22, 64), (23, 1000)])
def test_clustering_matches_oracle_with_many_representatives(eng, p_table, seed, tile):
    from ngspeciesid_b200 import engine as E
    recs = _mixed_reads(900, seed)
    srt = oc.sort_stage(recs, 13)
    ra = oc.read_array_from_sorted(srt)
    args = oc.default_args()
    p_emp = oc.load_p_emp(p_table, 13, 20)
    stats = oc.Stats()
    oc.single_clustering(ra, p_emp, args, stats)
    exp = [(-1 if w < 0 else w) for _rid, w, _how in stats.trace]
    eng.upload_records([(r[3], r[4]) for r in ra])
    eng.minimizers(13, 20)
    eng.quality_stats()
    assign, via, st = eng.cluster(13, 20, E.max_gap_table(p_emp, 0.1), np.arange(len(ra)),
                                  E.accession_ranks([r[2] for r in ra]), tile_reads=tile)
    assert list(assign) == exp
    how = {"new": 0, "map": 1, "align": 2}
    assert list(via) == [how[h] for _r, _w, h in stats.trace]
    assert st["n_new_reps"] == sum(1 for x in exp if x < 0) > 100


def test_clustering_properties_at_scale(eng, p_table):
    """Size-independent properties on a set too large for the oracle to finish quickly: the result
    does not depend on the speculation tile size, representatives are never assigned, every read
    points at a representative, a second run is identical, and the first reads agree with the
    oracle (the greedy pass on a prefix is the prefix of the greedy pass)."""
    import bench
    from ngspeciesid_b200 import engine as E
    seq, qual, off, acc = bench.make_workload(20000, 4242, cache=False)
    p_emp = oc.load_p_emp(p_table, 13, 20)
    mg = E.max_gap_table(p_emp, 0.1)
    eng.upload(seq, qual, off)
    eng.minimizers(13, 20)
    eng.quality_stats()
    ranks = E.accession_ranks(acc)
    order = np.arange(len(acc))
    a1, v1, st1 = eng.cluster(13, 20, mg, order, ranks, tile_reads=0)
    a2, v2, st2 = eng.cluster(13, 20, mg, order, ranks, tile_reads=1024)
    a3, v3, _ = eng.cluster(13, 20, mg, order, ranks, tile_reads=0)
    assert (a1 == a2).all() and (v1 == v2).all() and (a1 == a3).all()
    reps = np.nonzero(a1 == -1)[0]
    assert len(reps) >= 20 and st1["n_new_reps"] == len(reps)
    assigned = a1[a1 >= 0]
    assert np.isin(assigned, reps).all()                      # no chains: targets are representatives
    assert (assigned < np.nonzero(a1 >= 0)[0]).all()          # a read joins an earlier read
    assert st1["n_mapped"] + st1["n_aln_passed"] + len(reps) == len(acc)
    n0 = 1200
    ra = bench.read_array(seq, qual, off, acc, 0, n0)
    stats = oc.Stats()
    oc.single_clustering(ra, p_emp, oc.default_args(), stats)
    assert [w for _r, w, _h in stats.trace] == list(a1[:n0])


def test_k1_properties_at_scale(eng):
    """Minimizer invariants on 50k reads: positions strictly increase, consecutive positions are at
    most W apart, every k-mer code is the code at its position, window coverage is complete."""
    import bench
    seq, qual, off, acc = bench.make_workload(50000, 99, cache=False)
    eng.upload(seq, qual, off)
    k, w = 13, 20
    eng.minimizers(k, w)
    len_c, counts, kmer, pos = eng.get_minimizers()
    starts = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])
    assert (counts > 0).all()
    d = np.diff(pos.astype(np.int64))
    boundary = np.zeros(len(pos), dtype=bool)
    boundary[starts[1:-1]] = True
    inner = ~boundary[1:]
    assert (d[inner] > 0).all() and (d[inner] <= w - k + 1).all()
    first = pos[starts[:-1]]
    last = pos[starts[1:] - 1]
    assert (first <= w - k).all()
    assert (last.astype(np.int64) >= len_c.astype(np.int64) - w).all()
    # spot check codes against the oracle on a sample of reads
    from ngspeciesid_b200.engine import decode_kmer
    rng = np.random.default_rng(1)
    for i in rng.integers(0, len(acc), size=40):
        s = seq[off[i]:off[i + 1]].tobytes().decode()
        sc, _ = oc.hpol_compress(s)
        exp = oc.minimizers(sc, k, w)
        got = [(decode_kmer(kmer[j], k), int(pos[j])) for j in range(starts[i], starts[i + 1])]
        assert got == exp
    # the stream kernel hands out the record ranges per warp from an atomic cursor: sub-range
    # fetches (inside a warp, across warps, single reads) and a second run of the kernel (new
    # ranges) must return the same records as the full fetch
    for b, e in [(0, 1), (5, 37), (31, 33), (1000, 1100), (49999, 50000), (12345, 23456)]:
        lc2, c2, k2, p2 = eng.get_minimizers(b, e)
        assert (lc2 == len_c[b:e]).all() and (c2 == counts[b:e]).all()
        assert (k2 == kmer[starts[b]:starts[e]]).all() and (p2 == pos[starts[b]:starts[e]]).all()
    eng.minimizers_timed(k, w, 2)
    lc3, c3, k3, p3 = eng.get_minimizers()
    assert (lc3 == len_c).all() and (c3 == counts).all() and (k3 == kmer).all() and (p3 == pos).all()


# ---------------------------------------------------------------- sort stage (SURVEY.md 8 f, rank 1)
@pytest.mark.parametrize("tag", ["h1_t1", "supp1k_t1", "synth2k_t1", "synthpb_t1"])
def test_sort_stage_matches_reference_golden(eng, tag):
    """Scores (as the reference prints them), kept reads and their order against the vectors the
    reference itself produced, and against the oracle's restatement including the error rates."""
    from ngspeciesid_b200.modules import get_sorted_fastq_for_cluster as S
    g = load_golden("clusters_%s.json.gz" % tag)
    args = scenario_args(g)
    recs = scenario_reads(tag)
    read_array, error_rates = S.score_records(recs, args.k, 7.0, eng)
    names = [r[0] for r in recs]
    assert [r[0] for r in read_array] == [names[i] for i in g["sorted_input_index"]]
    assert ["{0}".format(r[3]) for r in read_array] == g["sorted_scores"]
    exp = oc.sort_stage(recs, args.k)
    assert [(a + "_{0}".format(s), q1, q2, s) for a, q1, q2, s in read_array] == exp
    # error rates of the kept reads, file order (the reference sorts them afterwards)
    kept = sorted(g["sorted_input_index"])
    exp_e = [oc.poisson_mean(recs[i][2], oc.PHRED_P_UNCAPPED) / float(len(recs[i][2])) for i in kept]
    assert error_rates == exp_e


def test_sort_stage_filters_and_file(eng, tmp_path):
    """Short / degenerate / low-quality reads are skipped exactly like the reference does, ties keep
    file order, and main() writes the reference's sorted.fastq and logfile."""
    import argparse
    from ngspeciesid_b200.modules import get_sorted_fastq_for_cluster as S
    rng = np.random.default_rng(17)
    recs = []
    for i in range(300):
        n = int(rng.integers(5, 120))
        s = "".join(rng.choice(list("ACGT"), size=n))
        if i % 17 == 0:
            s = "A" * n                                          # compresses to one base
        q = "".join(chr(33 + int(x)) for x in rng.integers(2 if i % 5 else 1, 45 if i % 7 else 9, size=n))
        recs.append(("r%d extra" % i, s, q))
    recs += [("dupA", recs[3][1], recs[3][2]), ("dupB", recs[3][1], recs[3][2])]     # equal scores: stable
    for k in (13, 15, 7):
        ra, er = S.score_records(recs, k, 7.0, eng)
        exp = oc.sort_stage(recs, k)
        assert [(a + "_{0}".format(s), x, y, s) for a, x, y, s in ra] == exp
    fq = tmp_path / "in.fastq"
    with open(fq, "w") as f:
        for a, s, q in recs:
            f.write("@%s\n%s\n+\n%s\n" % (a, s, q))
    args = argparse.Namespace(fastq=str(fq), outfolder=str(tmp_path), outfile=str(tmp_path / "sorted.fastq"), k=13,
                              quality_threshold=7.0, nr_cores=1, use_old_sorted_file=False)
    # the shared engine of modules/ is used by main()
    out = S.main(args)
    exp = oc.sort_stage(recs, 13)
    assert open(out).read() == "".join("@{0}\n{1}\n+\n{2}\n".format(a, s, q) for a, s, q, _ in exp)
    assert open(tmp_path / "logfile.txt").read().startswith("Lowest read error rate:")


def _cluster_vs_oracle(eng, p_table, recs, tile=0, presorted=False):
    from ngspeciesid_b200 import engine as E
    ra = oc.read_array_from_sorted(oc.sort_stage(recs, 13)) if not presorted else recs
    p_emp = oc.load_p_emp(p_table, 13, 20)
    stats = oc.Stats()
    oc.single_clustering(ra, p_emp, oc.default_args(), stats)
    exp = [(-1 if w < 0 else w) for _rid, w, _how in stats.trace]
    eng.upload_records([(r[3], r[4]) for r in ra])
    eng.minimizers(13, 20)
    eng.quality_stats()
    assign, via, st = eng.cluster(13, 20, E.max_gap_table(p_emp, 0.1), np.arange(len(ra)),
                                  E.accession_ranks([r[2] for r in ra]), tile_reads=tile)
    assert list(assign) == exp
    how = {"new": 0, "map": 1, "align": 2}
    assert list(via) == [how[h] for _r, _w, h in stats.trace]
    return st, stats


@pytest.mark.parametrize("inline", [1, 2, 8])
def test_alignment_result_chain_beyond_inline_cache(eng, p_table, inline, monkeypatch):
    """the reference tries every candidate tied for the top hit count
    (modules/cluster.py:174-205), so the number of failed candidates a read remembers over the
    rounds of a pass is unbounded. Templates that share a 288-base prefix tie on its minimizers and
    fail every alignment; with the inline part of the per-read cache cut to 1 or 2 entries
    (NGSID_TEST_ACACHE_INLINE) their chains of failed candidates run through the overflow pool."""
    monkeypatch.setenv("NGSID_TEST_ACACHE_INLINE", str(inline))
    rng = np.random.default_rng(77)
    prefix = "".join(rng.choice(list("ACGT"), size=260)) + "AC" * 14
    recs = []
    for t in range(14):
        tmpl = prefix + "T" + "".join(rng.choice(list("ACGT"), size=540))
        for c in range(3):
            s = list(tmpl)
            for _ in range(4):                      # a few substitutions outside the prefix
                p = int(rng.integers(320, len(s)))
                s[p] = "ACGT"[("ACGT".index(s[p]) + 1) % 4]
            recs.append(("t%d_c%d" % (t, c), "".join(s), "I" * len(s)))
    st, stats = _cluster_vs_oracle(eng, p_table, recs, tile=0)
    assert st["n_new_reps"] == 14
    assert stats.alignments >= 20 and stats.aln_passed == 0     # chains of 2-3 failed candidates


@pytest.mark.parametrize("seed", [31, 32])
def test_slot_and_node_pools_grow(eng, p_table, seed, monkeypatch):
    """slots of invalidated tentative representatives are not reused, so slot
    arrays and posting nodes are not bounded by the number of reads. The pools start tiny here
    (NGSID_TEST_SMALL_CAPS) so that the growth path runs many times; result = oracle."""
    monkeypatch.setenv("NGSID_TEST_SMALL_CAPS", "1")
    recs = _mixed_reads(600, seed)
    st, _ = _cluster_vs_oracle(eng, p_table, recs, tile=32)
    assert st["n_new_reps"] > 60


def test_hit_table_matches_get_all_hits(eng):
    """Stand-alone test of the hit table: per (read, representative) the number of read
    minimizers whose k-mer the representative holds and the sum of their positions, against a direct
    restatement of cluster.get_all_hits (modules/cluster.py:43-62) on the oracle's minimizers."""
    recs = scenario_reads("supp1k")[:400]
    eng.upload_records([(s, q) for _a, s, q in recs])
    eng.minimizers(13, 20)
    mins = []
    for _a, s, _q in recs:
        seqc, _ = oc.hpol_compress(s)
        mins.append(oc.minimizers(seqc, 13, 20) if len(seqc) >= 13 else [])
    reps = list(range(0, 400, 9))
    reads = list(range(400))
    cnt, psum = eng.hit_counts(reps, reads)
    db = {}
    for r in reps:
        for km, _p in mins[r]:
            db.setdefault(km, set()).add(r)
    col = {r: c for c, r in enumerate(reps)}
    exp_c = np.zeros_like(cnt); exp_s = np.zeros_like(psum)
    for i in reads:
        for km, p in mins[i]:
            for r in db.get(km, ()):
                if r != i:
                    exp_c[i, col[r]] += 1; exp_s[i, col[r]] += p
    assert (cnt == exp_c).all() and (psum == exp_s).all()
    assert exp_c.max() > 50 and (exp_c > 0).sum() > 2000
