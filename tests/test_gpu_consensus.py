"""Consensus hot path (ii) on the GPU against the CPU oracle (oracle/poa_oracle.cpp,
oracle/consensus_oracle.py). Stated tolerance: per-base edit distance between the CUDA consensus and
the oracle's <= 0.5 % of the consensus length (the kernels restate the same algorithm, so the
observed distance is 0); the distance to the synthetic template is reported alongside."""
import os

import numpy as np
import pytest

from oracle import consensus_oracle as co

pytestmark = pytest.mark.gpu
TOL = 0.005


@pytest.fixture(scope="module")
def eng():
    from ngspeciesid_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def species_reads(n, n_species, seed, lo=700, hi=800):
    from ngspeciesid_b200.synth import simulate_reads
    rs = simulate_reads(n, n_species=n_species, len_lo=lo, len_hi=hi, seed=seed)
    groups = {}
    for i in range(len(rs)):
        groups.setdefault((int(rs.species[i]), int(rs.strand[i])), []).append(i)
    tpl = [t.tobytes().decode() for t in rs.templates]
    return rs, groups, tpl


def within_tolerance(a, b):
    return co.edit_distance(a, b) <= TOL * max(len(a), len(b))


def test_draft_consensus_matches_oracle(eng):
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(240, 3, 31, 300, 420)
    recs = [rs.read(i) for i in range(len(rs))]
    eng.upload_records(recs)
    keys = sorted(groups)
    lists = [groups[k][:25] for k in keys]
    got, nodes = C.draft_consensus_batch(eng, lists)
    for k, lst, g in zip(keys, lists, got):
        exp = co.spoa_consensus([recs[i] for i in lst])
        assert within_tolerance(g, exp)
        assert g == exp                       # same algorithm: in practice identical
        t = tpl[k[0]] if k[1] == 0 else co.revcomp(tpl[k[0]])
        assert co.edit_distance(g, t) <= 0.03 * len(t)
    assert (nodes > 300).all()


def test_draft_single_and_tiny_jobs(eng):
    from ngspeciesid_b200.modules import consensus as C
    recs = [("ACGTACGTACGT", "555555555555"), ("ACGTACGAACGT", "555555555555"), ("A", "5"), ("ACGTTCGTACGT", "5+5+5+5+5+5+")]
    eng.upload_records(recs)
    got, _n = C.draft_consensus_batch(eng, [[0], [0, 1, 3], [2], [2, 0]])
    exp = [co.spoa_consensus([recs[i] for i in l]) for l in ([0], [0, 1, 3], [2], [2, 0])]
    assert got == exp


def test_polish_matches_oracle_and_template(eng):
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(160, 2, 33)
    recs = [rs.read(i) for i in range(len(rs))]
    eng.upload_records(recs)
    keys = sorted(groups)
    lists = [groups[k][:30] for k in keys]
    drafts, _n = C.draft_consensus_batch(eng, [l[:8] for l in lists])
    polished = C.polish_batch(eng, drafts, lists, 2)
    for k, lst, d, p in zip(keys, lists, drafts, polished):
        exp = co.racon_polish(d, [recs[i] for i in lst], 2)
        assert within_tolerance(p, exp)
        assert p == exp
        t = tpl[k[0]] if k[1] == 0 else co.revcomp(tpl[k[0]])
        assert co.edit_distance(p, t) <= 0.01 * len(t)


def test_polish_with_reads_that_do_not_span_their_windows(eng):
    """racon aligns a window layer that does not span the window to within 1 % to the sub-graph between its
    first and last backbone position (window.cpp generate_consensus, spoa Graph::subgraph): reads cut at
    both ends polish a draft together with full-length ones, GPU == oracle; and the sub-graph entry of the
    library agrees with the oracle's on explicit ranges."""
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(140, 1, 37, 1150, 1250)
    recs = [rs.read(i) for i in range(len(rs))]
    fw = [i for i in range(len(recs)) if rs.strand[i] == 0]
    rng = np.random.default_rng(5)
    cut = []
    for i in fw[30:42]:                                   # fragments: cut at both ends by 20-400 bases
        s, q = recs[i]
        a, b = int(rng.integers(20, 400)), int(rng.integers(20, 400))
        cut.append((s[a:len(s) - b], q[a:len(q) - b]))
    use = recs + cut
    eng.upload_records(use)
    full = fw[:30]                                         # enough of them to keep racon's coverage trimming off the ends
    parts = list(range(len(recs), len(use)))
    draft, _ = C.draft_consensus_batch(eng, [fw[:6]])
    pol = C.polish_batch(eng, draft, [full + parts], 2)[0]
    exp = co.racon_polish(draft[0], [use[i] for i in full + parts], 2)
    assert pol == exp
    assert co.edit_distance(pol, tpl[0]) <= 0.01 * len(tpl[0])
    # the fragments really took the sub-graph route: without them the result differs or they were layers
    only_full = co.racon_polish(draft[0], [use[i] for i in full], 2)
    assert co.edit_distance(exp, tpl[0]) <= co.edit_distance(only_full, tpl[0])
    # explicit ranges through the C ABI
    backbone = draft[0][:500]
    lay = [use[i] for i in full[:5]]
    segs = [(s[100:340], q[100:340]) for s, q in lay]
    eng.upload_records(segs)
    n = len(segs)
    job_off = [0, n + 1]
    src = [-1] + list(range(n)); beg = [0] * (n + 1); ln = [len(backbone)] + [len(s) for s, _ in segs]
    sb = [-1] + [95] * n; se = [-1] + [345] * n
    got, _nodes = eng.poa_consensus(job_off, src, beg, ln, aux=[backbone], mode=1, match=3, mismatch=-5, gap=-4, trim=False,
                                    layer_sub=(sb, se))
    want = co.poa_consensus([backbone] + [s for s, _ in segs], [""] + [q for _, q in segs], mode=1, match=3, mismatch=-5, gap=-4,
                            trim=False, sub=[None] + [(95, 345)] * n)
    assert got[0] == want
    whole = co.poa_consensus([backbone] + [s for s, _ in segs], [""] + [q for _, q in segs], mode=1, match=3, mismatch=-5, gap=-4,
                             trim=False)
    assert want != whole or len(want) == len(whole)


def test_polish_mixed_strands(eng):
    """A centre that absorbed its reverse-complement cluster (consensus.py:148-183): reads of both
    strands polish it, each in the orientation that aligns better."""
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(120, 1, 35)
    recs = [rs.read(i) for i in range(len(rs))]
    rc = [(co.revcomp(s), q[::-1]) for s, q in recs]
    eng.upload_records(recs + rc)
    n = len(recs)
    fw = [i for i in range(n) if rs.strand[i] == 0][:6]
    draft, _ = C.draft_consensus_batch(eng, [fw])
    use = list(range(40))
    pol = C.polish_batch(eng, draft, [use], 1, rc_lists=[[n + i for i in use]])[0]
    exp = co.racon_polish(draft[0], [recs[i] for i in use], 1, both_strands=True)
    assert pol == exp
    assert co.edit_distance(pol, tpl[0]) <= 0.01 * len(tpl[0])


def test_run_spoa_and_run_racon_files(eng, tmp_path):
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(60, 1, 37, 400, 450)
    idx = [i for i in range(len(rs)) if rs.strand[i] == 0][:20]
    fq = tmp_path / "reads_c_id_7.fq"
    with open(fq, "w") as f:
        for i in idx:
            s, q = rs.read(i)
            f.write("@%s_1.0\n%s\n+\n%s\n" % (rs.name(i), s, q))
    center = C.run_spoa(str(fq), str(tmp_path / "spoa_tmp.fa"), "spoa")
    assert open(tmp_path / "spoa_tmp.fa").readlines()[1].strip() == center
    assert center == co.spoa_consensus([rs.read(i) for i in idx])
    cf = tmp_path / "consensus_reference_7.fasta"
    cf.write_text(">consensus_cl_id_7_total_supporting_reads_20\n%s\n" % center)
    out = tmp_path / "racon_cl_id_7"
    os.makedirs(out)
    C.run_racon(str(fq), str(cf), str(out), "1", 2)
    lines = open(out / "consensus.fasta").readlines()
    assert len(lines) == 2 and lines[0].startswith(">consensus_cl_id_7")
    exp = co.racon_polish(center, [rs.read(i) for i in idx], 2, both_strands=True)
    assert lines[1].strip() == exp
    assert os.path.exists(out / "racon_polished_it_1.fasta")


def test_form_draft_consensus_and_polish_sequences_files(eng, tmp_path):
    """The two drivers with the reference's file layout (consensus.py:249-278, :186-246): clusters
    above the abundance cut-off get a draft (all in one batch) and are polished (all in one batch);
    every centre must equal what the oracle computes for it alone, and the files must be there."""
    from types import SimpleNamespace
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(90, 3, 53, 380, 440)
    sorted_fq = tmp_path / "sorted.fastq"
    accs = []
    with open(sorted_fq, "w") as f:
        for i in range(len(rs)):
            s, q = rs.read(i)
            accs.append("%s_%d.5" % (rs.name(i), 1000 - i))
            f.write("@%s\n%s\n+\n%s\n" % (accs[-1], s, q))
    clusters, reps = {}, {}
    for (sp, st), idx in groups.items():
        c_id = idx[0]
        clusters[c_id] = [accs[i] for i in idx]
        reps[c_id] = (c_id, 0, accs[c_id], "", "", float(1000 - c_id))
    clusters[10 ** 6] = ["lonely"]                                   # a singleton: no consensus
    reps[10 ** 6] = (10 ** 6, 0, "lonely", "", "", 1.0)
    work = tmp_path / "work"
    os.makedirs(work)
    args = SimpleNamespace(outfolder=str(tmp_path), max_seqs_for_consensus=12, racon=True, racon_iter=2, medaka=False, device=0)
    big = sorted(clusters.items(), key=lambda x: (len(x[1]), reps[x[0]][5]), reverse=True)
    cutoff = len(big[2][1])                                          # the three largest clusters pass
    centers = C.form_draft_consensus(clusters, reps, str(sorted_fq), str(work), cutoff, args)
    expect_ids = [c for c, a in big if len(a) >= cutoff]
    assert [c[1] for c in centers] == expect_ids
    idx_of = {a: i for i, a in enumerate(accs)}
    drafts = {}
    for n, c_id, cons, path in centers:
        used = [rs.read(idx_of[a]) for a in clusters[c_id][:12]]
        assert n == len(clusters[c_id]) and os.path.exists(path)
        assert cons == co.spoa_consensus(used)
        drafts[c_id] = (cons, used)
    out = C.polish_sequences([list(c[:3]) + [[c[3]]] for c in centers], args)
    for n, c_id, cons, _paths in out:
        d, used = drafts[c_id]
        assert cons == co.racon_polish(d, used, 2, both_strands=True)
        assert open(tmp_path / ("racon_cl_id_%d" % c_id) / "consensus.fasta").readlines()[1].strip() == cons
        assert open(tmp_path / ("consensus_reference_%d.fasta" % c_id)).readline().startswith(
            ">consensus_cl_id_%d_total_supporting_reads_%d" % (c_id, n))
        assert os.path.exists(tmp_path / ("reads_to_consensus_%d.fastq" % c_id))


def test_highest_aln_identity(eng):
    from ngspeciesid_b200.modules import consensus as C
    rng = np.random.default_rng(3)
    a = "".join(rng.choice(list("ACGT"), size=600))
    b = a[:300] + "T" + a[300:]
    for x, y in ((a, b), (a, co.revcomp(b)), (a, "".join(rng.choice(list("ACGT"), size=500)))):
        ops_f, _ = co.align_ops(x, y)
        ops_r, _ = co.align_ops(x, co.revcomp(y))
        exp = max(ops_f.count("=") / float(len(ops_f)), ops_r.count("=") / float(len(ops_r)))
        assert C.highest_aln_identity(x, y) == exp
    cents = [[10, 1, a, "p1"], [8, 2, co.revcomp(b), "p2"], [5, 3, "".join(rng.choice(list("ACGT"), size=600)), "p3"]]
    out = C.detect_reverse_complements(cents, 0.9)
    assert [c[:2] for c in out] == [[18, 1], [5, 3]] and out[0][3] == ["p1", "p2"]




def test_all_reads_of_a_large_cluster(eng):
    """The reference's default --max_seqs_for_consensus -1 (NGSpeciesID:204) feeds EVERY read of a
    cluster to the POA. 1 000 reads of one strand of one species: the graph grows to several
    thousand rows (host graphs grow on demand, no max_nodes bound); result = oracle."""
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(2100, 1, 61, 420, 460)
    recs = [rs.read(i) for i in range(len(rs))]
    eng.upload_records(recs)
    lst = groups[(0, 0)][:1000]
    assert len(lst) == 1000
    got, nodes = C.draft_consensus_batch(eng, [lst])
    exp = co.spoa_consensus([recs[i] for i in lst])
    assert within_tolerance(got[0], exp) and got[0] == exp
    assert int(nodes[0]) > 1500
    assert co.edit_distance(got[0], tpl[0]) <= 0.01 * len(tpl[0])
    assert eng.poa_cells() > 3e8
    pol = C.polish_batch(eng, got, [lst], 1)[0]
    assert pol == co.racon_polish(got[0], [recs[i] for i in lst], 1)


def test_five_thousand_read_cluster_completes(eng):
    """a cluster with >= 5 000 reads through the draft and one polishing round with all
    reads. Too large for the scalar oracle to follow in a test, so the check is against the
    template (<= 1 % per base) and that a capped graph (max_nodes) is a clean error."""
    from ngspeciesid_b200 import _lib
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(10400, 1, 67, 380, 420)
    recs = [rs.read(i) for i in range(len(rs))]
    eng.upload_records(recs)
    lst = groups[(0, 0)][:5000]
    assert len(lst) == 5000
    got, nodes = C.draft_consensus_batch(eng, [lst])
    assert co.edit_distance(got[0], tpl[0]) <= 0.05 * len(tpl[0])      # the local-mode draft has ragged ends
    pol = C.polish_batch(eng, got, [lst], 1)[0]
    assert co.edit_distance(pol, tpl[0]) <= 0.03 * len(tpl[0])           # ends are coverage-trimmed (racon)
    assert co.edit_distance(pol[15:-15], tpl[0][20:-20]) <= 0.02 * len(tpl[0]) + 10
    with pytest.raises(_lib.NgsidError):
        C.draft_consensus_batch(eng, [lst[:400]], max_nodes=600)
    got2, _n = C.draft_consensus_batch(eng, [lst[:20]])            # the context stays usable
    assert got2[0] == co.spoa_consensus([recs[i] for i in lst[:20]])


def test_long_layers_use_the_wide_block(eng):
    """Layers of 1 500 - 2 000 bases (PacBio configuration): more than 8 column tiles per row."""
    from ngspeciesid_b200.modules import consensus as C
    rs, groups, tpl = species_reads(40, 1, 71, 1500, 2000)
    recs = [rs.read(i) for i in range(len(rs))]
    eng.upload_records(recs)
    lst = groups[(0, 0)][:12]
    got, _n = C.draft_consensus_batch(eng, [lst])
    assert got[0] == co.spoa_consensus([recs[i] for i in lst])


def test_sample_h1_consensus_racon_all_reads(p_table):
    """BASELINE configs[0]: sample_h1 --ont --consensus --racon with every read of a cluster in the
    POA (max_seqs_for_consensus -1), through the N-GPU driver on one GPU, against the oracle."""
    import test_gpu_multi as TM
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200 import multi_gpu as M
    from conftest import scenario_reads
    from oracle import cluster_oracle as oc
    args = oc.default_args()
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads("h1"), args.k))
    p_emp = oc.load_p_emp(p_table, args.k, args.w)
    engs = [E.Engine(0) for _ in range(4)]
    try:
        engs[0].upload_records([(r[3], r[4]) for r in ra])
        pipe = M.Pipeline(*engs, k=args.k, w=args.w)
        pipe.cluster(E.max_gap_table(p_emp, 0.1), [r[2] for r in ra], [r[5] for r in ra], 0, len(ra))
        centers, info = pipe.consensus(0.1, -1, 2)
    finally:
        for e in engs:
            e.close()
    clusters, reps = oc.single_clustering(list(ra), p_emp, args)
    import test_multi_gpu_gloo as T
    by_acc = {r[2]: r for r in ra}
    drafts, exp = [], []
    # reference semantics with the oracle, all reads, 2 polishing rounds
    cents = []
    for c_id, accs in sorted(clusters.items(), key=lambda x: (len(x[1]), reps[x[0]][5]), reverse=True):
        if len(accs) >= int(0.1 * len(ra)):
            recs = [(by_acc[a][3], by_acc[a][4]) for a in accs]
            cents.append([len(accs), c_id, co.spoa_consensus(recs), recs])
    assert info["drafts"] == [c[2] for c in cents]
    assert len(cents) == 2 and len(centers) == 1                     # the two strand clusters merge
    merged = co.racon_polish(cents[0][2], cents[0][3] + cents[1][3], 2, both_strands=True)
    assert centers[0][:2] == [cents[0][0] + cents[1][0], cents[0][1]]
    assert co.edit_distance(centers[0][2], merged) <= TOL * len(merged)
    assert centers[0][2] == merged
