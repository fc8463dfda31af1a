"""The CPU oracle (oracle/) against the golden vectors produced by the reference itself
(tests/golden/make_golden.py). CPU only."""
import hashlib
import os

import pytest

from conftest import load_golden, scenario_reads, scenario_args
from oracle import cluster_oracle as oc


def test_minimizers_edge_cases():
    cases = load_golden("minimizers.json.gz")
    assert len(cases) > 100
    for c in cases:
        got = oc.minimizers(c["seq"], c["k"], c["w"])
        assert [[m, p] for m, p in got] == c["mins"], (c["k"], c["w"], len(c["seq"]))


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_primitives(idx):
    g = load_golden("primitives.json.gz")[idx]
    k, w = g["k"], g["w"]
    src = {"h1": "h1", "supp_pb": "supp1k", "synth2k": "synth2k"}[g["tag"]]
    n_take = {"h1": 280, "supp_pb": 200, "synth2k": 200}[g["tag"]]
    reads = scenario_reads(src)[:n_take]
    it = iter(g["reads"])
    n = 0
    for acc, seq, qual in reads:
        seqc, runs = oc.hpol_compress(seq)
        if len(seqc) < k or len(seq) < 2 * k:
            continue
        r = next(it)
        mins = oc.minimizers(seqc, k, w)
        assert len(seq) == r["len"] and len(seqc) == r["len_c"]
        assert [p for _, p in mins] == r["pos"]
        assert hashlib.sha1("".join(m for m, _ in mins).encode()).hexdigest()[:12] == r["kmer_sha"]
        qc = oc.compress_quality(qual, runs)
        assert repr(oc.poisson_mean(qc) / float(len(qc))) == r["err_c"]
        assert repr(oc.poisson_mean(qual) / float(len(seq))) == r["err_u"]
        assert repr(oc.expected_error_free_kmers_score(qual, k)) == r["score"]
        n += 1
    assert n == len(g["reads"])


def check_tsv_files(clusters, reps, g, folder):
    """The two output files against the SHA-1 of the files the reference wrote for this scenario."""
    from ngspeciesid_b200.modules import cluster_output
    n_big, n_all = cluster_output.write_cluster_tsvs(clusters, reps, folder)
    assert n_all == len(clusters) and n_big == sum(1 for a in clusters.values() if len(a) > 1)
    for name, key in (("final_clusters.tsv", "final_clusters_sha1"), ("final_cluster_origins.tsv", "final_cluster_origins_sha1")):
        with open(os.path.join(folder, name), "rb") as f:
            assert hashlib.sha1(f.read()).hexdigest() == g[key], name


def run_oracle_scenario(tag, p_table, tmp_path=None):
    g = load_golden("clusters_%s.json.gz" % tag)
    args = scenario_args(g)
    recs = scenario_reads(tag)
    srt = oc.sort_stage(recs, args.k)
    # sort-stage parity (order + printed score)
    names = [r[0] for r in recs]
    assert [names[i] for i in g["sorted_input_index"]] == ["_".join(s[0].split("_")[:-1]) for s in srt]
    assert [s[0].split("_")[-1] for s in srt] == g["sorted_scores"]
    ra = oc.read_array_from_sorted(srt)
    p_emp = oc.load_p_emp(p_table, args.k, args.w)
    if args.nr_cores > 1:
        clusters, reps = oc.parallel_clustering(ra, p_emp, args)
    else:
        clusters, reps = oc.single_clustering(ra, p_emp, args)
    idx_of = {r[2]: r[0] for r in ra}
    got = [[idx_of[a] for a in accs] for _rep, accs in oc.output_order(clusters, reps)]
    assert got == g["clusters"]
    origins = [[i, rep, repr(reps[rep][5]), repr(reps[rep][6])]
               for i, (rep, _a) in enumerate(oc.output_order(clusters, reps))]
    assert origins == [o[:4] for o in g["origins"]]
    if tmp_path is not None:
        check_tsv_files(clusters, reps, g, str(tmp_path))


@pytest.mark.parametrize("tag", ["h1_t1", "h1_t4", "h1_sym_t1", "supp1k_t1", "supp1k_t8",
                                 "synth2k_t1", "synth2k_t8", "synthpb_t1"])
def test_pipeline_matches_reference(tag, p_table, tmp_path):
    run_oracle_scenario(tag, p_table, tmp_path)
