#!/usr/bin/env python
"""
Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported as Python) with the parasail shim of oracle/parasail_shim.
Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (all committed):
  sample_h1.fastq.gz, supp1_1000.fastq.gz   reference test data (inputs)
  minimizers.json.gz                        cluster.get_kmer_minimizers on compressed reads
  primitives.json.gz                        per-read error rates, scores, hit tables, decisions
  clusters_<scenario>.json.gz               full pipeline results (final_clusters.tsv content)
  ../../ngspeciesid_b200/data/p_shared_table.npz   the empirical probability table (data)
"""
import gzip
import hashlib
import itertools
import json
import os
import runpy
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
SHIM = os.path.join(ROOT, "oracle", "parasail_shim")
sys.path.insert(0, SHIM)
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
os.environ["PYTHONPATH"] = SHIM + os.pathsep + REF + os.pathsep + os.environ.get("PYTHONPATH", "")

import numpy as np  # noqa: E402
from modules import cluster as ref_cluster  # noqa: E402
from modules import help_functions as ref_help  # noqa: E402
from modules import p_minimizers_shared as ref_table  # noqa: E402
from modules import get_sorted_fastq_for_cluster as ref_sort  # noqa: E402

from ngspeciesid_b200.synth import simulate_reads  # noqa: E402


def dump(name, obj):
    with gzip.open(os.path.join(HERE, name), "wt", compresslevel=9) as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name)


def copy_fixture(src, dst, n_records=None):
    with open(src) as f:
        lines = f.readlines()
    if n_records is not None:
        lines = lines[:4 * n_records]
    with gzip.open(os.path.join(HERE, dst), "wt", compresslevel=9) as f:
        f.writelines(lines)
    print("wrote", dst)


def write_plain(gz_name, path):
    with gzip.open(os.path.join(HERE, gz_name), "rt") as f, open(path, "w") as g:
        shutil.copyfileobj(f, g)


def run_reference(fastq, extra, tag):
    """One fresh interpreter per run: the reference calls mp.set_start_method('spawn')
    unconditionally (get_sorted_fastq_for_cluster.py:87), which can only happen once per process."""
    import subprocess
    out = tempfile.mkdtemp(prefix="ref_" + tag + "_")
    argv = ["NGSpeciesID", "--fastq", fastq, "--outfolder", out] + extra
    code = ("import sys, runpy; sys.path[:0] = [%r, %r]; sys.argv = %r; "
            "runpy.run_path(%r, run_name='__main__')" % (SHIM, REF, argv, os.path.join(REF, "NGSpeciesID")))
    subprocess.check_call([sys.executable, "-c", code], stdout=subprocess.DEVNULL)
    return out


def parse_outputs(out):
    sorted_acc = []
    with open(os.path.join(out, "sorted.fastq")) as f:
        for acc, (seq, qual) in ref_help.readfq(f):
            sorted_acc.append(acc)
    stripped = ["_".join(a.split("_")[:-1]) for a in sorted_acc]
    scores = [a.split("_")[-1] for a in sorted_acc]
    idx_of = {}
    for i, a in enumerate(stripped):
        idx_of.setdefault(a, []).append(i)
    clusters = []
    with open(os.path.join(out, "final_clusters.tsv")) as f:
        for line in f:
            cid, acc = line.rstrip("\n").split("\t")
            cid = int(cid)
            while len(clusters) <= cid:
                clusters.append([])
            lst = idx_of[acc]
            clusters[cid].append(lst[0] if len(lst) == 1 else lst.pop(0))
    origins = []
    with open(os.path.join(out, "final_cluster_origins.tsv")) as f:
        for line in f:
            cid, acc, seq, qual, score, err = line.rstrip("\n").split("\t")
            origins.append([int(cid), idx_of_first(stripped, acc), score, err,
                            hashlib.sha1((seq + "\t" + qual).encode()).hexdigest()[:12]])
    with open(os.path.join(out, "final_clusters.tsv"), "rb") as f:
        tsv_sha = hashlib.sha1(f.read()).hexdigest()
    with open(os.path.join(out, "final_cluster_origins.tsv"), "rb") as f:
        origins_sha = hashlib.sha1(f.read()).hexdigest()
    return dict(sorted_names=stripped, sorted_scores=scores, clusters=clusters, origins=origins,
                final_clusters_sha1=tsv_sha, final_cluster_origins_sha1=origins_sha)


def idx_of_first(stripped, acc):
    return stripped.index(acc)


def scenario(tag, fastq, extra):
    out = run_reference(fastq, extra, tag)
    res = parse_outputs(out)
    res["args"] = extra
    # names are only needed to map the reference's TSV back to sorted indices; keep the order
    # information as indices into the input file instead (smaller)
    in_names = [acc for acc, _ in ref_help.readfq(open(fastq))]
    pos = {}
    for i, a in enumerate(in_names):
        pos.setdefault(a, []).append(i)
    res["sorted_input_index"] = [pos[a].pop(0) if len(pos[a]) > 1 else pos[a][0]
                                 for a in res.pop("sorted_names")]
    dump("clusters_%s.json.gz" % tag, res)
    shutil.rmtree(out)
    print(tag, "clusters:", len(res["clusters"]), "top sizes:",
          sorted((len(c) for c in res["clusters"]), reverse=True)[:6])


def primitives(fastq, k, w, n_take, tag):
    """Per-read primitives straight from the reference's functions."""
    p_emp = {}
    for kk, ww, p, e1, e2 in ref_table.read_empirical_p():
        if int(kk) == k and abs(int(ww) - w) <= 2:
            p_emp[(float(e1), float(e2))] = float(p)
            p_emp[(float(e2), float(e1))] = float(p)
    phred = {chr(i): min(10 ** (-(ord(chr(i)) - 33) / 10.0), 0.79433) for i in range(128)}
    recs = []
    reads = [(acc, seq, qual) for acc, (seq, qual) in ref_help.readfq(open(fastq))][:n_take]
    for acc, seq, qual in reads:
        seqc = "".join(ch for ch, _ in itertools.groupby(seq))
        if len(seqc) < k or len(seq) < 2 * k:
            continue
        mins = ref_cluster.get_kmer_minimizers(seqc, k, w)
        runs = [len(list(g)) for _, g in itertools.groupby(seq)]
        qc, start = [], 0
        for h in runs:
            qc.append(min(qual[start:start + h], key=lambda x: phred[x]))
            start += h
        qc = "".join(qc)
        err_c = sum([qc.count(c) * phred[c] for c in set(qc)]) / float(len(qc))
        err_u = sum([qual.count(c) * phred[c] for c in set(qual)]) / float(len(seq))
        exp_err = ref_sort.expected_number_of_erroneous_kmers(qual, k)
        score = (1.0 - exp_err / float(len(seq) - k + 1)) * (len(seq) - k + 1)
        recs.append(dict(len=len(seq), len_c=len(seqc), pos=[p for _, p in mins],
                         kmer_sha=hashlib.sha1("".join(m for m, _ in mins).encode()).hexdigest()[:12],
                         err_c=repr(err_c), err_u=repr(err_u), score=repr(score),
                         bucket=ref_cluster.p_shared_minimizer_empirical(err_c, 0.01, p_emp)))
    return dict(tag=tag, k=k, w=w, reads=recs)


def minimizer_cases():
    """Edge cases of get_kmer_minimizers incl. inputs shorter than w (truncated windows)."""
    rng = np.random.default_rng(5)
    cases = []
    for k, w in ((13, 20), (15, 50), (13, 13), (10, 100)):
        for ln in list(range(k, w + 3)) + [w + 10, 200, 777]:
            for rep in range(2):
                # homopolymer-free random string (as after compression) and a low-complexity one
                if rep == 0:
                    s = []
                    while len(s) < ln:
                        c = "ACGT"[rng.integers(4)]
                        if not s or s[-1] != c:
                            s.append(c)
                    s = "".join(s)
                else:
                    s = ("ACAC" * ln)[:ln]
                mins = ref_cluster.get_kmer_minimizers(s, k, w)
                cases.append(dict(k=k, w=w, seq=s, mins=[[m, p] for m, p in mins]))
    return cases


def main():
    copy_fixture(os.path.join(REF, "test", "sample_h1.fastq"), "sample_h1.fastq.gz")
    copy_fixture(os.path.join(REF, "test", "Supplementary_File1_reads.fastq"), "supp1_1000.fastq.gz", 1000)

    # probability table (data, not code): dense arrays indexed [k][w][e1][e2]
    rows = ref_table.read_empirical_p()
    arr = np.array(rows, dtype=np.float64)
    os.makedirs(os.path.join(ROOT, "ngspeciesid_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "ngspeciesid_b200", "data", "p_shared_table.npz"),
                        k=arr[:, 0].astype(np.int16), w=arr[:, 1].astype(np.int16), p=arr[:, 2],
                        e1=np.rint(arr[:, 3] * 100).astype(np.int8),
                        e2=np.rint(arr[:, 4] * 100).astype(np.int8))
    print("wrote p_shared_table.npz", arr.shape)

    tmp = tempfile.mkdtemp(prefix="golden_in_")
    h1 = os.path.join(tmp, "sample_h1.fastq")
    s1k = os.path.join(tmp, "supp1_1000.fastq")
    write_plain("sample_h1.fastq.gz", h1)
    write_plain("supp1_1000.fastq.gz", s1k)
    syn = os.path.join(tmp, "synth2k.fastq")
    simulate_reads(2000, n_species=5, seed=7).write_fastq(syn)
    synpb = os.path.join(tmp, "synthpb.fastq")
    simulate_reads(600, n_species=4, len_lo=1900, len_hi=2000, seed=11, profile="pacbio",
                   per_read_len=(500, 2000)).write_fastq(synpb)

    dump("minimizers.json.gz", minimizer_cases())
    dump("primitives.json.gz", [primitives(h1, 13, 20, 280, "h1"), primitives(s1k, 15, 50, 200, "supp_pb"),
                                primitives(syn, 13, 20, 200, "synth2k")])

    scenario("h1_t1", h1, ["--ont", "--t", "1"])
    scenario("h1_t4", h1, ["--ont", "--t", "4"])
    scenario("h1_sym_t1", h1, ["--ont", "--t", "1", "--symmetric_map_align_thresholds"])
    scenario("supp1k_t1", s1k, ["--ont", "--t", "1"])
    scenario("supp1k_t8", s1k, ["--ont", "--t", "8"])
    scenario("synth2k_t1", syn, ["--ont", "--t", "1"])
    scenario("synth2k_t8", syn, ["--ont", "--t", "8"])
    scenario("synthpb_t1", synpb, ["--isoseq", "--t", "1"])
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
