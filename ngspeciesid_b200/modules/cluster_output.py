"""The output files of the clustering path (SURVEY.md section 8 b "on-disk contract", 8 f rank 3).
The reference writes them inline in its main() (NGSpeciesID:96-119); here the same bytes come from
one function so that a caller of the modules gets final_clusters.tsv and
final_cluster_origins.tsv without the CLI. Pinned by the SHA-1 of the files the reference itself
wrote for the golden scenarios (tests/golden/clusters_*.json.gz)."""
import os


def _name(acc):
    # the sort stage appended "_<score>" to every read name (get_sorted_fastq_for_cluster.py:176)
    return "_".join(acc.split("_")[:-1])


def write_cluster_tsvs(clusters, representatives, outfolder):
    """clusters: {rep id: [accession, ...]}, representatives: {rep id: 8-tuple} as returned by
    single_clustering / parallel_clustering. Clusters by (size, representative score) descending,
    members by their score suffix descending (both sorts stable). Returns (number of clusters with
    more than one read, number of clusters)."""
    nontrivial = 0
    with open(os.path.join(outfolder, "final_clusters.tsv"), "w") as out, \
            open(os.path.join(outfolder, "final_cluster_origins.tsv"), "w") as origins:
        ordered = sorted(clusters.items(), key=lambda x: (len(x[1]), representatives[x[0]][5]), reverse=True)
        for out_id, (c_id, accs) in enumerate(ordered):
            _rid, _b, acc, seq, qual, score, error_rate, _comp = representatives[c_id]
            origins.write("{0}\t{1}\t{2}\t{3}\t{4}\t{5}\n".format(out_id, _name(acc), seq, qual, score, error_rate))
            for r_acc in sorted(accs, key=lambda a: float(a.split("_")[-1]), reverse=True):
                out.write("{0}\t{1}\n".format(out_id, _name(r_acc)))
            if len(accs) > 1:
                nontrivial += 1
    return nontrivial, len(clusters)
