"""Worker of tests/test_ranks_gloo.py: one rank of a world_size-N gloo group on the CPU. Runs the
rank-sharded --t N driver (ngspeciesid_b200.modules.parallelize.parallel_clustering_ranks) with the
oracle's reads_to_clusters as the per-batch operator and writes the clusters it returns."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, out_dir = sys.argv[1], sys.argv[2]
    import numpy as np
    import torch.distributed as dist
    from conftest import load_golden, scenario_reads, scenario_args
    from oracle import cluster_oracle as oc
    from ngspeciesid_b200.modules import parallelize

    dist.init_process_group("gloo")
    rank = dist.get_rank()
    g = load_golden("clusters_%s.json.gz" % tag)
    args = scenario_args(g)
    args.outfolder = out_dir                      # rank 0 leaves the per-round snapshots there
    z = np.load(os.path.join(ROOT, "ngspeciesid_b200", "data", "p_shared_table.npz"))
    p_table = [(int(k), int(w), float(p), e1 / 100.0, e2 / 100.0)
               for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads(tag), args.k))
    p_emp = oc.load_p_emp(p_table, args.k, args.w)
    calls = []

    def fn(clusters, reps, reads, p, db, bi, a):
        calls.append((bi, len(reads)))
        return oc.reads_to_clusters(clusters, reps, reads, p, db, bi, a)

    clusters, reps = parallelize.parallel_clustering_ranks(ra, p_emp, args, cluster_fn=fn)
    idx_of = {r[2]: r[0] for r in ra}
    got = [[idx_of[a] for a in accs] for _rep, accs in oc.output_order(clusters, reps)]
    origins = [[i, rep, repr(reps[rep][5]), repr(reps[rep][6])]
               for i, (rep, _a) in enumerate(oc.output_order(clusters, reps))]
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump({"clusters": got, "origins": origins, "calls": calls}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
