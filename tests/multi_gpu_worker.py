"""Worker of tests/test_multi_gpu_gloo.py: one rank of a gloo group on the CPU running the N-GPU driver
(ngspeciesid_b200.multi_gpu.Pipeline) on oracle-backed stand-in engines (tests/fake_engine.py)."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, out_dir, max_seqs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    import numpy as np
    import torch.distributed as dist
    from conftest import scenario_reads
    from fake_engine import GlooOps, OracleEngine
    from oracle import cluster_oracle as oc
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200 import multi_gpu as M

    on_gpu = len(sys.argv) > 4 and sys.argv[4] == "gpu"
    dist.init_process_group("gloo")
    ops = GlooOps(dist)
    rank, world = ops.rank, ops.world
    args = oc.default_args(nr_cores=world)
    z = np.load(os.path.join(ROOT, "ngspeciesid_b200", "data", "p_shared_table.npz"))
    p_table = [(int(k), int(w), float(p), e1 / 100.0, e2 / 100.0)
               for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads(tag), args.k))
    p_emp = oc.load_p_emp(p_table, args.k, args.w)
    lens = np.array([len(r[3]) for r in ra])
    bounds = M.batch_bounds(lens, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = ra[lo:hi]
    if on_gpu:
        # the product path: real engines, NCCL inside libngsid.so (the id travels over the gloo group)
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        engs = [E.Engine(dev) for _ in range(4)]
        if world > 1:
            uid = ops.allgather(E.nccl_unique_id() if rank == 0 else None)[0]
            engs[0].nccl_init(uid, rank, world)
            for e in engs[1:]:
                e.nccl_share(engs[0])
    else:
        engs = [OracleEngine(p_emp, args, ops) for _ in range(4)]
    engs[0].upload_records([(r[3], r[4]) for r in mine])
    pipe = M.Pipeline(*engs, rank=rank, world=world, k=args.k, w=args.w)
    roots = pipe.cluster(E.max_gap_table(p_emp, args.min_prob_no_hits), [r[2] for r in mine], [r[5] for r in mine],
                         lo, len(ra))
    # the same steps through cluster_stream (local pass of step s + 1 on a second batch engine and host thread under
    # the exchange + merge rounds of step s): same roots, and the state of the last step feeds the consensus below
    if on_gpu:
        alt = E.Engine(dev)
        if world > 1:
            alt.nccl_share(engs[0])
    else:
        alt = OracleEngine(p_emp, args, ops)
    alt.upload_records([(r[3], r[4]) for r in mine])
    pipe.alt = alt
    roots_s = pipe.cluster_stream(3, E.max_gap_table(p_emp, args.min_prob_no_hits), [r[2] for r in mine], [r[5] for r in mine],
                                  lo, len(ra))
    assert [int(x) for x in roots_s] == [int(x) for x in roots], "cluster_stream differs from cluster"
    # member lists of the local round-0 clusters (global ids), for the concatenation order
    local = {}
    for i, rp in enumerate(pipe.local_rep_of):
        if rp >= 0:
            local.setdefault(int(lo + rp), []).append(int(lo + i))
    centers, info = pipe.consensus(0.1, max_seqs, 1)
    out = {"roots": [int(x) for x in roots], "lo": lo, "local": local,
           "glist": {str(pipe.g_gid[r]): [int(pipe.g_gid[g]) for g in gl] for r, gl in pipe.ms.glist.items()},
           "centers": centers, "drafts": info["drafts"], "owner": info["owner"], "rounds": pipe.merge_rounds}
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
