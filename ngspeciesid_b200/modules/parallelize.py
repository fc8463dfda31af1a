"""
Drop-in for the reference's modules/parallelize.py (ksahlin/NGSpeciesID v0.3.1): the batch split
and the log2 rounds of pairwise batch merges are the reference's (results depend on --t exactly
as there). The batches of a round are independent; where the reference hands them to a process pool
(modules/parallelize.py:129-164) they run here as concurrent passes on the one GPU -- a few host
threads, each with its own context and stream (`GPU_LANES`): the latency-bound start of one pass (small
speculation tiles, single-pair alignment launches) fills the gaps of another pass's long launches.
"""
import math
import os
from concurrent.futures import ThreadPoolExecutor

from .. import engine as _engine

GPU_LANES = int(os.environ.get("NGSID_GPU_LANES", "3"))      # passes in flight per GPU (measured: 1 / 2 / 3 = 2.94 / 3.57 / 3.67 M reads/s)

from . import cluster, help_functions


def batch_list(lst, nr_cores=1, batch_type="nr_reads", merge_consecutive=False):
    """Reference: modules/parallelize.py:33-81 (generator of batches)."""
    if merge_consecutive:
        bound, cur = 2, []
        for info in lst:
            if info[1] <= bound:
                cur.append(info)
            else:
                yield cur
                bound += 2
                cur = [info]
        yield cur
        return
    if batch_type == "nr_reads":
        size = int(len(lst) / nr_cores) + 1
        for i in range(0, len(lst), size):
            yield lst[i:i + size]
        return
    if batch_type == "total_nt":
        weight = lambda r: len(r[3])
    elif batch_type == "read_lengths_squared":
        weight = lambda r: math.pow(len(r[3]), 2)
    else:
        return
    limit = int(sum(weight(r) for r in lst) / nr_cores) + 1
    cur, acc = [], 0
    for info in lst:
        acc += weight(info)
        cur.append(info)
        if acc >= limit:
            yield cur
            cur, acc = [], 0
    yield cur


def print_intermediate_results(clusters, cluster_seq_origin, args, iter_nr):
    """Snapshot of a merge round as the reference leaves it (modules/parallelize.py:83-104):
    <outfolder>/<iter_nr>/pre_clusters.csv -- cluster id and read name (score suffix cut off), clusters
    largest first, members in list order -- and cluster_origins.csv, one line per representative.
    Nothing reads these files back, in the reference or here."""
    folder = "{0}/{1}".format(args.outfolder, iter_nr)
    help_functions.mkdir_p(folder)
    by_size = sorted(clusters, key=lambda c: len(clusters[c]), reverse=True)
    member_lines, origin_lines = [], []
    for c_id in by_size:
        member_lines.extend("{0}\t{1}\n".format(c_id, acc.rsplit("_", 1)[0] if "_" in acc else "") for acc in clusters[c_id])
        rid, _batch, acc, seq, qual, score, err, _comp = cluster_seq_origin[c_id]
        origin_lines.append("{0}\t{1}\t{2}\t{3}\t{4}\t{5}\n".format(rid, acc, seq, qual, score, err))
    with open(os.path.join(folder, "pre_clusters.csv"), "w") as f:
        f.writelines(member_lines)
    with open(os.path.join(folder, "cluster_origins.csv"), "w") as f:
        f.writelines(origin_lines)


def _snapshot(all_cl, all_rp, args, it):
    # the reference always has an output folder; library callers (and the tests) may not
    if getattr(args, "outfolder", None):
        print_intermediate_results(all_cl, all_rp, args, it)


def _run_batches(cl, rp, batches, p_emp_probs, db, args):
    """reads_to_clusters of every batch of a round (new batch index i + 1), up to GPU_LANES of them at a time."""
    n = len(batches)
    lanes = max(1, min(GPU_LANES, n))
    if lanes == 1:
        return [cluster.reads_to_clusters(cl[i], rp[i], batches[i], p_emp_probs, db[i], i + 1, args) for i in range(n)]
    free = list(range(lanes))

    def work(i):
        slot = free.pop()                                    # list.pop / append are atomic under the GIL
        try:
            _engine.set_engine_slot(slot)
            return cluster.reads_to_clusters(cl[i], rp[i], batches[i], p_emp_probs, db[i], i + 1, args)
        finally:
            free.append(slot)
    with ThreadPoolExecutor(max_workers=lanes) as ex:
        return list(ex.map(work, range(n)))


def parallel_clustering(read_array, p_emp_probs, args):
    """Reference: modules/parallelize.py:107-217 -> (clusters, representatives)."""
    batches = list(batch_list(read_array, args.nr_cores, batch_type=args.batch_type))
    num = args.nr_cores
    cl = [{r[0]: [r[2]] for r in b} for b in batches]
    rp = [{r[0]: tuple(r) for r in b} for b in batches]
    db = [{} for _ in batches]
    it = 1
    while True:
        if len(batches) == 1:
            res = cluster.reads_to_clusters(cl[0], rp[0], batches[0], p_emp_probs, db[0], 1, args)
            return res[1][0], res[1][1]
        all_cl, all_rp, all_db = {}, {}, {}
        results = _run_batches(cl, rp, batches, p_emp_probs, db, args)
        for i in range(len(batches)):
            c, r, d, bi = results[i][i + 1]
            all_cl.update(c)
            all_rp.update(r)
            all_db[bi] = d
        read_array = [(v[0], v[1], v[2], v[3], v[4], v[5]) for _, v in
                      sorted(all_rp.items(), key=lambda x: x[1][5], reverse=True)]
        if num == 1:
            return all_cl, all_rp
        _snapshot(all_cl, all_rp, args, it)
        it += 1
        batches = list(batch_list(read_array, num, batch_type=args.batch_type, merge_consecutive=True))
        num = len(batches)
        cl, rp, db = [], [], []
        for b in batches:
            low = min(r[1] for r in b)
            cl.append({r[0]: all_cl[r[0]] for r in b})
            rp.append({r[0]: all_rp[r[0]] for r in b})
            db.append(all_db[low])


def parallel_clustering_ranks(read_array, p_emp_probs, args, group=None, cluster_fn=None):
    """`parallel_clustering` with the batches of every round sharded over the ranks of a
    torch.distributed process group (one process per GPU; NCCL on the GPU box, gloo in the CPU
    tests). Every rank passes the same score-sorted `read_array`; batch i of a round runs on rank
    i mod world_size, so the result depends on --t (args.nr_cores) exactly as in the reference and
    NOT on the number of GPUs, and every rank returns the same (clusters, representatives).

    The only exchange is the one the reference has when it joins its process pool
    (modules/parallelize.py:153-187): the (clusters, representatives, minimizer_database, batch
    index) tuples of the finished batches, here one all_gather_object per round. After round 1
    these hold the surviving representatives only (tens to thousands of reads).

    `cluster_fn` replaces cluster.reads_to_clusters; the CPU tests pass the oracle's restatement
    (the product path needs a GPU)."""
    import torch.distributed as dist
    fn = cluster_fn if cluster_fn is not None else cluster.reads_to_clusters
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if cluster_fn is None and getattr(args, "device", None) is None:
        args.device = int(os.environ.get("LOCAL_RANK", "0"))        # one GPU per rank
    batches = list(batch_list(read_array, args.nr_cores, batch_type=args.batch_type))
    num = args.nr_cores
    cl = [{r[0]: [r[2]] for r in b} for b in batches]
    rp = [{r[0]: tuple(r) for r in b} for b in batches]
    db = [{} for _ in batches]
    it = 1
    while True:
        if len(batches) == 1:
            # last round runs in one process in the reference too (parallelize.py:142-149)
            out = [None, None]
            if rank == 0:
                res = fn(cl[0], rp[0], batches[0], p_emp_probs, db[0], 1, args)
                out = [res[1][0], res[1][1]]
            dist.broadcast_object_list(out, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            return out[0], out[1]
        mine = []
        for i in range(rank, len(batches), world):
            res = fn(cl[i], rp[i], batches[i], p_emp_probs, db[i], i + 1, args)
            mine.append(res[i + 1])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        all_cl, all_rp, all_db = {}, {}, {}
        for c, r, d, bi in sorted((t for g in gathered for t in g), key=lambda t: t[3]):
            all_cl.update(c)
            all_rp.update(r)
            all_db[bi] = d
        read_array = [(v[0], v[1], v[2], v[3], v[4], v[5]) for _, v in
                      sorted(all_rp.items(), key=lambda x: x[1][5], reverse=True)]
        if num == 1:
            return all_cl, all_rp
        if rank == 0:
            _snapshot(all_cl, all_rp, args, it)
        it += 1
        batches = list(batch_list(read_array, num, batch_type=args.batch_type, merge_consecutive=True))
        num = len(batches)
        cl, rp, db = [], [], []
        for b in batches:
            low = min(r[1] for r in b)
            cl.append({r[0]: all_cl[r[0]] for r in b})
            rp.append({r[0]: all_rp[r[0]] for r in b})
            db.append(all_db[low])


def single_clustering(read_array, p_emp_probs, args):
    """Reference: NGSpeciesID:20-33."""
    clusters = {r[0]: [r[2]] for r in read_array}
    reps = {r[0]: tuple(r) for r in read_array}
    res = cluster.reads_to_clusters(clusters, reps, read_array, p_emp_probs, {}, 1, args)
    return res[1][0], res[1][1]
