"""The N-GPU driver on the CPU: world_size-2 and -4 gloo groups run ngspeciesid_b200.multi_gpu.Pipeline
on oracle-backed stand-in engines. Checked against single-process computations with the oracle
(itself pinned to the reference's --t N golden clusterings): final clusters incl. the order in which
the reference concatenates merged clusters, the consensus of the final clusters (draft of the capped
read lists, reverse-complement merge, one polishing round with the union of the merged clusters'
reads), identical on every rank."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden, scenario_reads
from oracle import cluster_oracle as oc
from oracle import consensus_oracle as co

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(tag, world, tmp_path, max_seqs):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "multi_gpu_worker.py"), tag, str(tmp_path), str(max_seqs)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    return [json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r))) for r in range(world)]


def expected_consensus(ra, clusters, reps, max_seqs, ratio=0.1, thr=0.9):
    """Reference semantics (NGSpeciesID:124-158) computed with the oracle in one process."""
    by_acc = {r[2]: r for r in ra}
    cutoff = int(ratio * len(ra))
    centers = []
    for c_id, accs in sorted(clusters.items(), key=lambda x: (len(x[1]), reps[x[0]][5]), reverse=True):
        if len(accs) >= cutoff:
            recs = [(by_acc[a][3], by_acc[a][4]) for a in accs[:max_seqs]]
            centers.append([len(accs), c_id, co.spoa_consensus(recs), recs])
    drafts = [c[2] for c in centers]

    def ident(s1, s2):
        best = 0.0
        for t in (s2, co.revcomp(s2)):
            ops, _sc = co.align_ops(s1, t)
            best = max(best, ops.count("=") / float(len(ops)))
        return best
    gone, out = set(), []
    for i, (n, c_id, seq, recs) in enumerate(centers):
        if i in gone:
            continue
        tot, reads = n, list(recs)
        for j in range(i + 1, len(centers)):
            if ident(seq, centers[j][2]) >= thr:
                tot += centers[j][0]
                gone.add(j)
                reads += centers[j][3]
        out.append([tot, c_id, co.racon_polish(seq, reads, 1, both_strands=True)])
    return drafts, out


@pytest.mark.parametrize("tag,world", [("h1", 2), ("h1", 4)])
def test_pipeline_matches_single_process_oracle(tag, world, tmp_path):
    max_seqs = 6
    outs = run_world(tag, world, tmp_path, max_seqs)
    args = oc.default_args(nr_cores=world)
    z = np.load(os.path.join(HERE, "..", "ngspeciesid_b200", "data", "p_shared_table.npz"))
    p_table = [(int(k), int(w), float(p), e1 / 100.0, e2 / 100.0)
               for k, w, p, e1, e2 in zip(z["k"], z["w"], z["p"], z["e1"], z["e2"])]
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads(tag), args.k))
    clusters, reps = oc.parallel_clustering(list(ra), oc.load_p_emp(p_table, args.k, args.w), args)
    id_of = {r[2]: r[0] for r in ra}
    # ---- membership: root of every read
    exp_root = {}
    for c_id, accs in clusters.items():
        for a in accs:
            exp_root[id_of[a]] = c_id
    got_root = {}
    for o in outs:
        for i, r in enumerate(o["roots"]):
            got_root[o["lo"] + i] = r
    assert got_root == exp_root
    if world == 4:          # the reference itself, --t 4
        g = load_golden("clusters_h1_t4.json.gz")
        by_root = {}
        for rid, r in got_root.items():
            by_root.setdefault(r, set()).add(rid)
        assert sorted(sorted(s) for s in by_root.values()) == sorted(sorted(c) for c in g["clusters"])
    # ---- concatenation order of the final cluster lists (modules/cluster.py:338-345 over the rounds)
    local = {}
    for o in outs:
        local.update({int(k): v for k, v in o["local"].items()})
    for o in outs:
        assert o["glist"] == outs[0]["glist"]
    for root, gl in outs[0]["glist"].items():
        got = [m for g in gl for m in local[g]]
        assert got == [id_of[a] for a in clusters[int(root)]]
    # ---- consensus of the final clusters, identical on every rank
    drafts, centers = expected_consensus(ra, clusters, reps, max_seqs)
    for o in outs:
        assert o["drafts"] == drafts
        assert o["centers"] == centers
    assert len(centers) >= 1 and outs[0]["rounds"] == {2: 1, 4: 2}[world]
