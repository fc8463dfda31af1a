"""The N-GPU driver on real GPUs: the data plane of libngsid.so (NCCL) under torchrun with 2 ranks
(skipped on a box with one GPU), and the single-rank degenerate path on one GPU."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import scenario_reads
from oracle import cluster_oracle as oc
from oracle import consensus_oracle as co

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _expected(tag, world, p_table, max_seqs):
    import test_multi_gpu_gloo as T
    args = oc.default_args(nr_cores=world)
    ra = oc.read_array_from_sorted(oc.sort_stage(scenario_reads(tag), args.k))
    if world == 1:
        clusters, reps = oc.single_clustering(list(ra), oc.load_p_emp(p_table, args.k, args.w), args)
    else:
        clusters, reps = oc.parallel_clustering(list(ra), oc.load_p_emp(p_table, args.k, args.w), args)
    drafts, centers = T.expected_consensus(ra, clusters, reps, max_seqs)
    id_of = {r[2]: r[0] for r in ra}
    roots = {}
    for c_id, accs in clusters.items():
        for a in accs:
            roots[id_of[a]] = c_id
    return roots, drafts, centers


@pytest.mark.parametrize("world", [1, 2])
def test_pipeline_on_gpus(world, p_table, tmp_path):
    """Pipeline (cluster -> NCCL gather of representatives -> merge rounds -> exchange of cluster reads ->
    draft -> reverse-complement merge -> second exchange -> polish) against the single-process oracle."""
    if world > _n_gpus():
        pytest.skip("needs %d GPUs" % world)
    max_seqs = 12
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "multi_gpu_worker.py"), "h1", str(tmp_path), str(max_seqs), "gpu"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    outs = [json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r))) for r in range(world)]
    roots, drafts, centers = _expected("h1", world, p_table, max_seqs)
    got = {}
    for o in outs:
        for i, r in enumerate(o["roots"]):
            got[o["lo"] + i] = r
    assert got == roots                                   # bit-identical assignments
    for o in outs:
        assert len(o["drafts"]) == len(drafts) and len(o["centers"]) == len(centers)
        for a, b in zip(o["drafts"], drafts):             # stated tolerance: 0.5 % per base (equal in practice)
            assert co.edit_distance(a, b) <= 0.005 * len(b)
        for a, b in zip(o["centers"], centers):
            assert a[:2] == b[:2]
            assert co.edit_distance(a[2], b[2]) <= 0.005 * len(b[2])
