"""The N > 1 path on the CPU: world_size-2 (and 3) gloo process groups run the rank-sharded --t N
driver with the oracle as the per-batch operator; every rank must return the clustering the
REFERENCE produced with the same --t (tests/golden/clusters_*_t4/_t8), whatever the world size."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("tag,world", [("h1_t4", 2), ("supp1k_t8", 2), ("h1_t4", 3)])
def test_rank_sharded_clustering_matches_reference(tag, world, tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "ranks_worker.py"), tag, str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    g = load_golden("clusters_%s.json.gz" % tag)
    n_batches = int(g["args"][g["args"].index("--t") + 1])
    # per-round snapshots of the --t N path (parallelize.py:83-104), written once (rank 0)
    rounds = 0
    while (n_batches >> rounds) > 1:
        rounds += 1
    for it in range(1, rounds + 1):
        pre = open(os.path.join(str(tmp_path), str(it), "pre_clusters.csv")).read().splitlines()
        org = open(os.path.join(str(tmp_path), str(it), "cluster_origins.csv")).read().splitlines()
        assert len(pre) == sum(len(c) for c in g["clusters"])
        assert len(org) == len(set(l.split("\t")[0] for l in pre))
    assert not os.path.exists(os.path.join(str(tmp_path), str(rounds + 1)))
    for r in range(world):
        out = json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r)))
        assert out["clusters"] == g["clusters"]
        assert out["origins"] == [o[:4] for o in g["origins"]]
        # round 1: batch i (1-based index i + 1) ran on rank i mod world
        mine = list(range(r + 1, n_batches + 1, world))
        assert [bi for bi, _n in out["calls"][: len(mine)]] == mine
