"""Times the K5 draft / polish kernels at several cluster depths (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200.engine import Engine
from ngspeciesid_b200.modules import consensus as C
from ngspeciesid_b200.synth import simulate_reads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rs = simulate_reads(n, n_species=10, seed=1003)
groups = {}
for i in range(len(rs)):
    groups.setdefault((int(rs.species[i]), int(rs.strand[i])), []).append(i)
eng = Engine(0)
eng.upload(rs.seq, rs.qual, rs.offsets)
keys = sorted(groups)
for depth in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "50,200,800").split(",")]:
    lists = [groups[k][:depth] for k in keys]
    nb = sum(int(rs.lengths()[l].sum()) for l in lists)
    t = time.time()
    try:
        drafts, nodes = C.draft_consensus_batch(eng, lists)
    except Exception as e:
        print("depth", depth, "draft failed:", e); continue
    dt = time.time() - t
    t = time.time()
    pol = C.polish_batch(eng, drafts, lists, 1)
    dt2 = time.time() - t
    print("depth %d: %d clusters, draft %.2f s (%.2f Mbp/s, max nodes %d), 1 polish round %.2f s (%.2f Mbp/s)"
          % (depth, len(lists), dt, nb / dt / 1e6, int(nodes.max()), dt2, nb / dt2 / 1e6), flush=True)
