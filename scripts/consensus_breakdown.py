"""Wall-clock breakdown of the consensus leg (draft + polish rounds) into K4 / K5 / host time (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200.engine import Engine
from ngspeciesid_b200.modules import consensus as C
from ngspeciesid_b200.synth import simulate_reads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rs = simulate_reads(n, n_species=10, seed=1003)
groups = {}
for i in range(len(rs)):
    groups.setdefault((int(rs.species[i]), int(rs.strand[i])), []).append(i)
eng = Engine(0)
eng.upload(rs.seq, rs.qual, rs.offsets)
lists = [groups[k][:depth] for k in sorted(groups)]
acc = {"k4": 0.0, "k5": 0.0}
_a, _p = eng.sg_align_paths, eng.poa_consensus
def a(*x, **k):
    t = time.time(); r = _a(*x, **k); acc["k4"] += time.time() - t; return r
def p(*x, **k):
    t = time.time(); r = _p(*x, **k); acc["k5"] += time.time() - t; return r
eng.sg_align_paths, eng.poa_consensus = a, p
for rep in range(2):
    acc["k4"] = acc["k5"] = 0.0
    t = time.time(); drafts, _ = C.draft_consensus_batch(eng, lists); td = time.time() - t
    k5d = acc["k5"]
    print("rep %d draft %.3f s (K5 %.3f)" % (rep, td, k5d), flush=True)
    for it in range(3):
        acc["k4"] = acc["k5"] = 0.0
        t = time.time(); drafts = C.polish_round_batch(eng, drafts, lists); tp = time.time() - t
        print("   polish %d: %.3f s (K4 %.3f, K5 %.3f, host %.3f)" % (it, tp, acc["k4"], acc["k5"], tp - acc["k4"] - acc["k5"]), flush=True)
