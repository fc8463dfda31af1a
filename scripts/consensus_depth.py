"""Consensus accuracy and time against cluster depth (GPU box): draft (spoa-equivalent) and draft + 3
polishing rounds (racon-equivalent) of the same species from its first 20 / 200 / 2000 forward reads;
per-base edit distance of each result to the species template (the checker's edit distance, oracle/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200.engine import Engine
from ngspeciesid_b200.modules import consensus as C
from ngspeciesid_b200.synth import simulate_reads
from oracle import consensus_oracle as co            # checker only

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
depths = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "20,200,2000").split(",")]
n_sp = 4
rs = simulate_reads(n, n_species=10, seed=1003)
tpl = [t.tobytes().decode() if hasattr(t, "tobytes") else t for t in rs.templates]
groups = {}
for i in range(len(rs)):
    if int(rs.strand[i]) == 0:
        groups.setdefault(int(rs.species[i]), []).append(i)
eng = Engine(0)
eng.upload(rs.seq, rs.qual, rs.offsets)
species = sorted(groups)[:n_sp]
print("| depth | draft: edit distance per base (4 species) | after 3 polishing rounds | draft s | polish s |")
print("|---|---|---|---|---|")
for depth in depths:
    lists = [groups[sp][:depth] for sp in species]
    t = time.time(); drafts, nodes = C.draft_consensus_batch(eng, lists); td = time.time() - t
    t = time.time(); pol = C.polish_batch(eng, drafts, lists, 3); tp = time.time() - t
    dd = [co.edit_distance(d, tpl[sp]) / float(len(tpl[sp])) for d, sp in zip(drafts, species)]
    dp = [co.edit_distance(d, tpl[sp]) / float(len(tpl[sp])) for d, sp in zip(pol, species)]
    print("| %d | %s | %s | %.2f | %.2f |" % (min(len(l) for l in lists), " ".join("%.4f" % x for x in dd),
                                              " ".join("%.4f" % x for x in dp), td, tp), flush=True)
