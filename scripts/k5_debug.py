"""Draft jobs of growing depth on both K5 kernel shapes against the oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200.engine import Engine
from ngspeciesid_b200.modules import consensus as C
from ngspeciesid_b200.synth import simulate_reads
from oracle import consensus_oracle as co

eng = Engine(0)
for (lo, hi) in ((300, 420), (500, 520), (700, 800)):
    rs = simulate_reads(120, n_species=1, len_lo=lo, len_hi=hi, seed=33)
    recs = [rs.read(i) for i in range(len(rs))]
    idx = [i for i in range(len(rs)) if rs.strand[i] == 0]
    eng.upload_records(recs)
    for n in (1, 2, 3, 4, 6, 8, 12, 20, 40):
        lst = idx[:n]
        res = {}
        for shape, om in ((0, 0), (1, 0)):
            eng.poa_shape, eng.poa_order_mode = shape, om
            got, nodes = C.draft_consensus_batch(eng, [lst])
            res[(shape, om)] = (got[0], int(nodes[0]))
        exp0 = co.poa_consensus([recs[i][0] for i in lst], [recs[i][1] for i in lst], order_mode=0)
        print(lo, n, "wave==oracle", res[(0, 0)][0] == exp0, "row==oracle", res[(1, 0)][0] == exp0,
              "nodes", res[(0, 0)][1], res[(1, 0)][1], flush=True)
