"""Probe for k2_insert_kernel (run under ncu): 24 unrelated reads -> one tile, all new representatives."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200 import engine as E
from ngspeciesid_b200.modules import p_minimizers_shared
rng = np.random.default_rng(1)
recs = [("".join(rng.choice(list("ACGT"), size=750)), "5" * 750) for _ in range(24)]
eng = E.Engine(0)
eng.upload_records(recs)
eng.minimizers(13, 20)
eng.quality_stats()
p_emp = p_minimizers_shared.p_emp_for(13, 20)
a, v, st = eng.cluster(13, 20, E.max_gap_table(p_emp, 0.1), np.arange(24), E.accession_ranks(["r%02d" % i for i in range(24)]))
print(list(a), st["n_new_reps"])
