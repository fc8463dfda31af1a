"""Latency of one merge round of the --t N path (GPU box): ~20 surviving representatives of one batch
clustered against the table of ~20 representatives of another (modules/parallelize.py:153-217)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from ngspeciesid_b200 import engine as E
from ngspeciesid_b200.modules import p_minimizers_shared

n = 100000
seq, qual, off, acc = bench.make_workload(2 * n, 1002)
p_emp = p_minimizers_shared.p_emp_for(13, 20)
mg_tab = E.max_gap_table(p_emp, 0.1)
eng = E.Engine(0)
eng.upload(seq, qual, off)
eng.minimizers(13, 20); eng.quality_stats()
ranks = E.accession_ranks(acc)
a1, _, _ = eng.cluster(13, 20, mg_tab, np.arange(0, n, dtype=np.int32), ranks)
a2, _, _ = eng.cluster(13, 20, mg_tab, np.arange(n, 2 * n, dtype=np.int32), ranks)
lo = np.nonzero(a1 == -1)[0].astype(np.int32)
hi = (np.nonzero(a2 == -1)[0] + n).astype(np.int32)
print("representatives:", len(lo), len(hi))
for rep in range(3):
    eng.sync(); eng.reset_launch_count()
    t = time.perf_counter()
    for _ in range(20):
        a, _v, st = eng.cluster(13, 20, mg_tab, hi, ranks, init_reps=lo)
    eng.sync()
    dt = (time.perf_counter() - t) / 20
    print("merge pass: %.3f ms, %d launches, device cluster %.3f ms (k4 %.3f, map %.3f), stats %s" % (
        dt * 1e3, eng.launch_count() // 20, eng.phase_ms(3), eng.phase_ms(4), eng.phase_ms(5),
        {k: st[k] for k in ("n_new_reps", "n_alignments", "n_tiles", "n_chain_steps", "n_surprises")}))
