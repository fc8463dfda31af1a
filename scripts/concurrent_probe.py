"""Two (or more) independent clustering passes in flight on one GPU: separate contexts, streams and host threads (GPU box)."""
import os, sys, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from ngspeciesid_b200 import engine as E
from ngspeciesid_b200.modules import p_minimizers_shared

n = 100000
seq, qual, off, acc = bench.make_workload(n, 1002)
mg_tab = E.max_gap_table(p_minimizers_shared.p_emp_for(13, 20), 0.1)
ranks = E.accession_ranks(acc)
order = np.arange(n, dtype=np.int32)
NE = int(sys.argv[1]) if len(sys.argv) > 1 else 3
engs = [E.Engine(0) for _ in range(NE)]
for e in engs:
    e.upload(seq, qual, off); e.minimizers(13, 20); e.quality_stats(); e.cluster(13, 20, mg_tab, order, ranks)

def run(k, passes):
    def work(e):
        for _ in range(passes):
            e.minimizers(13, 20); e.quality_stats(); e.cluster(13, 20, mg_tab, order, ranks)
        e.sync()
    th = [threading.Thread(target=work, args=(engs[i],)) for i in range(k)]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return time.perf_counter() - t

for k in range(1, NE + 1):
    run(k, 2)
    dt = run(k, 6)
    print("%d passes in flight: %.2f ms per pass, %.2f M reads/s" % (k, dt * 1e3 / (6 * k), n * 6 * k / dt / 1e6), flush=True)
