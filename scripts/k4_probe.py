"""Times K4 (block alignment) on n pairs of synthetic reads in one call (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ngspeciesid_b200.engine import Engine
from ngspeciesid_b200.synth import simulate_reads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
payload = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rs = simulate_reads(n + 1, n_species=1, seed=7)
eng = Engine(0)
eng.upload(rs.seq, rs.qual, rs.offsets)
eng.set_option(2, payload)
a = np.arange(1, n + 1, dtype=np.int32)
b = np.zeros(n, dtype=np.int32)
o = np.full(n, 2, dtype=np.int32)
m = np.full(n, 9, dtype=np.int32)
eng.sg_block_align(a[:64], b[:64], o[:64], m[:64], 13)
for _ in range(2):
    t = time.time()
    cnt = eng.sg_block_align(a, b, o, m, 13)
    dt = time.time() - t
    cells = float((rs.lengths()[1:] * rs.lengths()[0]).sum())
    print("pairs %d: %.2f ms, %.1f GCUPS (payload=%d), mean count %.1f" % (n, dt * 1e3, cells / dt / 1e9, payload, cnt.mean()), flush=True)
