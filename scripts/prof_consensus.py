"""Where the time of the consensus leg goes (run on the GPU box): python scripts/prof_consensus.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["NGSID_POA_TIMING"] = "1"
import numpy as np
import bench
from ngspeciesid_b200 import engine as E, multi_gpu as M
from ngspeciesid_b200.modules import p_minimizers_shared

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seq, qual, off, acc = bench.make_workload(n, 1002)
p_emp = p_minimizers_shared.p_emp_for(13, 20)
engs = [E.Engine(0) for _ in range(4)]
engs[0].upload(seq, qual, off)
pipe = M.Pipeline(*engs)
pipe.cluster(E.max_gap_table(p_emp, 0.1), acc, [float(a.split("_")[-1]) for a in acc], 0, n)
pipe.consensus(0.02, 200, 3)
pipe.phase = {}
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
pipe.consensus(0.02, 200, 3)
pr.disable()
print("consensus step %.3f s" % (time.perf_counter() - t0), {k: round(v, 3) for k, v in pipe.phase.items()})
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
