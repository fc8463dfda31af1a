"""Where the time of the drop-in call modules.parallelize.single_clustering goes (GPU box)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from baseline import reference_arm
from ngspeciesid_b200.modules import parallelize, p_minimizers_shared
n = 100000
seq, qual, off, acc = bench.make_workload(n, 1002)
ra = bench.read_array(seq, qual, off, acc, 0, n)
a = reference_arm.reference_args(13, 20, 1, None); a.device = 0
p_emp = p_minimizers_shared.p_emp_for(13, 20)
for _ in range(2):
    t = time.perf_counter(); parallelize.single_clustering(list(ra), p_emp, a); print("call %.3f s" % (time.perf_counter() - t))
pr = cProfile.Profile(); pr.enable(); parallelize.single_clustering(list(ra), p_emp, a); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
