"""What does a transfer on a second context cost the clustering pass that runs under it? (GPU box)"""
import os, sys, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import bench
from ngspeciesid_b200 import engine as E
from ngspeciesid_b200.modules import p_minimizers_shared

n = 100000
seq, qual, off, acc = bench.make_workload(n, 1002)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
seq, qual, off = pin(seq), pin(qual), pin(off)
mg_tab = E.max_gap_table(p_minimizers_shared.p_emp_for(13, 20), 0.1)
a, b = E.Engine(0), E.Engine(0)
ranks = E.accession_ranks(acc)
order = np.arange(n, dtype=np.int32)
for e in (a, b):
    e.upload(seq, qual, off); e.minimizers(13, 20); e.quality_stats(); e.cluster(13, 20, mg_tab, order, ranks)

def timed_pass(side):
    th = threading.Thread(target=side) if side else None
    a.sync(); b.sync()
    t = time.perf_counter()
    if th: th.start()
    a.cluster(13, 20, mg_tab, order, ranks)
    dt = time.perf_counter() - t
    if th: th.join()
    b.sync()
    return dt * 1e3

def up(): b.upload(seq, qual, off); b.sync()
def k1k0(): b.minimizers(13, 20); b.quality_stats(); b.sync()
def both(): up(); k1k0()
def sleep(): time.sleep(0.004)
def spin():
    t = time.perf_counter()
    while time.perf_counter() - t < 0.004: pass
tseq, tqual = torch.from_numpy(seq), torch.from_numpy(qual)
dseq, dqual = torch.empty_like(tseq, device="cuda"), torch.empty_like(tqual, device="cuda")
tstream = torch.cuda.Stream()
def torch_h2d():
    with torch.cuda.stream(tstream):
        dseq.copy_(tseq, non_blocking=True); dqual.copy_(tqual, non_blocking=True)
    tstream.synchronize()
def torch_d2d():
    with torch.cuda.stream(tstream):
        for _ in range(40): dseq.copy_(dqual, non_blocking=True)
    tstream.synchronize()
def up_nosync(): b.upload(seq, qual, off)
for name, f in (("nothing", None), ("torch H2D 150 MB (no library code)", torch_h2d), ("torch D2D 40 x 75 MB", torch_d2d), ("upload call without sync", up_nosync), ("sleeping thread", sleep), ("spinning python thread (GIL)", spin), ("H2D upload + pack", up), ("K1 + K0", k1k0), ("upload + K1 + K0", both)):
    r = [timed_pass(f) for _ in range(6)]
    print("%-32s pass %.2f ms (min %.2f)" % (name, sum(r[1:]) / 5, min(r)))
