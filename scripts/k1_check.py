"""K1 stream kernel: parity on a sample + timing on 2 M reads (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from ngspeciesid_b200 import engine as E
from oracle import cluster_oracle as oc
seq, qual, off, acc = bench.make_workload(100000, 1002)
eng = E.Engine(0)
eng.upload(seq, qual, off)
eng.minimizers(13, 20)
len_c, counts, kmer, pos = eng.get_minimizers(0, 3000)
o = 0
bad = 0
for i in range(3000):
    s = seq[off[i]:off[i + 1]].tobytes().decode()
    sc, _ = oc.hpol_compress(s)
    exp = oc.minimizers(sc, 13, 20)
    got = [(E.decode_kmer(kmer[o + j], 13), int(pos[o + j])) for j in range(counts[i])]
    bad += (got != exp) or len_c[i] != len(sc)
    o += counts[i]
print("parity: %d of 3000 reads differ" % bad)
eng.set_option(1, 1)            # generic kernel for every read: the same records?
eng.minimizers(13, 20)
lc2, c2, k2, p2 = eng.get_minimizers()
eng.set_option(1, 0)
eng.minimizers(13, 20)
lc1, c1, k1, p1 = eng.get_minimizers()
print("stream == generic on 100 k reads:", bool((c1 == c2).all() and (k1 == k2).all() and (p1 == p2).all() and (lc1 == lc2).all()))
rep = 20
big_seq = np.tile(seq, rep); big_qual = np.tile(qual, rep)
bl = np.tile(np.diff(off), rep)
bo = np.zeros(len(bl) + 1, dtype=np.int64); np.cumsum(bl, out=bo[1:])
eng.upload(big_seq, big_qual, bo)
eng.minimizers_timed(13, 20, 3)
ms = eng.minimizers_timed(13, 20, 10)
nm = int(c1.sum()) * rep
alg = int(((bl + 3) // 4).sum()) + 8 * len(bl) + 8 * nm + 4 * len(bl)
print("K1: %.4f ms for %d reads, %.1f GB/s, frac %.4f" % (ms, len(bl), alg / ms / 1e6, alg / ms / 1e6 / 6552.3))
