"""TEST INFRASTRUCTURE: a CPU stand-in for ngspeciesid_b200.engine.Engine that computes everything
with the oracle and moves data over a torch.distributed (gloo) group. It lets the host logic of the
N-GPU driver (ngspeciesid_b200/multi_gpu.py: merge-round scheduling, final cluster composition,
consensus sharding plan, both read exchanges) run at world_size 2 on a machine without GPUs. Never
imported by the product package."""
import numpy as np

from oracle import cluster_oracle as oc
from oracle import consensus_oracle as co


class OracleEngine(object):
    def __init__(self, p_emp, args, group_ops=None):
        self.p_emp, self.args = p_emp, args
        self.rank, self.world = (group_ops.rank, group_ops.world) if group_ops else (0, 1)
        self.ops = group_ops
        self.recs = []
        self.n_reads = 0
        self.offsets = self.h_seq = self.h_qual = None

    # ---- reads
    def _set(self, recs):
        self.recs = list(recs)
        self.n_reads = len(self.recs)
        lens = [len(s) for s, _q in self.recs]
        self.offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.h_seq = np.frombuffer("".join(s for s, _q in self.recs).encode(), dtype=np.uint8)
        self.h_qual = np.frombuffer("".join(q for _s, q in self.recs).encode(), dtype=np.uint8)

    def upload(self, seq, qual, offsets):
        self._set([(seq[offsets[i]:offsets[i + 1]].tobytes().decode(), qual[offsets[i]:offsets[i + 1]].tobytes().decode())
                   for i in range(len(offsets) - 1)])

    def upload_records(self, recs):
        self._set(recs)

    def minimizers(self, k, w):
        pass

    def quality_stats(self):
        pass

    def sync(self):
        pass

    def adopt_device_reads(self):
        pass

    def append_revcomp(self):
        self._set(self.recs + [(co.revcomp(s), q[::-1]) for s, q in self.recs])

    # ---- clustering pass with the semantics of ngsid_cluster
    def cluster(self, k, w, max_gap, order, acc_rank, init_reps=None, tile_reads=0, **kw):
        a = self.args
        acc = lambda r: "%012d" % int(acc_rank[r])          # same order as the real accession strings
        clusters, reps, db, reads = {}, {}, {}, []
        init = [] if init_reps is None else [int(x) for x in init_reps]
        for r in init:
            seq, qual = self.recs[r]
            seqc, runs = oc.hpol_compress(seq)
            qc = oc.compress_quality(qual, runs)
            reps[r] = (r, 1, acc(r), seq, qual, 0.0, oc.poisson_mean(qc) / float(len(qc)), seqc)
            clusters[r] = [acc(r)]
            for km, _p in oc.minimizers(seqc, k, w):
                db.setdefault(km, set()).add(r)
            reads.append((r, 1, acc(r), seq, qual, 0.0))
        b = 2 if init else 0
        for r in order:
            r = int(r)
            seq, qual = self.recs[r]
            reps[r] = (r, b, acc(r), seq, qual, 0.0)
            clusters[r] = [acc(r)]
            reads.append((r, b, acc(r), seq, qual, 0.0))
        st = oc.Stats()
        oc.reads_to_clusters(clusters, reps, reads, self.p_emp, db, 3, a, st)
        win = {rid: wnr for rid, wnr, _h in st.trace}
        how = {rid: h for rid, _w, h in st.trace}
        assign = np.array([win.get(int(r), -2) for r in order], dtype=np.int32)
        via = np.array([{"new": 0, "map": 1, "align": 2}[how.get(int(r), "new")] for r in order], dtype=np.uint8)
        return assign, via, {"n_new_reps": int((assign == -1).sum()), "n_alignments": st.alignments}

    # ---- data plane over gloo
    def allgather_bytes(self, data):
        return self.ops.allgather(bytes(data)) if self.ops else [bytes(data)]

    def allreduce(self, arr, op="sum"):
        if self.ops:
            parts = self.ops.allgather(arr.copy())
            arr[...] = np.sum(parts, axis=0) if op == "sum" else np.max(parts, axis=0)
        return arr

    def gather_representatives(self, reps, dst):
        mine = [self.recs[int(r)] for r in reps]
        parts = self.ops.allgather(mine) if self.ops else [mine]
        dst._set([x for p in parts for x in p])
        return np.array([len(p) for p in parts], dtype=np.int64)

    def exchange_reads(self, read_idx, dest, tags, dst, expect):
        assert all(dest[i] <= dest[i + 1] for i in range(len(dest) - 1))
        mine = [(int(d), int(t), self.recs[int(r)]) for r, d, t in zip(read_idx, dest, tags)]
        parts = self.ops.allgather(mine) if self.ops else [mine]
        got = [[(t, rec) for d, t, rec in p if d == self.rank] for p in parts]
        dst._set([rec for p in got for _t, rec in p])
        assert dst.n_reads <= max(1, expect)
        return np.array([t for p in got for t, _r in p], dtype=np.int64), np.array([len(p) for p in got], dtype=np.int64)

    # ---- consensus kernels through the oracle
    def _src(self, idx, aux):
        return (self.recs[idx][0], self.recs[idx][1]) if idx >= 0 else (aux[-idx - 1], None)

    def poa_consensus(self, job_off, layer_src, layer_begin, layer_len, aux=None, mode=0, match=5, mismatch=-4, gap=-2,
                      trim=False, max_nodes=0, shape=None, order_mode=0, layer_sub=None):
        out = []
        for j in range(len(job_off) - 1):
            seqs, quals, sub = [], [], []
            for l in range(int(job_off[j]), int(job_off[j + 1])):
                s, q = self._src(int(layer_src[l]), aux)
                b, n = int(layer_begin[l]), int(layer_len[l])
                seqs.append(s[b:b + n])
                quals.append(q[b:b + n] if q is not None else "")
                sub.append(None if layer_sub is None or int(layer_sub[0][l]) < 0 else (int(layer_sub[0][l]), int(layer_sub[1][l])))
            out.append(co.poa_consensus(seqs, quals, mode=mode, match=match, mismatch=mismatch, gap=gap, trim=trim, sub=sub))
        return out, np.zeros(len(out), dtype=np.int32)

    def sg_align_paths(self, a, b, open_pen, aux=None, window=500, want_windows=False):
        n = len(a)
        score, nm, nc = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        win = np.full((n, 16, 4), -1, dtype=np.int32)
        for i in range(n):
            s1, s2 = self._src(int(a[i]), aux)[0], self._src(int(b[i]), aux)[0]
            ops, sc = co.align_ops(s1, s2, int(open_pen[i]))
            score[i], nm[i], nc[i] = sc, ops.count("="), len(ops)
            if want_windows:
                for wi, seg in co.window_segments(ops, len(s2), window).items():
                    win[i, wi] = seg
        return (score, nm, nc, win) if want_windows else (score, nm, nc)


class GlooOps(object):
    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out
