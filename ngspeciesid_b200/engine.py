"""
Array-level host API over the C ABI: one Engine per GPU. This is the call surface bench.py times
end to end (host numpy buffers in, host numpy results out) and what the reference-shaped
wrappers in ngspeciesid_b200.modules sit on.
"""
import ctypes
import threading
import math

import numpy as np

from . import _lib
from ._lib import ClusterParams, ClusterStats, NgsidError, PoaParams, as_array, ptr

# PHRED char -> capped error probability, exactly the reference's table (modules/cluster.py:233)
PHRED_P = np.array([min(10 ** (-(c - 33) / 10.0), 0.79433) for c in range(128)], dtype=np.float64)
# uncapped table of the sort stage (modules/get_sorted_fastq_for_cluster.py:21)
PHRED_P_UNCAPPED = np.array([10 ** (-(c - 33) / 10.0) for c in range(128)], dtype=np.float64)


def bucket_thresholds():
    """14 doubles: the smallest x with round(x, 2) >= (b+2)/100, b = 0..13, found with Python's
    own round() so that the device bucketing equals cluster.p_shared_minimizer_empirical
    (modules/cluster.py:356-366)."""
    out = []
    for b in range(14):
        target = round((b + 2) / 100.0, 2)
        x = (b + 1) / 100.0 + 0.005
        while round(x, 2) >= target:
            x = math.nextafter(x, 0.0)
        while round(x, 2) < target:
            x = math.nextafter(x, 1.0)
        out.append(x)
    return np.array(out, dtype=np.float64)


def bucket_values():
    return [round(0.01 * (b + 1), 2) for b in range(15)]


def max_gap_table(p_emp_probs, min_prob_no_hits):
    """max_gap[b1*15+b2]: largest gap g for which the left-to-right product of g factors
    (1 - p_emp[(e1,e2)]) starting from 1 is not < min_prob_no_hits (modules/cluster.py:97-112)."""
    vals = bucket_values()
    out = np.zeros(225, dtype=np.int32)
    for b1, e1 in enumerate(vals):
        for b2, e2 in enumerate(vals):
            q = 1.0 - p_emp_probs[(e1, e2)]
            g, prod = -1, 1
            while not (prod < min_prob_no_hits):
                g += 1
                if g > 1 << 20:
                    break
                prod = prod * q
            out[b1 * 15 + b2] = g
    return out


def accession_ranks(accessions):
    """Rank of each accession string in ascending order (Python compares str by code point, which
    is the byte order of UTF-8) -- the tie-break of modules/cluster.py:79. Equal strings share a
    rank."""
    n = len(accessions)
    if n == 0:
        return np.zeros(0, dtype=np.uint32)
    try:
        arr = np.array(accessions, dtype="S")             # ASCII accessions: one pass inside numpy
    except UnicodeEncodeError:
        arr = np.array([a.encode("utf-8") for a in accessions])
    order = np.argsort(arr, kind="stable")
    srt = arr[order]
    start = np.ones(n, dtype=bool)
    start[1:] = srt[1:] != srt[:-1]
    group_first = np.maximum.accumulate(np.where(start, np.arange(n), 0))
    rank = np.empty(n, dtype=np.uint32)
    rank[order] = group_first.astype(np.uint32)
    return rank


class Engine(object):
    def __init__(self, device=0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.ngsid_ctx_create(device, ctypes.byref(h))
        if rc != 0 or not h:
            raise NgsidError(rc, "ngsid_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device
        self.n_reads = 0
        self.offsets = None
        self._q_done = False
        self.rank, self.world = 0, 1

    def close(self):
        if getattr(self, "h", None):
            pinned = self.__dict__.pop("_pinned", None)
            if pinned is not None:
                self.h_seq = self.h_qual = None              # views into the page-locked buffers
                for p_ in pinned[1]:
                    self.lib.ngsid_pinned_free(p_)
            self.lib.ngsid_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NgsidError(rc, self.lib.ngsid_last_error(self.h).decode("utf-8", "replace"))

    # ---- reads
    def upload(self, seq, qual, offsets):
        seq = as_array(seq, np.uint8)
        qual = as_array(qual, np.uint8)
        offsets = as_array(offsets, np.int64)
        n = len(offsets) - 1
        if len(seq) != offsets[-1] or len(qual) != offsets[-1]:
            raise ValueError("seq/qual length does not match offsets")
        self._check(self.lib.ngsid_upload_reads(self.h, ptr(seq), ptr(qual), ptr(offsets), n))
        self.n_reads = n
        self.offsets = offsets
        self.h_seq, self.h_qual = seq, qual
        self._qcs = None
        self._q_done = False

    def pinned_pair(self, nbytes):
        """Two page-locked uint8 arrays of at least `nbytes` (kept and grown by the engine) for bases and
        qualities: the host layer packs records into them in place and uploads from them."""
        cur = self.__dict__.get("_pinned")
        if cur is None or cur[0] < nbytes:
            if cur is not None:
                self.h_seq = self.h_qual = None              # they may be views into the buffers about to go
                self._pinned = None
                for p_ in cur[1]:
                    self.lib.ngsid_pinned_free(p_)
            cap = int(nbytes + nbytes // 4 + 4096)
            ptrs = []
            for _ in range(2):
                p_ = ctypes.c_void_p()
                rc = self.lib.ngsid_pinned_alloc(ctypes.byref(p_), cap)
                if rc != 0 or not p_:
                    raise NgsidError(rc, "ngsid_pinned_alloc failed")
                ptrs.append(p_)
            arrs = [np.ctypeslib.as_array((ctypes.c_uint8 * cap).from_address(p_.value)) for p_ in ptrs]
            self._pinned = cur = (cap, ptrs, arrs)
        return cur[2][0], cur[2][1]

    def upload_records(self, records):
        """records: iterable of (seq:str, qual:str)."""
        seqs, quals = [], []
        for s, q in records:
            seqs.append(s.encode("ascii"))
            quals.append(q.encode("ascii"))
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offs[1:])
        self.upload(np.frombuffer(b"".join(seqs), dtype=np.uint8), np.frombuffer(b"".join(quals), dtype=np.uint8), offs)

    # ---- K1
    def minimizers(self, k, w):
        self._check(self.lib.ngsid_minimizers(self.h, k, w))

    def minimizers_timed(self, k, w, iters):
        ms = ctypes.c_float(0)
        self._check(self.lib.ngsid_minimizers_timed(self.h, k, w, iters, ctypes.byref(ms)))
        return ms.value

    def get_minimizers(self, begin=0, end=None):
        end = self.n_reads if end is None else end
        n = end - begin
        len_c = np.zeros(n, dtype=np.uint32)
        counts = np.zeros(n, dtype=np.uint32)
        total = ctypes.c_int64(0)
        self._check(self.lib.ngsid_get_minimizers(self.h, begin, end, ptr(len_c), ptr(counts), None, None, 0, ctypes.byref(total)))
        kmer = np.zeros(total.value, dtype=np.uint32)
        pos = np.zeros(total.value, dtype=np.uint32)
        self._check(self.lib.ngsid_get_minimizers(self.h, begin, end, None, None, ptr(kmer), ptr(pos), total.value, ctypes.byref(total)))
        return len_c, counts, kmer, pos

    # ---- K0
    def quality_stats(self):
        thr = bucket_thresholds()
        self._check(self.lib.ngsid_quality_stats(self.h, ptr(PHRED_P), ptr(thr)))
        self._q_done = True

    def sort_scores(self, k):
        """Sort-stage keys of the uploaded reads (get_sorted_fastq_for_cluster.py:23-33,145-152):
        (score, mean error probability with the uncapped table), both float64, bit-identical to the
        reference's arithmetic."""
        score, err = np.zeros(self.n_reads), np.zeros(self.n_reads)
        if self.n_reads:
            self._check(self.lib.ngsid_sort_scores(self.h, k, ptr(PHRED_P), ptr(PHRED_P_UNCAPPED), ptr(score), ptr(err)))
        return score, err

    def get_quality_stats(self, begin=0, end=None):
        end = self.n_reads if end is None else end
        n = end - begin
        ec, eu, bk = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
        self._check(self.lib.ngsid_get_quality_stats(self.h, begin, end, ptr(ec), ptr(eu), ptr(bk)))
        return ec, eu, bk

    # ---- clustering pass
    def cluster(self, k, w, max_gap, order, acc_rank, init_reps=None, min_shared=5, min_fraction=0.8,
                mapped_threshold=0.7, aligned_threshold=0.4, symmetric=False, tile_reads=0):
        p = ClusterParams()
        p.k, p.w, p.min_shared, p.symmetric = k, w, min_shared, 1 if symmetric else 0
        p.min_fraction, p.mapped_threshold, p.aligned_threshold = min_fraction, mapped_threshold, aligned_threshold
        mg = as_array(max_gap, np.int32)
        if mg.shape != (225,):
            raise ValueError("max_gap must have 225 entries")
        for i in range(225):
            p.max_gap[i] = int(mg[i])
        p.tile_reads = tile_reads
        order = as_array(order, np.int32)
        acc_rank = as_array(acc_rank, np.uint32)
        if len(acc_rank) != self.n_reads:
            raise ValueError("acc_rank needs one entry per uploaded read")
        init = None if init_reps is None or len(init_reps) == 0 else as_array(init_reps, np.int32)
        assign = np.zeros(len(order), dtype=np.int32)
        via = np.zeros(len(order), dtype=np.uint8)
        st = ClusterStats()
        self._check(self.lib.ngsid_cluster(self.h, ctypes.byref(p), ptr(order), len(order), ptr(init),
                                           0 if init is None else len(init), ptr(acc_rank), ptr(assign), ptr(via),
                                           ctypes.byref(st)))
        return assign, via, st.as_dict()

    def hit_counts(self, reps, reads):
        """get_all_hits for every (read, representative) pair -> (counts, position sums), shape (len(reads), len(reps))."""
        reps, reads = as_array(reps, np.int32), as_array(reads, np.int32)
        cnt = np.zeros((len(reads), len(reps)), dtype=np.uint32)
        psum = np.zeros((len(reads), len(reps)), dtype=np.uint32)
        self._check(self.lib.ngsid_hit_counts(self.h, ptr(reps), len(reps), ptr(reads), len(reads), ptr(cnt), ptr(psum)))
        return cnt, psum

    # ---- K4 alone
    def sg_block_align(self, read_a, read_b, open_pen, match_id, k, want_score=False):
        a, b = as_array(read_a, np.int32), as_array(read_b, np.int32)
        o, m = as_array(open_pen, np.int32), as_array(match_id, np.int32)
        cnt = np.zeros(len(a), dtype=np.int32)
        score = np.zeros(len(a), dtype=np.int32) if want_score else None
        self._check(self.lib.ngsid_sg_block_align(self.h, ptr(a), ptr(b), ptr(o), ptr(m), len(a), k, ptr(cnt), ptr(score)))
        return (cnt, score) if want_score else cnt

    def sg_align_paths(self, a, b, open_pen, aux=None, window=500, want_windows=False):
        """Semi-global alignments of rows a[i] against columns b[i]; indices < 0 name aux strings.
        Returns (score, n_match, n_cols[, windows (n,16,4)])."""
        a, b = as_array(a, np.int32), as_array(b, np.int32)
        o = as_array(open_pen, np.int32)
        n = len(a)
        aux_seq, aux_off, n_aux = None, None, 0
        if aux:
            enc = [s.encode("ascii") for s in aux]
            aux_off = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(s) for s in enc], out=aux_off[1:])
            aux_seq = np.frombuffer(b"".join(enc), dtype=np.uint8)
            n_aux = len(enc)
        score = np.zeros(n, dtype=np.int32)
        nmatch = np.zeros(n, dtype=np.int32)
        ncols = np.zeros(n, dtype=np.int32)
        win = np.zeros((n, 16, 4), dtype=np.int32) if want_windows else None
        self._check(self.lib.ngsid_sg_align_paths(self.h, ptr(a), ptr(b), ptr(o), n, ptr(aux_seq), ptr(aux_off), n_aux,
                                                  window, ptr(score), ptr(nmatch), ptr(ncols), ptr(win)))
        return (score, nmatch, ncols, win) if want_windows else (score, nmatch, ncols)

    def poa_consensus(self, job_off, layer_src, layer_begin, layer_len, aux=None, mode=0, match=5,
                      mismatch=-4, gap=-2, trim=False, max_nodes=0, layer_sub=None):
        """K5: one POA consensus per job. Layers index uploaded reads (>= 0) or aux strings (< 0).
        max_nodes > 0 bounds the graph of a job (an error beyond it); 0 = as large as it gets.
        layer_sub = (begin, end) arrays: backbone positions a layer is aligned between (racon's sub-graph
        alignment of layers that do not span their window), -1 = the whole graph.
        Returns (list of consensus strings, node counts)."""
        job_off = as_array(job_off, np.int64)
        src, beg, ln = as_array(layer_src, np.int32), as_array(layer_begin, np.int32), as_array(layer_len, np.int32)
        n_jobs = len(job_off) - 1
        if n_jobs <= 0:
            return [], np.zeros(0, dtype=np.int32)
        aux_seq, aux_off, n_aux = None, None, 0
        if aux:
            enc = [s.encode("ascii") for s in aux]
            aux_off = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(s) for s in enc], out=aux_off[1:])
            aux_seq = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
            n_aux = len(enc)
        p = PoaParams()
        p.mode, p.match, p.mismatch, p.gap, p.trim, p.max_nodes = mode, match, mismatch, gap, 1 if trim else 0, max_nodes
        stride = 4 * int(ln.max() if len(ln) else 1) + 64
        out = np.zeros((n_jobs, stride), dtype=np.uint8)
        out_len = np.zeros(n_jobs, dtype=np.int32)
        nodes = np.zeros(n_jobs, dtype=np.int32)
        if layer_sub is not None:
            sb, se = as_array(layer_sub[0], np.int32), as_array(layer_sub[1], np.int32)
            if len(sb) != len(src) or len(se) != len(src):
                raise ValueError("layer_sub arrays must have one entry per layer")
            self._check(self.lib.ngsid_poa_consensus_sub(self.h, ctypes.byref(p), n_jobs, ptr(job_off), ptr(src), ptr(beg), ptr(ln),
                                                         ptr(sb), ptr(se), ptr(aux_seq), ptr(aux_off), n_aux, ptr(out), stride,
                                                         ptr(out_len), ptr(nodes)))
        else:
            self._check(self.lib.ngsid_poa_consensus(self.h, ctypes.byref(p), n_jobs, ptr(job_off), ptr(src), ptr(beg), ptr(ln),
                                                     ptr(aux_seq), ptr(aux_off), n_aux, ptr(out), stride, ptr(out_len), ptr(nodes)))
        # running totals for whoever reports K5 (bench.py): DP cells, layer steps (= the longest job of a call:
        # a call advances every job one layer per kernel launch), kernel + copy / host graph / whole-call time
        acc = self.__dict__.setdefault("poa_acc", {"calls": 0, "jobs": 0, "cells": 0, "layer_steps": 0, "device_ms": 0.0, "host_ms": 0.0, "call_ms": 0.0})
        acc["calls"] += 1; acc["jobs"] += n_jobs; acc["cells"] += self.poa_cells()
        acc["layer_steps"] += int(np.diff(job_off).max())
        acc["device_ms"] += self.phase_ms(6); acc["host_ms"] += self.phase_ms(7); acc["call_ms"] += self.phase_ms(8)
        return [out[j, :out_len[j]].tobytes().decode("ascii") for j in range(n_jobs)], nodes

    # ---- multi-GPU data plane (NCCL inside the library; a world of one rank without a communicator)
    def nccl_init(self, unique_id, rank, nranks):
        uid = np.frombuffer(unique_id, dtype=np.uint8).copy()
        self._check(self.lib.ngsid_nccl_init(self.h, ptr(uid), rank, nranks))
        self.rank, self.world = rank, nranks

    def nccl_share(self, owner):
        """Use the communicator of `owner` (an Engine on the same GPU that called nccl_init)."""
        self._check(self.lib.ngsid_nccl_share(self.h, owner.h))
        self.rank, self.world = owner.rank, owner.world

    def allgather_bytes(self, data):
        """-> list of bytes objects, one per rank."""
        world = self.world
        send = np.frombuffer(bytes(data), dtype=np.uint8)
        counts = np.zeros(world, dtype=np.int64)
        self._check(self.lib.ngsid_allgather_bytes(self.h, ptr(send), len(send), None, 0, ptr(counts)))     # sizes
        cap = max(1, int(counts.sum()))
        recv = np.zeros(cap, dtype=np.uint8)
        self._check(self.lib.ngsid_allgather_bytes(self.h, ptr(send), len(send), ptr(recv), cap, ptr(counts)))
        out, o = [], 0
        for c in counts:
            out.append(recv[o:o + c].tobytes())
            o += int(c)
        return out

    def allreduce(self, arr, op="sum"):
        """In-place all-reduce of an int32 / int64 numpy array."""
        assert arr.dtype in (np.int32, np.int64) and arr.flags["C_CONTIGUOUS"]
        self._check(self.lib.ngsid_allreduce(self.h, ptr(arr), arr.size, arr.dtype.itemsize, 0 if op == "sum" else 1))
        return arr

    def gather_representatives(self, reps, dst):
        """Collective: representatives (local read indices) of every rank -> `dst` engine (same GPU),
        with their minimizers and quality statistics. Returns counts per rank."""
        reps = as_array(reps, np.int32)
        counts = np.zeros(self.world, dtype=np.int64)
        self._check(self.lib.ngsid_gather_representatives(self.h, ptr(reps), len(reps), dst.h, ptr(counts)))
        dst.n_reads = int(counts.sum())
        dst.offsets, dst.h_seq, dst.h_qual = None, None, None
        dst._q_done = True
        return counts

    def exchange_reads(self, read_idx, dest, tags, dst, expect):
        """Collective all-to-all of reads into `dst` (same GPU); `expect` = reads this rank will
        receive (upper bound). Returns (tags of the received reads, counts per source rank)."""
        read_idx, dest, tags = as_array(read_idx, np.int32), as_array(dest, np.int32), as_array(tags, np.int64)
        out = np.zeros(max(1, int(expect)), dtype=np.int64)
        counts = np.zeros(self.world, dtype=np.int64)
        self._check(self.lib.ngsid_exchange_reads(self.h, ptr(read_idx), ptr(dest), ptr(tags), len(read_idx), dst.h,
                                                  ptr(out), len(out), ptr(counts)))
        n = int(counts.sum())
        dst.n_reads = n
        dst.offsets, dst.h_seq, dst.h_qual = None, None, None
        dst._q_done = False
        return out[:n], counts

    def download_reads(self, begin=0, end=None):
        """(seq u8, qual u8, offsets i64) of reads [begin, end) as they are on the device."""
        end = self.n_reads if end is None else end
        off = np.zeros(end - begin + 1, dtype=np.int64)
        self._check(self.lib.ngsid_download_reads(self.h, begin, end, None, None, ptr(off)))
        seq, qual = np.zeros(int(off[-1]), dtype=np.uint8), np.zeros(int(off[-1]), dtype=np.uint8)
        self._check(self.lib.ngsid_download_reads(self.h, begin, end, ptr(seq), ptr(qual), ptr(off)))
        return seq, qual, off

    def adopt_device_reads(self):
        """Host mirror (offsets, bases, qualities) for a read set that arrived through the data plane."""
        self.h_seq, self.h_qual, self.offsets = self.download_reads()
        self._qcs = None

    def append_revcomp(self):
        self._check(self.lib.ngsid_append_revcomp(self.h))
        n = self.n_reads
        self.n_reads = 2 * n
        if self.offsets is not None:
            self.adopt_device_reads()

    def set_option(self, option, value):
        self._check(self.lib.ngsid_set_option(self.h, option, value))

    def kmer_string(self, code, k):
        """The k-mer string behind a minimizer code, including codes of the exception path (k-mers with a
        base outside ACGT: 1 << 30 | dictionary index)."""
        code = int(code)
        if code & (1 << 30) and not code & (1 << 31):
            buf = ctypes.create_string_buffer(256)
            n = self.lib.ngsid_kmer_string(self.h, code, buf, 256)
            if n < 0:
                self._check(n)
            return buf.value.decode("ascii")
        return decode_kmer(code, k)

    def kmer_strings(self, codes, k):
        """[kmer_string(c, k) for c in codes], the plain 2k-bit codes decoded in one numpy pass."""
        codes = np.asarray(codes, dtype=np.uint32)
        special = (codes >> 30) != 0                       # truncated suffix (bit 31) or dictionary entry (bit 30)
        shifts = np.arange(2 * (k - 1), -1, -2, dtype=np.uint32)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)[(codes[:, None] >> shifts[None, :]) & 3]
        out = [row.tobytes().decode("ascii") for row in letters] if k else [""] * len(codes)
        for i in np.nonzero(special)[0].tolist():
            out[i] = self.kmer_string(codes[i], k)
        return out

    def poa_cells(self):
        """DP cells (graph rows x layer bases, summed) of the last poa_consensus call."""
        return int(self.lib.ngsid_poa_cells(self.h))

    def phase_ms(self, which):
        return float(self.lib.ngsid_phase_ms(self.h, which))

    def launch_count(self):
        return int(self.lib.ngsid_launch_count(self.h))

    def reset_launch_count(self):
        self.lib.ngsid_reset_launch_count(self.h)

    def sync(self):
        self._check(self.lib.ngsid_sync(self.h))


def nccl_unique_id():
    """128-byte NCCL id created by the calling rank (hand it to the other ranks out of band)."""
    buf = np.zeros(128, dtype=np.uint8)
    n = _lib.load().ngsid_nccl_unique_id(ptr(buf), 128)
    if n <= 0:
        raise NgsidError(n, "NCCL is not available (libnccl.so.2 not found; set NGSID_NCCL_LIB)")
    return buf[:n].tobytes()


_ENGINES = {}
_ENGINES_LOCK = threading.Lock()
_TLS = threading.local()


def set_engine_slot(slot):
    """Engine slot of the calling thread: threads that run independent batches on one GPU at the same time
    (modules.parallelize.parallel_clustering) each work on their own context and stream."""
    _TLS.slot = int(slot)


def get_engine(device=0, slot=None):
    """One cached Engine per (device, slot) for the reference-shaped wrappers; slot defaults to the calling
    thread's (0 unless set_engine_slot was called)."""
    if slot is None:
        slot = getattr(_TLS, "slot", 0)
    key = (device, slot)
    with _ENGINES_LOCK:
        e = _ENGINES.get(key)
        if e is None or e.h is None:
            e = _ENGINES[key] = Engine(device)
    return e


def decode_kmer(code, k):
    """2k-bit code (or truncated-suffix code, bit 31 set) -> string."""
    code = int(code)
    if code & (1 << 31):
        body = code & ~(1 << 31)
        t = (body.bit_length() - 1) // 2
        body &= (1 << (2 * t)) - 1
        n = t
    else:
        body, n = code, k
    return "".join("ACGT"[(body >> (2 * (n - 1 - i))) & 3] for i in range(n))
