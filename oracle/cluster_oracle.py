"""
oracle/cluster_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (plain Python, integer k-mer codes) of the reference's clustering hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path (ngspeciesid_b200/) never does.

Pinned against the reference itself: tests/golden/make_golden.py imports /root/reference with the
parasail shim, runs it on the reference's fixtures and on synthetic reads, and stores the
results under tests/golden/; tests/test_oracle.py checks this restatement against those vectors.
The alignment step inside (oracle/sg_align.c) stands in for the un-vendored parasail==1.2.4 and is
PARITY UNPINNED against real parasail (see that file's header).

Reference functions restated (file:line relative to /root/reference):
  homopolymer compression            modules/cluster.py:265
  get_kmer_minimizers                modules/cluster.py:16-39
  quality compression + error rate   modules/cluster.py:273-292
  get_all_hits                       modules/cluster.py:43-62
  get_best_cluster                   modules/cluster.py:67-127
  p_shared_minimizer_empirical       modules/cluster.py:356-368
  get_best_cluster_block_align       modules/cluster.py:172-205
  parasail_block_alignment           modules/cluster.py:130-169
  reads_to_clusters                  modules/cluster.py:207-353
  batch_list / parallel_clustering   modules/parallelize.py:33-81,107-217
  single_clustering                  NGSpeciesID:20-33
"""
import ctypes
import math
import os
from types import SimpleNamespace

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_sg_block_align.restype = ctypes.c_int
        _LIB.oracle_sg_block_align.argtypes = [
            ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return _LIB


def default_args(**over):
    """The hot-path knobs with the reference CLI's defaults (NGSpeciesID:188-245, --ont)."""
    d = dict(k=13, w=20, min_shared=5, mapped_threshold=0.7, aligned_threshold=0.4,
             symmetric_map_align_thresholds=False, min_fraction=0.8, min_prob_no_hits=0.1,
             batch_type="total_nt", nr_cores=1, print_output=10000, outfolder=None)
    d.update(over)
    return SimpleNamespace(**d)


# ---------------------------------------------------------------- per-read primitives
PHRED_P = [min(10 ** (-(c - 33) / 10.0), 0.79433) for c in range(128)]       # cluster.py:233
PHRED_P_UNCAPPED = [10 ** (-(c - 33) / 10.0) for c in range(128)]              # get_sorted_fastq_for_cluster.py:21


def hpol_compress(seq):
    """-> (compressed string, list of run lengths)   [cluster.py:265,279]"""
    if not seq:
        return "", []
    chars, runs = [seq[0]], [1]
    prev = seq[0]
    for ch in seq[1:]:
        if ch == prev:
            runs[-1] += 1
        else:
            chars.append(ch)
            runs.append(1)
            prev = ch
    return "".join(chars), runs


def poisson_mean(qual, table=PHRED_P):
    """sum over distinct quality chars of count*p, with Python >= 3.12's compensated builtin sum.
    The reference iterates a set() (hash order); terms are taken in ascending character order
    here. [cluster.py:290, 185-186]"""
    counts = {}
    for ch in qual:
        counts[ch] = counts.get(ch, 0) + 1
    return sum([counts[ch] * table[ord(ch)] for ch in sorted(counts)])


def compress_quality(qual, runs):
    """Per homopolymer run keep the quality char with the lowest error probability (first such
    char on ties of probability).  [cluster.py:279-286]"""
    out, start = [], 0
    for n in runs:
        best = qual[start]
        bp = PHRED_P[ord(best)]
        for ch in qual[start + 1:start + n]:
            p = PHRED_P[ord(ch)]
            if p < bp:
                best, bp = ch, p
        out.append(best)
        start += n
    return "".join(out)


def minimizers(seqc, k, w):
    """(k,w) lexicographic minimizers of the compressed read: [(kmer, pos)].  [cluster.py:16-39]

    For len >= w this is: leftmost minimum of every window of W = w-k+1 consecutive k-mers,
    reported whenever the position of the minimum changes. Shorter inputs reproduce the
    reference's behaviour of comparing truncated (possibly empty) suffix strings.
    """
    W = w - k + 1
    n_kmers = len(seqc) - k + 1
    first = [seqc[i:i + k] for i in range(W)]
    best = min(first)
    out = [(best, first.index(best))]
    if n_kmers <= W:
        return out
    cur_pos = out[0][1]
    for right in range(W, n_kmers):           # right = index of the k-mer entering the window
        left = right - W + 1
        new = seqc[right:right + k]
        if cur_pos < left:                     # minimum fell out: rescan, leftmost wins
            best, cur_pos = None, -1
            for p in range(left, right + 1):
                km = seqc[p:p + k]
                if best is None or km < best:
                    best, cur_pos = km, p
            out.append((best, cur_pos))
        elif new < best:
            best, cur_pos = new, right
            out.append((best, cur_pos))
    return out


def error_bucket(e):
    """round(e, 2) clamped to [0.01, 0.15]  [cluster.py:356-366]"""
    r = round(e, 2)
    if r > 0.15:
        r = 0.15
    if r < 0.01:
        r = 0.01
    return r


# ---------------------------------------------------------------- pairwise statistics
def mapped_length(hit_idx, hit_pos, n_minimizers, len_c, q_no_share, min_prob_no_hits):
    """total_mapped of cluster.py:99-115 for one (read, representative) pair."""
    def bridged(gap):
        p = 1
        for _ in range(gap):
            p = p * q_no_share
        return not (p < min_prob_no_hits)
    total = 0
    prev_idx, prev_pos = -1, 0
    for j, pos in zip(hit_idx, hit_pos):
        if bridged(j - prev_idx - 1):
            total += pos - prev_pos
        prev_idx, prev_pos = j, pos
    if bridged(n_minimizers - prev_idx - 1):
        total += len_c - prev_pos
    return total


def block_align_ratio(s1, s2, k, match_id, open_pen, ext=1):
    """(alignment_ratio, target_ratio, window_count) [cluster.py:130-169]"""
    b1, b2 = s1.encode(), s2.encode()
    ncols = ctypes.c_int(0)
    score = ctypes.c_int(0)
    cnt = _lib().oracle_sg_block_align(b1, len(b1), b2, len(b2), open_pen, ext, k, match_id,
                                       ctypes.byref(ncols), ctypes.byref(score))
    if cnt < 0:
        raise MemoryError("oracle_sg_block_align")
    return cnt / float(len(s1)), cnt / float(len(s2)), cnt


def gap_open_and_match_id(err_sum, k):
    """cluster.py:189-198"""
    if err_sum <= 0.01:
        go = 5
    elif err_sum <= 0.04:
        go = 4
    elif err_sum <= 0.1:
        go = 3
    else:
        go = 2
    return go, math.floor((1.0 - err_sum) * k)


# ---------------------------------------------------------------- the greedy pass
class Stats(object):
    def __init__(self):
        self.mapped = 0
        self.aln_called = 0
        self.aln_passed = 0
        self.alignments = 0
        self.trace = []      # per processed read: (read id, rep id or -1, 'map'|'align'|'new')


def reads_to_clusters(clusters, representatives, sorted_reads, p_emp_probs, minimizer_database,
                      new_batch_index, args, stats=None):
    """Same contract as the reference's cluster.reads_to_clusters (cluster.py:207-353): mutates
    `clusters`, `representatives`, `minimizer_database`; returns
    {new_batch_index: (clusters, representatives, minimizer_database, new_batch_index)}."""
    k, w = args.k, args.w
    st = stats if stats is not None else Stats()
    batches = [r[1] for r in sorted_reads]
    lowest = max(1, min(batches or [1]))
    moved = {}

    for (rid, prev_batch, acc, seq, qual, score) in sorted_reads:
        if prev_batch == lowest:                                   # cluster.py:243-248
            rec = representatives[rid]
            representatives[rid] = rec[:1] + (new_batch_index,) + rec[2:]
            continue
        seqc, runs = hpol_compress(seq)
        if len(seqc) < k:
            continue
        mins = minimizers(seqc, k, w)
        rec = representatives[rid]
        if len(rec) == 8:
            representatives[rid] = rec[:1] + (new_batch_index,) + rec[2:]
        else:
            qc = compress_quality(qual, runs)
            err = poisson_mean(qc) / float(len(qc))
            representatives[rid] = (rid, new_batch_index, acc, seq, qual, score, err, seqc)
        err_read = representatives[rid][6]

        # hits per representative, in read order                     cluster.py:43-62
        hits = {}
        for j, (km, pos) in enumerate(mins):
            for rep in minimizer_database.get(km, ()):
                if rep == rid:
                    continue
                h = hits.get(rep)
                if h is None:
                    hits[rep] = h = ([], [])
                h[0].append(j)
                h[1].append(pos)

        best_map, best_aln, top = -1, -1, 0
        ranked = None
        if hits:
            ranked = sorted(hits.items(),
                            key=lambda it: (len(it[1][1]), sum(it[1][1]), representatives[it[0]][2]),
                            reverse=True)
            top = len(ranked[0][1][1])
            if top >= args.min_shared:
                for rep, (hidx, hpos) in ranked:                     # cluster.py:84-125
                    n = len(hidx)
                    if n < args.min_fraction * top or n < args.min_shared:
                        break
                    rrec = representatives[rep]
                    q = 1.0 - p_emp_probs[(error_bucket(err_read), error_bucket(rrec[6]))]
                    total = mapped_length(hidx, hpos, len(mins), len(seqc), q, args.min_prob_no_hits)
                    ratio = total / float(len(seqc))
                    if args.symmetric_map_align_thresholds:
                        ratio = min(ratio, total / float(len(rrec[7])))
                    if ratio > args.mapped_threshold:
                        best_map = rep
                        break
        if best_map >= 0:
            st.mapped += 1
        elif top >= args.min_shared:                                 # cluster.py:310-316, 172-205
            st.aln_called += 1
            for rep, (hidx, hpos) in ranked:
                if len(hidx) < top:
                    break
                rrec = representatives[rep]
                e1 = poisson_mean(qual) / float(len(seq))
                e2 = poisson_mean(rrec[4]) / float(len(rrec[3]))
                go, mid = gap_open_and_match_id(e1 + e2, k)
                a_ratio, t_ratio, _ = block_align_ratio(seq, rrec[3], k, mid, go)
                st.alignments += 1
                r = min(a_ratio, t_ratio) if args.symmetric_map_align_thresholds else a_ratio
                if r >= args.aligned_threshold:
                    best_aln = rep
                    st.aln_passed += 1
                    break

        winner = max(best_map, best_aln)
        if winner >= 0:
            moved[rid] = winner
            st.trace.append((rid, winner, "map" if best_map >= 0 else "align"))
        else:
            st.trace.append((rid, -1, "new"))
            for km, _pos in mins:
                s = minimizer_database.get(km)
                if s is None:
                    minimizer_database[km] = s = set()
                s.add(rid)

    for rid, winner in moved.items():                                # cluster.py:338-345
        clusters[winner].extend(clusters[rid])
        del clusters[rid]
        del representatives[rid]
    return {new_batch_index: (clusters, representatives, minimizer_database, new_batch_index)}


def single_clustering(read_array, p_emp_probs, args, stats=None):
    """NGSpeciesID:20-33"""
    clusters = {r[0]: [r[2]] for r in read_array}
    reps = {r[0]: tuple(r) for r in read_array}
    res = reads_to_clusters(clusters, reps, read_array, p_emp_probs, {}, 1, args, stats)
    return res[1][0], res[1][1]


def split_batches(read_array, nr_cores, batch_type="total_nt"):
    """First-round batches of modules/parallelize.py:46-81 (consecutive chunks; may end with an
    empty batch)."""
    if batch_type == "nr_reads":
        size = int(len(read_array) / nr_cores) + 1
        return [read_array[i:i + size] for i in range(0, len(read_array), size)]
    if batch_type == "total_nt":
        weight = lambda r: len(r[3])
    elif batch_type == "read_lengths_squared":
        weight = lambda r: math.pow(len(r[3]), 2)
    else:
        return []
    total = sum(weight(r) for r in read_array)
    limit = int(total / nr_cores) + 1
    out, cur, acc = [], [], 0
    for r in read_array:
        acc += weight(r)
        cur.append(r)
        if acc >= limit:
            out.append(cur)
            cur, acc = [], 0
    out.append(cur)
    return out


def pair_batches(read_array):
    """Merge-round batches (parallelize.py:34-45): consecutive runs of reads whose previous batch
    index is <= 2, <= 4, ...; the list is in score order so a batch pair can appear as more than
    one run only if indices interleave -- the reference yields a new batch each time the index
    exceeds the current bound, which is reproduced."""
    out, cur, bound = [], [], 2
    for r in read_array:
        if r[1] <= bound:
            cur.append(r)
        else:
            out.append(cur)
            bound += 2
            cur = [r]
    out.append(cur)
    return out


def parallel_clustering(read_array, p_emp_probs, args):
    """modules/parallelize.py:107-217 without the process pool (batches are independent, so
    running them one after another gives the same result)."""
    batches = split_batches(read_array, args.nr_cores, args.batch_type)
    num = args.nr_cores
    cl = [{r[0]: [r[2]] for r in b} for b in batches]
    rp = [{r[0]: tuple(r) for r in b} for b in batches]
    db = [{} for _ in batches]
    while True:
        if len(batches) == 1:
            res = reads_to_clusters(cl[0], rp[0], batches[0], p_emp_probs, db[0], 1, args)
            return res[1][0], res[1][1]
        all_cl, all_rp, all_db = {}, {}, {}
        for i in range(len(batches)):
            res = reads_to_clusters(cl[i], rp[i], batches[i], p_emp_probs, db[i], i + 1, args)
            c, r, d, bi = res[i + 1]
            all_cl.update(c)
            all_rp.update(r)
            all_db[bi] = d
        read_array = [(v[0], v[1], v[2], v[3], v[4], v[5]) for _, v in
                      sorted(all_rp.items(), key=lambda x: x[1][5], reverse=True)]
        if num == 1:
            return all_cl, all_rp
        batches = pair_batches(read_array)
        num = len(batches)
        cl, rp, db = [], [], []
        for b in batches:
            low = min(r[1] for r in b)
            cl.append({r[0]: all_cl[r[0]] for r in b})
            rp.append({r[0]: all_rp[r[0]] for r in b})
            db.append(all_db[low])


# ---------------------------------------------------------------- pipeline glue used by tests / bench
def expected_error_free_kmers_score(qual, k):
    """The sort key of modules/get_sorted_fastq_for_cluster.py:23-33,150-152 (row f.1, used here
    only to order synthetic reads the way the reference's sort stage would)."""
    pe = [PHRED_P[ord(c)] for c in qual]
    cur = 1
    for p in pe[:k]:
        cur = cur * (1.0 - p)
    total = cur
    for i in range(k, len(pe)):
        cur *= ((1.0 - pe[i]) / (1.0 - pe[i - k]))
        total += cur
    n = len(qual) - k + 1
    exp_err = n - total
    p_no_err = 1.0 - exp_err / float(n)
    return p_no_err * n


def sort_stage(records, k, q_threshold=7.0):
    """records: iterable of (acc, seq, qual) -> list of (acc_with_score, seq, qual, score) in the
    order the reference's sort stage emits (get_sorted_fastq_for_cluster.py:124-155, 174-177)."""
    out = []
    for acc, seq, qual in records:
        seqc, _ = hpol_compress(seq)
        if len(seq) < 2 * k or len(seqc) < k:
            continue
        e = poisson_mean(qual, PHRED_P_UNCAPPED) / float(len(qual))
        if 10 * -math.log(e, 10) <= q_threshold:
            continue
        s = expected_error_free_kmers_score(qual, k)
        out.append((acc + "_{0}".format(s), seq, qual, s))
    out.sort(key=lambda x: x[3], reverse=True)
    return out


def read_array_from_sorted(sorted_records):
    """NGSpeciesID:54-58: (i, 0, acc, seq, qual, float(acc.split('_')[-1]))"""
    return [(i, 0, acc, seq, qual, float(acc.split("_")[-1]))
            for i, (acc, seq, qual, _s) in enumerate(sorted_records)]


def load_p_emp(table_rows, k, w):
    """NGSpeciesID:72-77"""
    out = {}
    for kk, ww, p, e1, e2 in table_rows:
        if int(kk) == k and abs(int(ww) - w) <= 2:
            out[(float(e1), float(e2))] = float(p)
            out[(float(e2), float(e1))] = float(p)
    return out


def output_order(clusters, representatives):
    """Cluster/member order of the reference's TSV writer (NGSpeciesID:104-114): clusters by
    (size, representative score) descending, members by their score suffix descending; both sorts
    are stable. Returns [(rep_id, [accessions...])]."""
    out = []
    for c_id, accs in sorted(clusters.items(), key=lambda x: (len(x[1]), representatives[x[0]][5]),
                             reverse=True):
        out.append((c_id, sorted(accs, key=lambda a: float(a.split("_")[-1]), reverse=True)))
    return out
