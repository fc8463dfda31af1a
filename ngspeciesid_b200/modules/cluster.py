"""
Drop-in for the reference's modules/cluster.py hot path (ksahlin/NGSpeciesID v0.3.1), computed on
the GPU through libngsid.so. Function names, arguments, mutation of the caller's dicts and return
values follow the reference (file:line cited per function); there is no CPU fallback.
"""
import collections
import itertools
import logging
import math
import operator

import numpy as np

from .. import engine as _engine
from ..build import load_hostpack

_hostpack = load_hostpack()          # csrc/hostpack.c: tuples of str -> byte arrays (host glue)


def _hpol(seq):
    return "".join(ch for ch, _ in itertools.groupby(seq))


def get_kmer_minimizers(seq, k_size, w_size):
    """Reference: modules/cluster.py:16-39. `seq` is the (already homopolymer-compressed) string;
    returns [(kmer, position)]. The kernel compresses its input itself, which is the identity on a
    compressed string; a string that still contains homopolymers is rejected, because the reference
    would not compress it."""
    if _hpol(seq) != seq:
        raise ValueError("get_kmer_minimizers expects a homopolymer-compressed sequence "
                         "(the reference only ever passes seq_hpol_comp, modules/cluster.py:269)")
    eng = _engine.get_engine()
    eng.upload_records([(seq, "I" * len(seq))])
    eng.minimizers(k_size, w_size)
    _lc, _cnt, kmer, pos = eng.get_minimizers(0, 1)
    return [(eng.kmer_string(c, k_size), int(p)) for c, p in zip(kmer, pos)]


def p_shared_minimizer_empirical(error_rate_read, error_rate_center, p_emp_probs):
    """Reference: modules/cluster.py:356-368 (host-side lookup; the device uses the bucketed form)."""
    def bucket(e):
        e = round(e, 2)
        return 0.15 if e > 0.15 else (0.01 if e < 0.01 else e)
    return p_emp_probs[(bucket(error_rate_read), bucket(error_rate_center))]


def parasail_block_alignment(s1, s2, k, match_id, match_score=2, mismatch_penalty=-2, opening_penalty=5, gap_ext=1):
    """Reference: modules/cluster.py:130-169. Returns (s1, s2, (None, None, alignment_ratio,
    target_alignment_ratio)): the gapped strings are not materialised (the kernel computes the
    block statistic during the DP); every caller in the reference only uses the two ratios."""
    if (match_score, mismatch_penalty, gap_ext) != (2, -2, 1):
        raise NotImplementedError("the kernel implements the reference's fixed scoring 2/-2/ext 1")
    eng = _engine.get_engine()
    eng.upload_records([(s1, "I" * len(s1)), (s2, "I" * len(s2))])
    cnt = eng.sg_block_align([0], [1], [opening_penalty], [match_id], k)
    c = int(cnt[0])
    return (s1, s2, (None, None, c / float(len(s1)), c / float(len(s2))))


def reads_to_clusters(clusters, representatives, sorted_reads, p_emp_probs, minimizer_database, new_batch_index, args):
    """Reference: modules/cluster.py:207-353. Same contract: iterates `sorted_reads` in order,
    mutates `clusters`, `representatives` (values become 8-tuples) and `minimizer_database`
    ({kmer: set(ids)}), returns {new_batch_index: (clusters, representatives,
    minimizer_database, new_batch_index)}."""
    k, w = args.k, args.w
    prev_b = list(map(operator.itemgetter(1), sorted_reads))
    lowest = max(1, min(prev_b or [1]))

    # reads whose batch index equals the lowest one are the table's own representatives:
    # only their batch index changes (modules/cluster.py:243-248)
    if lowest in prev_b:
        todo = []
        for rec in sorted_reads:
            rid = rec[0]
            if rec[1] == lowest:
                t = representatives[rid]
                representatives[rid] = t[:1] + (new_batch_index,) + t[2:]
            else:
                todo.append(rec)
    else:
        todo = sorted_reads if isinstance(sorted_reads, list) else list(sorted_reads)

    # representatives already in the caller's table
    init_ids = set()
    for ids in minimizer_database.values():
        init_ids.update(ids)
    init_ids = sorted(init_ids)

    if todo:
        # one upload for the table's representatives (first) and the reads to cluster. Everything per read
        # below runs inside C or numpy: the two str fields of the records go straight into byte arrays
        # (_hostpack.pack_fields), Python objects are touched for survivors and in bulk for the moves.
        n_init, n_todo = len(init_ids), len(todo)
        recs = [representatives[rid] for rid in init_ids] + todo if n_init else todo
        eng = _engine.get_engine(getattr(args, "device", 0))
        total = _hostpack.measure_fields(recs, 3)
        h_seq, h_qual = eng.pinned_pair(total)                       # page-locked, owned by the engine
        offs = np.frombuffer(_hostpack.pack_fields_into(recs, 3, 4, h_seq.ctypes.data, h_qual.ctypes.data, total),
                             dtype=np.int64)
        ids = np.fromiter(map(operator.itemgetter(0), recs), np.int64, len(recs))
        accs = list(map(operator.itemgetter(2), recs))
        eng.upload(h_seq[:total], h_qual[:total], offs)
        eng.minimizers(k, w)
        eng.quality_stats()
        max_gap = _engine.max_gap_table(p_emp_probs, args.min_prob_no_hits)
        assign, via, stats = eng.cluster(k, w, max_gap, np.arange(n_init, n_init + n_todo, dtype=np.int32),
                                         _engine.accession_ranks(accs), init_reps=np.arange(n_init, dtype=np.int32),
                                         min_shared=args.min_shared, min_fraction=args.min_fraction,
                                         mapped_threshold=args.mapped_threshold,
                                         aligned_threshold=args.aligned_threshold,
                                         symmetric=bool(args.symmetric_map_align_thresholds))
        new_local = np.nonzero(assign == -1)[0]
        # survivors: 8-tuples (error rate of the compressed qualities, compressed sequence) and their
        # minimizers into the caller's table (cluster.py:292, 328-334)
        for i in new_local.tolist():
            li = n_init + i
            rid, _b, acc, seq, qual, score = todo[i][:6]
            t = representatives[rid]
            if len(t) == 8:
                representatives[rid] = t[:1] + (new_batch_index,) + t[2:]
            else:
                err_c, _eu, _bk = eng.get_quality_stats(li, li + 1)
                representatives[rid] = (rid, new_batch_index, acc, seq, qual, score, float(err_c[0]), _hpol(seq))
            _lc, _cnt, kmer, _pos = eng.get_minimizers(li, li + 1)
            for m in eng.kmer_strings(kmer, k):
                s = minimizer_database.get(m)
                if s is None:
                    minimizer_database[m] = s = set()
                s.add(rid)
        # assigned reads join their representative in processing order (cluster.py:338-345): per winner, the
        # accession lists of its reads are concatenated in that order; the loops run inside C (map / chain)
        moved = np.nonzero(assign >= 0)[0]
        if len(moved):
            win_local = assign[moved]
            by_winner = np.argsort(win_local, kind="stable")             # stable: processing order within a winner
            moved_ids = ids[n_init + moved[by_winner]]
            win_sorted = win_local[by_winner]
            cuts = np.nonzero(np.diff(win_sorted))[0] + 1
            starts = [0] + cuts.tolist() + [len(moved)]
            moved_list = moved_ids.tolist()
            pop_cluster = clusters.pop
            for g in range(len(starts) - 1):
                group = moved_list[starts[g]:starts[g + 1]]
                winner = int(ids[win_sorted[starts[g]]])
                clusters[winner].extend(itertools.chain.from_iterable(map(pop_cluster, group)))
            collections.deque(map(representatives.__delitem__, moved_list), maxlen=0)
        logging.debug("Total number of reads iterated through:{0}".format(len(sorted_reads)))
        logging.debug("Passed mapping criteria:{0}".format(stats["n_mapped"]))
        logging.debug("Passed alignment criteria in this process:{0}".format(stats["n_aln_passed"]))
        logging.debug("Total calls to alignment module in this process:{0}".format(stats["n_alignments"]))
    return {new_batch_index: (clusters, representatives, minimizer_database, new_batch_index)}
