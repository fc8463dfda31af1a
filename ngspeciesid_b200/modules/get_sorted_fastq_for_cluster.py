"""The sort stage in front of the clustering path, with the reference's names
(modules/get_sorted_fastq_for_cluster.py): reads are filtered (len >= 2k, homopolymer-compressed
length >= k, mean quality above the threshold), scored with the expected number of error-free
k-mers and written to sorted.fastq in descending score order with the score appended to the name.

The per-read arithmetic (score and mean error probability) runs on the GPU through
ngsid_sort_scores, bit-identical to the reference's doubles; filtering, the stable sort and the file
are host work, as in the reference. Reads must be ACGT (limit of this build)."""
import logging
import math
import os
from time import time

import numpy as np

from .. import engine as _engine
from . import help_functions


def _threshold_error_rate(q_threshold):
    """Smallest double e with 10 * -math.log(e, 10) <= q_threshold (the reference's skip rule,
    get_sorted_fastq_for_cluster.py:147, is monotone in e), found with math.log itself so that the
    vectorised comparison e >= e* decides every read exactly like the reference."""
    def skipped(e):
        return 10 * -math.log(e, 10) <= q_threshold
    lo, hi = 1e-300, 1e300          # not skipped / skipped
    if skipped(lo):
        return 0.0
    if not skipped(hi):
        return float("inf")
    while True:
        mid = math.sqrt(lo) * math.sqrt(hi) if hi / lo > 4 else lo + (hi - lo) / 2
        if mid <= lo or mid >= hi:
            break
        if skipped(mid):
            hi = mid
        else:
            lo = mid
    e = hi
    while True:                      # walk down to the first double that is skipped
        prev = np.nextafter(e, 0.0)
        if prev > 0 and skipped(float(prev)):
            e = float(prev)
        else:
            return e


def _score_uploaded(eng, k, q_threshold):
    """Scores, filter and stable order of the reads currently uploaded to `eng` -> (order, kept
    indices in file order, score, err): reference :124-155 / :41-66 with the arithmetic on the GPU."""
    score, err = eng.sort_scores(k)
    seqb = eng.h_seq
    off = eng.offsets
    lens = np.diff(off)
    # homopolymer-compressed length: 1 + number of positions whose base differs from the previous one
    diff = np.ones(len(seqb), dtype=np.int64)
    if len(seqb) > 1:
        diff[1:] = seqb[1:] != seqb[:-1]
    starts = off[:-1][lens > 0]
    diff[starts] = 1
    csum = np.concatenate([[0], np.cumsum(diff)])
    lenc = csum[off[1:]] - csum[off[:-1]]
    keep = ~((lens < 2 * k) | (lenc < k))
    e_star = _threshold_error_rate(q_threshold)
    keep &= ~(err >= e_star)
    idx = np.nonzero(keep)[0]
    order = idx[np.argsort(-score[idx], kind="stable")]
    return order, idx, score, err


def score_records(records, k, q_threshold, eng=None):
    """records: [(acc, seq, qual)] in file order -> (read_array, error_rates) exactly as
    fastq_single_core / fastq_parallel build them (reference :124-155, :41-66): read_array =
    [(acc, seq, qual, score)] sorted by score descending (stable), error_rates of the kept reads
    in file order."""
    eng = eng or _engine.get_engine()
    records = list(records)
    if not records:
        return [], []
    eng.upload_records([(s, q) for _a, s, q in records])
    order, idx, score, err = _score_uploaded(eng, k, q_threshold)
    error_rates = [float(err[i]) for i in idx]
    read_array = [(records[i][0], records[i][1], records[i][2], float(score[i])) for i in order]
    return read_array, error_rates


def score_fastq_arrays(fa, k, q_threshold, eng=None):
    """The same for a file parsed by help_functions.read_fastq_arrays (ngsid_fastq_parse): the
    concatenated arrays go to the GPU as they are and Python strings are only made for the reads
    that pass the filter."""
    eng = eng or _engine.get_engine()
    if len(fa) == 0:
        return [], []
    if not fa.has_qual.all() or (np.diff(fa.seq_off) != np.diff(fa.qual_off)).any():
        # FASTA records or qualities longer than the read: the generic record path decides
        return score_records([(n, s, q) for n, (s, q) in fa.records()], k, q_threshold, eng)
    eng.upload(fa.seq, fa.qual, fa.seq_off)
    order, idx, score, err = _score_uploaded(eng, k, q_threshold)
    error_rates = [float(err[i]) for i in idx]
    read_array = []
    for i in order:
        name, (s, q) = fa.record(i)
        read_array.append((name, s, q, float(score[i])))
    return read_array, error_rates


def fastq_single_core(args):
    return score_fastq_arrays(help_functions.read_fastq_arrays(args.fastq), args.k, args.quality_threshold)


def fastq_parallel(args):
    # the reference splits the reads over a process pool and concatenates the batches in order
    # before the stable sort; one GPU pass gives the same list
    return fastq_single_core(args)


def main(args):
    start = time()
    logfile = open(os.path.join(args.outfolder, "logfile.txt"), 'w')
    if os.path.isfile(args.outfile) and args.use_old_sorted_file:
        logging.warning("Using already existing sorted file in specified directory, in not intended, specify different outfolder or delete the current file.")
        return args.outfile
    elif args.fastq:
        read_array, error_rates = fastq_parallel(args) if args.nr_cores > 1 else fastq_single_core(args)

    reads_sorted_outfile = open(args.outfile, "w")
    for i, (acc, seq, qual, score) in enumerate(read_array):
        reads_sorted_outfile.write("@{0}\n{1}\n+\n{2}\n".format(acc + "_{0}".format(score), seq, qual))
    reads_sorted_outfile.close()
    logging.debug(f"{len(read_array)} reads passed quality critera (avg phred Q val over {args.quality_threshold} and length > 2*k) and will be clustered.")
    error_rates.sort()
    min_e = error_rates[0]
    max_e = error_rates[-1]
    median_e = error_rates[int(len(error_rates)/2)]
    mean_e = sum(error_rates)/len(error_rates)
    logfile.write("Lowest read error rate:{0}\n".format(min_e))
    logfile.write("Highest read error rate:{0}\n".format(max_e))
    logfile.write("Median read error rate:{0}\n".format(median_e))
    logfile.write("Mean read error rate:{0}\n".format(mean_e))
    logfile.write("\n")
    logfile.close()
    logging.debug("Sorted all reads in {0} seconds.".format(time() - start))
    return reads_sorted_outfile.name
