"""
Drop-in for the reference's modules/consensus.py hot path (ksahlin/NGSpeciesID v0.3.1): the
subprocess calls to spoa / minimap2 / racon are replaced by in-process CUDA kernels reached through
libngsid.so (K5 partial-order alignment, K4 read-to-draft alignment). Function names, arguments,
return values and the files written follow the reference (file:line cited per function).
Medaka polishing (run_medaka, modules/consensus.py:94-104) is out of scope and not provided.
"""
import glob
import logging
import os
import shutil

import numpy as np

from .. import engine as _engine
from . import help_functions

WINDOW = 500                     # racon's default window length
_RC = bytes.maketrans(b"ACGT", b"TGCA")


def reverse_complement(string):
    """Reference: modules/consensus.py:75-81 (ACGT alphabet of this build)."""
    return string.encode().translate(_RC)[::-1].decode()


# ------------------------------------------------------------------------------------- array level
def draft_consensus_batch(eng, read_lists, max_nodes=0):
    """One spoa-equivalent consensus per list of uploaded-read indices (file order), all lists in
    one launch (one thread block per list)."""
    job_off = np.zeros(len(read_lists) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in read_lists], out=job_off[1:])
    src = np.concatenate([np.asarray(r, dtype=np.int32) for r in read_lists]) if read_lists else np.zeros(0, np.int32)
    lens = np.diff(eng.offsets)[src].astype(np.int32)
    cons, nodes = eng.poa_consensus(job_off, src, np.zeros(len(src), dtype=np.int32), lens,
                                    mode=0, match=5, mismatch=-4, gap=-2, trim=False, max_nodes=max_nodes)
    return cons, nodes


def _quality_prefix(eng, reads):
    """Prefix sums of (quality - 33) over the given uploaded reads only (a few thousand reads polish
    a cluster; a prefix over the whole upload would cost more than the round itself).
    Returns (prefix array, start of each read of `reads` in it)."""
    reads = np.asarray(reads, dtype=np.int64)
    uniq, inv = np.unique(reads, return_inverse=True)
    lens = (eng.offsets[uniq + 1] - eng.offsets[uniq]).astype(np.int64)
    starts = np.zeros(len(uniq) + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    # gather the quality bytes of the involved reads into one contiguous array
    idx = np.repeat(eng.offsets[uniq].astype(np.int64) - starts[:-1], lens) + np.arange(starts[-1], dtype=np.int64)
    q = eng.h_qual[idx].astype(np.int64) - 33
    qcs = np.zeros(starts[-1] + 1, dtype=np.int64)
    np.cumsum(q, out=qcs[1:])
    return qcs, starts[:-1][inv]


def polish_round_batch(eng, targets, read_lists, rc_lists=None, max_nodes=0):
    """One racon-equivalent round for every target at once. read_lists[t]: uploaded-read indices
    polishing target t; rc_lists[t] (optional): for each of them the index of its uploaded reverse
    complement, or -1 -- when given, a read is used in the orientation with the higher score."""
    n_t = len(targets)
    A, B = [], []
    for t, rl in enumerate(read_lists):
        A.extend(rl)
        B.extend([-t - 1] * len(rl))
    A = np.asarray(A, dtype=np.int32)
    B = np.asarray(B, dtype=np.int32)
    if len(A) == 0:
        return list(targets)
    score, _m, _c, win = eng.sg_align_paths(A, B, np.full(len(A), 3, dtype=np.int32), aux=targets,
                                            window=WINDOW, want_windows=True)
    if rc_lists is not None:
        R = np.concatenate([np.asarray(r, dtype=np.int32) for r in rc_lists])
        has = R >= 0
        if has.any():
            s2, _m2, _c2, w2 = eng.sg_align_paths(R[has], B[has], np.full(int(has.sum()), 3, dtype=np.int32),
                                                  aux=targets, window=WINDOW, want_windows=True)
            better = np.zeros(len(A), dtype=bool)
            better[np.nonzero(has)[0]] = s2 > score[has]
            idx = np.nonzero(has)[0]
            take = s2 > score[has]
            win[idx[take]] = w2[take]
            A = A.copy()
            A[idx[take]] = R[has][take]
    tlen = np.array([len(t) for t in targets], dtype=np.int64)
    tl = tlen[-B - 1]
    qcs, roff = _quality_prefix(eng, A)
    jobs, job_keys = [], []
    layer_rows = []
    n_pairs = len(A)
    for w in range(16):
        q0, q1, t0, t1 = win[:, w, 0], win[:, w, 1], win[:, w, 2], win[:, w, 3]
        ws = w * WINDOW
        wlen = np.minimum(WINDOW, tl - ws)
        ok = (q0 >= 0) & (wlen > 0)
        if not ok.any():
            continue
        seg = (q1 - q0).astype(np.int64)
        ok &= ~(seg < 0.02 * wlen)
        segc = np.maximum(seg, 1)
        mean_q = (qcs[roff + np.maximum(q1, 0)] - qcs[roff + np.maximum(q0, 0)]) / segc.astype(np.float64)
        ok &= ~(mean_q < 10.0)
        # a layer that spans the window to within 1 % at both ends is aligned to the whole window graph, a
        # shorter one to the sub-graph between its first and last backbone position (racon window.cpp)
        off = 0.01 * wlen
        b, e = t0 - ws, t1 - ws - 1
        spans = (b < off) & (e > wlen - off)
        rows = np.nonzero(ok)[0]
        if len(rows):
            sb = np.where(spans[rows], -1, b[rows]).astype(np.int64)
            se = np.where(spans[rows], -1, e[rows]).astype(np.int64)
            layer_rows.append(np.stack([(-B[rows] - 1).astype(np.int64), np.full(len(rows), w, dtype=np.int64), b[rows].astype(np.int64),
                                        rows.astype(np.int64), A[rows].astype(np.int64), q0[rows].astype(np.int64), seg[rows],
                                        sb, se], axis=1))
    out = []
    if not layer_rows:
        return list(targets)
    L = np.concatenate(layer_rows)
    # per (target, window): layers sorted by start, ties in read order
    order = np.lexsort((L[:, 3], L[:, 2], L[:, 1], L[:, 0]))
    L = L[order]
    key = L[:, 0] * 16 + L[:, 1]
    uniq, start, count = np.unique(key, return_index=True, return_counts=True)
    job_off, src, beg, ln, job_key = [0], [], [], [], []
    sub_b, sub_e = [], []
    for u, s0, c in zip(uniq, start, count):
        if c + 1 < 3:
            continue
        t, w = int(u // 16), int(u % 16)
        ws = w * WINDOW
        wlen = min(WINDOW, len(targets[t]) - ws)
        src.append(-t - 1); beg.append(ws); ln.append(wlen); sub_b.append(-1); sub_e.append(-1)
        rows = L[s0:s0 + c]
        src.extend(rows[:, 4].tolist()); beg.extend(rows[:, 5].tolist()); ln.extend(rows[:, 6].tolist())
        sub_b.extend(rows[:, 7].tolist()); sub_e.extend(rows[:, 8].tolist())
        job_off.append(len(src))
        job_key.append((t, w))
    cons = {}
    if job_key:
        res, _nodes = eng.poa_consensus(job_off, src, beg, ln, aux=targets, mode=1, match=3, mismatch=-5, gap=-4,
                                        trim=True, max_nodes=max_nodes, layer_sub=(sub_b, sub_e))
        cons = dict(zip(job_key, res))
    for t, tgt in enumerate(targets):
        parts = []
        for w in range((len(tgt) + WINDOW - 1) // WINDOW):
            parts.append(cons.get((t, w), tgt[w * WINDOW:(w + 1) * WINDOW]))
        out.append("".join(parts))
    return out


def polish_batch(eng, targets, read_lists, iters, rc_lists=None, max_nodes=0):
    for _ in range(iters):
        targets = polish_round_batch(eng, targets, read_lists, rc_lists, max_nodes)
    return targets


# ------------------------------------------------------------------------------------- reference-shaped
def _read_fastq(path):
    with open(path, "r") as f:
        return [(acc, seq, qual) for acc, (seq, qual) in help_functions.readfq(f)]


def run_spoa(reads, spoa_out_file, spoa_path):
    """Reference: modules/consensus.py:83-92. `reads` is a FASTQ path; the consensus is also left
    in `spoa_out_file` as a 2-line FASTA (line 2 = sequence), `spoa_path` is ignored."""
    recs = _read_fastq(reads)
    eng = _engine.get_engine()
    eng.upload_records([(s, q) for _a, s, q in recs])
    cons, _n = draft_consensus_batch(eng, [list(range(len(recs)))])
    with open(spoa_out_file, "w") as f:
        f.write(">Consensus LN:i:{0} RC:i:{1} XC:f:1.000000\n{2}\n".format(len(cons[0]), len(recs), cons[0]))
    return cons[0]


def _racon_batch(jobs, racon_iter):
    """jobs: [(reads fastq, centre fasta, outfolder)]. All centres are polished together: one
    upload, and per iteration one alignment launch and one window-POA launch over every centre.
    The jobs are independent, so every centre gets what a call of its own would give it."""
    eng = _engine.get_engine()
    flat, lists, rc_lists, names, targets = [], [], [], [], []
    for reads_to_center, center_file, _out in jobs:
        recs = _read_fastq(reads_to_center)
        with open(center_file) as f:
            lines = f.readlines()
        names.append(lines[0].strip()[1:])
        targets.append(lines[1].strip())
        fwd = [(s, q) for _a, s, q in recs]
        lists.append(list(range(len(flat), len(flat) + len(fwd))))
        flat.extend(fwd)
    n_fwd = len(flat)
    for li in lists:                              # reverse complements behind all forward reads
        rc_lists.append([n_fwd + i for i in li])
    flat.extend([(reverse_complement(s), q[::-1]) for s, q in flat[:n_fwd]])
    eng.upload_records(flat)
    logs = [open(os.path.join(out, "stdout.txt"), "w") for _r, _c, out in jobs]
    try:
        for i in range(racon_iter):
            targets = polish_round_batch(eng, targets, lists, rc_lists)
            for (_r, _c, out), name, target, li, log in zip(jobs, names, targets, lists, logs):
                open(os.path.join(out, "read_alignments_it_{0}.paf".format(i)), "w").close()
                with open(os.path.join(out, "racon_polished_it_{0}.fasta".format(i)), "w") as f:
                    f.write(">{0} LN:i:{1} RC:i:{2} XC:f:1.000000\n{3}\n".format(name, len(target), len(li), target))
                log.write("iteration {0}: {1} bases\n".format(i, len(target)))
        for (_r, _c, out), name, target, li in zip(jobs, names, targets, lists):
            with open(os.path.join(out, "consensus.fasta"), "w") as f:
                f.write(">{0} LN:i:{1} RC:i:{2} XC:f:1.000000\n{3}\n".format(name, len(target), len(li), target))
    finally:
        for log in logs:
            log.close()
    return targets


def run_racon(reads_to_center, center_file, outfolder, cores, racon_iter):
    """Reference: modules/consensus.py:107-126. Writes racon_polished_it_<i>.fasta per iteration and
    consensus.fasta (2-line FASTA) into `outfolder`; reads of either strand are used in the
    orientation that aligns better to the centre (what minimap2 decides in the reference)."""
    _racon_batch([(reads_to_center, center_file, outfolder)], racon_iter)


def run_medaka(reads_to_center, center_file, outfolder, cores, medaka_model, outfastq=False):
    """Reference: modules/consensus.py:94-104. medaka is a neural polisher outside this
    implementation: like the reference this shells out to the
    `medaka_consensus` executable, which has to be installed."""
    import subprocess
    with open(os.path.join(outfolder, "stdout.txt"), "w") as out, open(os.path.join(outfolder, "stderr.txt"), "w") as err:
        cmd = ["medaka_consensus", "-i", reads_to_center, "-d", center_file, "-o", outfolder, "-t", cores]
        if medaka_model:
            cmd += ["-m", medaka_model]
        if outfastq:
            cmd += ["-q"]
        subprocess.check_call(cmd, stdout=out, stderr=err)


def _pair_identities(eng, firsts, seconds):
    """highest_aln_identity for many pairs in ONE K4 launch: identity (equal columns / all columns of
    the semi-global alignment, gap open 3 / extend 1) of firsts[i] against seconds[i] and against its
    reverse complement; the larger of the two."""
    n = len(firsts)
    if n == 0:
        return np.zeros(0)
    if eng.n_reads == 0:
        eng.upload_records([("ACGT", "5555")])
    aux = list(firsts) + list(seconds) + [reverse_complement(x) for x in seconds]
    a = np.concatenate([-(np.arange(n) + 1), -(np.arange(n) + 1)]).astype(np.int32)
    b = np.concatenate([-(np.arange(n) + n + 1), -(np.arange(n) + 2 * n + 1)]).astype(np.int32)
    _score, match, cols = eng.sg_align_paths(a, b, np.full(2 * n, 3, dtype=np.int32), aux=aux)
    ident = match / np.maximum(cols, 1).astype(np.float64)
    return ident[:n], ident[n:]


def highest_aln_identity(seq, seq2):
    """Reference: modules/consensus.py:129-145: identity (equal columns / all columns) of the
    semi-global alignment (open 3, extend 1) in the better of the two orientations."""
    fw, rc = _pair_identities(_engine.get_engine(), [seq], [seq2])
    logging.debug("Rec comp orientation identity %: {0}".format(rc[0]))
    logging.debug("Forward orientation identity %: {0}".format(fw[0]))
    return max([float(fw[0]), float(rc[0])])


def detect_reverse_complements(centers, rc_identity_threshold):
    """Merges centres that are (reverse-complement) copies of an earlier, larger centre; behaviour of
    the reference's modules/consensus.py:148-183: a centre that has not been absorbed looks at EVERY
    later centre (also at ones an earlier centre absorbed already) and absorbs those whose identity
    reaches the threshold; the result rows are [merged read count, c_id, sequence, [read files]].
    All pair identities come from one alignment launch instead of one call per pair."""
    from ..multi_gpu import merge_reverse_complements
    n = len(centers)
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    ident = np.zeros((n, n))
    if pairs:
        fw, rc = _pair_identities(_engine.get_engine(), [centers[i][2] for i, _j in pairs], [centers[j][2] for _i, j in pairs])
        for (i, j), x, y in zip(pairs, fw, rc):
            ident[i, j] = max(x, y)
    files = [list(c[3]) if isinstance(c[3], list) else [c[3]] for c in centers]
    merged = []
    for i, total, members in merge_reverse_complements([c[0] for c in centers], ident, rc_identity_threshold):
        merged.append([total, centers[i][1], centers[i][2], [f for m in members for f in files[m]]])
    logging.debug("{0} consensus formed.".format(len(merged)))
    return merged


def form_draft_consensus(clusters, representatives, sorted_reads_fastq_file, work_dir, abundance_cutoff, args):
    """Reference: modules/consensus.py:249-278 -> [[n_reads, c_id, consensus, reads_path], ...].
    The per-cluster FASTQ files are written as in the reference; the drafts of all clusters are
    computed in one batched launch."""
    reads = {acc: (seq, qual) for acc, seq, qual in _read_fastq(sorted_reads_fastq_file)}
    centers, todo = [], []
    singletons, discarded = 0, []
    for c_id, all_read_acc in sorted(clusters.items(), key=lambda x: (len(x[1]), representatives[x[0]][5]), reverse=True):
        n = len(all_read_acc)
        if n >= abundance_cutoff:
            path = os.path.join(work_dir, "reads_c_id_{0}.fq".format(c_id))
            used = []
            with open(path, "w") as f:
                for i, acc in enumerate(all_read_acc):
                    if args.max_seqs_for_consensus >= 0 and i >= args.max_seqs_for_consensus:
                        break
                    seq, qual = reads[acc]
                    f.write("@{0}\n{1}\n{2}\n{3}\n".format(acc, seq, "+", qual))
                    used.append((seq, qual))
            todo.append((n, c_id, path, used))
        elif n == 1:
            singletons += 1
        elif n > 1:
            discarded.append(n)
    if todo:
        eng = _engine.get_engine(getattr(args, "device", 0))
        flat, lists = [], []
        for _n, _c, _p, used in todo:
            lists.append(list(range(len(flat), len(flat) + len(used))))
            flat.extend(used)
        eng.upload_records(flat)
        cons, _nodes = draft_consensus_batch(eng, lists)
        for (n, c_id, path, _u), c in zip(todo, cons):
            with open(os.path.join(work_dir, "spoa_tmp.fa"), "w") as f:
                f.write(">Consensus LN:i:{0}\n{1}\n".format(len(c), c))
            centers.append([n, c_id, c, path])
    logging.debug("{0} singletons were discarded".format(singletons))
    logging.debug("{0} clusters were discarded due to not passing the abundance_cutoff: a total of {1} reads "
                  "were discarded. Highest abundance among them: {2} reads.".format(len(discarded), sum(discarded), max(discarded or [0])))
    return centers


def _clear_previous_outputs(outfolder, prefix):
    for folder in glob.glob(os.path.join(outfolder, prefix + "*")):
        shutil.rmtree(folder)
    for stale in glob.glob(os.path.join(outfolder, "consensus_reference_*")):
        os.remove(stale)


def _write_centre_inputs(outfolder, c_id, n_reads, centre, read_files):
    """consensus_reference_<c_id>.fasta and reads_to_consensus_<c_id>.fastq of one centre
    (modules/consensus.py:201-215 of the reference: one record per accession and file, the accession
    cut at its first blank). Returns both paths."""
    fasta = os.path.join(outfolder, "consensus_reference_{0}.fasta".format(c_id))
    with open(fasta, "w") as f:
        f.write(">consensus_cl_id_{0}_total_supporting_reads_{1}\n{2}\n".format(c_id, n_reads, centre))
    fastq = os.path.join(outfolder, "reads_to_consensus_{0}.fastq".format(c_id))
    with open(fastq, "w") as f:
        for path in read_files:
            unique = {}
            for acc, seq, qual in _read_fastq(path):
                unique[acc] = (seq, qual)
            f.writelines("@{0}\n{1}\n+\n{2}\n".format(acc.split()[0], seq, qual) for acc, (seq, qual) in unique.items())
    return fasta, fastq


def _second_line(path):
    with open(path, "r") as f:
        return f.readlines()[1].strip()


def polish_sequences(centers, args):
    """Polishes every centre in place (centers[i][2]) and returns `centers`; files and folders as the
    reference writes them (modules/consensus.py:186-246): consensus_reference_<id>.fasta,
    reads_to_consensus_<id>.fastq, racon_cl_id_<id>/ or medaka_cl_id_<id>/ with consensus.fasta.
    With --racon all centres go through the kernels in one batch (they are independent); --medaka
    shells out per centre like the reference."""
    use_medaka = bool(getattr(args, "medaka", False))
    prefix = "medaka_cl_id_" if use_medaka else "racon_cl_id_"
    _clear_previous_outputs(args.outfolder, prefix)
    racon_jobs = []
    for row in centers:
        n_reads, c_id, centre, read_files = row
        fasta, fastq = _write_centre_inputs(args.outfolder, c_id, n_reads, centre, read_files)
        folder = os.path.join(args.outfolder, prefix + str(c_id))
        if use_medaka:
            help_functions.mkdir_p(folder)
            run_medaka(fastq, fasta, folder, "1", args.medaka_model, outfastq=getattr(args, "medaka_fastq", False))
            produced = [p for p in (os.path.join(folder, "consensus.fasta"), os.path.join(folder, "consensus.fastq")) if os.path.isfile(p)]
            if produced:
                row[2] = _second_line(produced[0])
            assert row[2], "Medaka consensus sequence not found"
        elif args.racon:
            help_functions.mkdir_p(folder)
            racon_jobs.append((fastq, fasta, folder, row))
    if racon_jobs:
        _racon_batch([job[:3] for job in racon_jobs], args.racon_iter)
        for _fq, _fa, folder, row in racon_jobs:
            row[2] = _second_line(os.path.join(folder, "consensus.fasta"))
    return centers
