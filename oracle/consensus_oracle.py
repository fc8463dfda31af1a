"""
oracle/consensus_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's consensus hot path (modules/consensus.py), whose arithmetic
lives in the external binaries spoa / minimap2 / racon (absent here; PARITY UNPINNED, see
oracle/poa_oracle.cpp):

  run_spoa      modules/consensus.py:83-92    `spoa reads.fq -l 0 -r 0 -g -2`
                -> local POA, match 5, mismatch -4, linear gap -2, quality weights, heaviest bundle
  run_racon     modules/consensus.py:107-126  racon_iter x (minimap2 -x map-ont; racon)
                -> per iteration: every read is aligned to the current target (here: the semi-global
                   aligner of oracle/sg_align.c, open 3 / extend 1, instead of minimap2 + edlib);
                   the target is cut into 500-base windows; a read contributes to a window the
                   stretch between its first and last aligned (=/X) column inside the window
                   (racon's breaking points) when that stretch is at least 2 % of the window, has
                   mean quality >= 10; a layer that spans the window to within 1 % at both ends is
                   aligned to the whole window graph, a shorter one to the sub-graph between its first
                   and last backbone position (spoa Graph::subgraph, oracle/poa_oracle.cpp);
                   window consensus = global POA (3 / -5 / -4) of backbone (weight 0) + layers
                   sorted by start, coverage-trimmed; windows with < 3 sequences keep the backbone;
                   the polished target is the concatenation.
"""
import ctypes

from . import cluster_oracle as _co

WINDOW = 500


def _lib():
    lib = _co._lib()
    if not hasattr(lib, "_poa_ready"):
        lib.oracle_poa_consensus.restype = ctypes.c_int
        lib.oracle_poa_consensus.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_poa_consensus_ex.restype = ctypes.c_int
        lib.oracle_poa_consensus_ex.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_int)]
        lib.oracle_poa_consensus_sub.restype = ctypes.c_int
        lib.oracle_poa_consensus_sub.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                 ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_int)]
        lib.oracle_sg_align.restype = ctypes.c_int
        lib.oracle_sg_align.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        lib._poa_ready = True
    return lib


ORDER_MODE = 0        # 0 = spoa's re-sort (reference-faithful), 1 = path insertion (the kernel's order)


def poa_consensus(seqs, quals=None, mode=0, match=5, mismatch=-4, gap=-2, trim=False, order_mode=None, sub=None):
    """sub: optional [(begin, end) or None per sequence]: backbone positions of the first sequence that the
    sequence is aligned between (racon's sub-graph alignment); None / missing = the whole graph."""
    n = len(seqs)
    arr = (ctypes.c_char_p * n)(*[s.encode() for s in seqs])
    qarr = None
    if quals is not None:
        qarr = (ctypes.c_char_p * n)(*[q.encode() for q in quals])
    cap = sum(len(s) for s in seqs) + 16
    out = ctypes.create_string_buffer(cap)
    nn = ctypes.c_int(0)
    om = ORDER_MODE if order_mode is None else order_mode
    if sub is not None:
        sb = (ctypes.c_int * n)(*[(-1 if x is None else int(x[0])) for x in sub])
        se = (ctypes.c_int * n)(*[(-1 if x is None else int(x[1])) for x in sub])
        r = _lib().oracle_poa_consensus_sub(arr, qarr, n, mode, match, mismatch, gap, 1 if trim else 0, om, sb, se, out, cap,
                                            ctypes.byref(nn))
    else:
        r = _lib().oracle_poa_consensus_ex(arr, qarr, n, mode, match, mismatch, gap, 1 if trim else 0, om, out, cap, ctypes.byref(nn))
    if r < 0:
        raise MemoryError("oracle_poa_consensus")
    return out.value.decode()


def spoa_consensus(records):
    """records: [(seq, qual)] in file order -> consensus (what run_spoa returns)."""
    return poa_consensus([r[0] for r in records], [r[1] for r in records], mode=0, match=5, mismatch=-4, gap=-2)


def align_ops(read, target, open_pen=3, ext=1):
    b1, b2 = read.encode(), target.encode()
    buf = ctypes.create_string_buffer(len(b1) + len(b2) + 2)
    score, ei, ej = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    n = _lib().oracle_sg_align(b1, len(b1), b2, len(b2), 2, -2, open_pen, ext, buf, ctypes.byref(score),
                               ctypes.byref(ei), ctypes.byref(ej))
    return buf.raw[:n].decode(), score.value


_RC = str.maketrans("ACGT", "TGCA")


def revcomp(s):
    return s.translate(_RC)[::-1]


def window_segments(ops, target_len, window=WINDOW):
    """Breaking points per window: {window index: (q_first, q_last_exclusive, t_first, t_last_exclusive)}."""
    out = {}
    qi = ti = 0
    first = None
    last = None
    for op in ops:
        if op in "=X":
            if first is None:
                first = (ti, qi)
            last = (ti + 1, qi + 1)
            if (ti + 1) % window == 0 or ti + 1 == target_len:
                out[ti // window] = (first[1], last[1], first[0], last[0])
                first = None
            qi += 1
            ti += 1
        elif op == "I":
            qi += 1
        else:  # D: target base against a gap
            if ((ti + 1) % window == 0 or ti + 1 == target_len) and first is not None:
                out[ti // window] = (first[1], last[1], first[0], last[0])
                first = None
            ti += 1
    return out


def racon_round(target, reads, window=WINDOW, both_strands=False):
    """One polishing round. reads: [(seq, qual)]."""
    n_win = (len(target) + window - 1) // window
    layers = [[] for _ in range(n_win)]
    for seq, qual in reads:
        ops, score = align_ops(seq, target)
        if both_strands:
            rs, rq = revcomp(seq), qual[::-1]
            ops2, score2 = align_ops(rs, target)
            if score2 > score:
                seq, qual, ops = rs, rq, ops2
        for wi, (q0, q1, t0, t1) in window_segments(ops, len(target), window).items():
            ws = wi * window
            wlen = min(window, len(target) - ws)
            if q1 - q0 < 0.02 * wlen:
                continue
            sq = qual[q0:q1]
            if sum(ord(c) - 33 for c in sq) / float(len(sq)) < 10.0:
                continue
            off = 0.01 * wlen
            b, e = t0 - ws, t1 - ws - 1
            spans = b < off and e > wlen - off
            layers[wi].append((b, seq[q0:q1], sq, None if spans else (b, e)))
    out = []
    for wi in range(n_win):
        ws = wi * window
        backbone = target[ws:ws + window]
        ls = sorted(layers[wi], key=lambda x: x[0])
        if len(ls) + 1 < 3:
            out.append(backbone)
            continue
        seqs = [backbone] + [l[1] for l in ls]
        quals = [""] + [l[2] for l in ls]
        out.append(poa_consensus(seqs, quals, mode=1, match=3, mismatch=-5, gap=-4, trim=True, sub=[None] + [l[3] for l in ls]))
    return "".join(out)


def racon_polish(target, reads, iters, both_strands=False):
    for _ in range(iters):
        target = racon_round(target, reads, both_strands=both_strands)
    return target


def edit_distance(a, b):
    """Plain Levenshtein distance (small inputs; used for the consensus tolerance)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]
