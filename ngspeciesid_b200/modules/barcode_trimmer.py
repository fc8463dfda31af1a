"""Primer / universal-tail trimming of the consensus ends with the reference's names
(modules/barcode_trimmer.py, SURVEY.md section 8 f rank 4). Host code: at most a few consensus
sequences per run, two windows of `trim_window` bases each, primers of ~20-30 bases.

The reference calls edlib (C++, pip, absent here): `edlib.align(primer, window, mode="HW",
task="locations", k=primer_max_ed, additionalEqualities=IUPAC_map)`. `find_locations` restates what
that call returns for this use -- PARITY UNPINNED against edlib itself (edlib >= 1.1.2 is not
installable in the build image and the reference holds no vectors); the properties it is tested
on are edlib's documented contract: infix ("HW") edit distance = minimum over all substrings of the
target, every end position that reaches it in ascending order, for each end the EARLIEST start
that reaches it, nothing when the distance exceeds k.
"""
import logging

from . import help_functions

# modules/barcode_trimmer.py:40-45: IUPAC codes of the primer match the bases they stand for
IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
         "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC"}
_EQUAL = set()
for _c, _bases in IUPAC.items():
    for _b in _bases:
        _EQUAL.add((_c, _b))
        _EQUAL.add((_b, _c))           # edlib's additional equalities are symmetric


def _same(a, b):
    return a == b or (a, b) in _EQUAL


def reverse_complement(string):
    """Reference: modules/barcode_trimmer.py:6-13 (IUPAC-aware, case preserving)."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "X": "X", "Y": "R", "R": "Y", "K": "M",
            "M": "K", "S": "S", "W": "W", "B": "V", "V": "B", "H": "D", "D": "H"}
    comp.update({k.lower(): v.lower() for k, v in list(comp.items()) if k != "X"})
    return "".join(comp[c] for c in reversed(string))


def _end_distances(query, target):
    """Sellers' recurrence: d[j] = edit distance of `query` against the best substring of `target`
    that ends at target position j (exclusive end j; j = 0 is the empty substring)."""
    m = len(query)
    prev = list(range(m + 1))                    # column for the empty target prefix: i deletions
    out = [prev[m]]
    for tj in target:
        cur = [0] * (m + 1)                      # a match may start anywhere: row 0 is free
        for i in range(1, m + 1):
            cur[i] = min(prev[i - 1] + (0 if _same(query[i - 1], tj) else 1), prev[i] + 1, cur[i - 1] + 1)
        out.append(cur[m])
        prev = cur
    return out


def find_locations(query, target, k):
    """-> (edit distance, [(start, end)]) with inclusive ends, as edlib reports them in HW mode with
    task="locations"; (-1, []) when the best infix distance exceeds k (k < 0: no limit)."""
    if not query or not target:
        return -1, []
    d = _end_distances(query, target)
    best = min(d[1:]) if len(d) > 1 else d[0]
    if len(d) == 1 or (k >= 0 and best > k):
        return -1, []
    ends = [j for j in range(1, len(d)) if d[j] == best]
    locs = []
    rq = query[::-1]
    for e in ends:
        # start of the alignment that ends at e: align the reversed query against the reversed
        # prefix, anchored at its first base (prefix mode); the longest extent that reaches `best`
        rt = target[:e][::-1]
        m = len(rq)
        prev = list(range(m + 1))
        far = 0 if prev[m] == best else -1
        for j, tj in enumerate(rt, 1):
            cur = [j] + [0] * m
            for i in range(1, m + 1):
                cur[i] = min(prev[i - 1] + (0 if _same(rq[i - 1], tj) else 1), prev[i] + 1, cur[i - 1] + 1)
            if cur[m] == best:
                far = j
            prev = cur
        locs.append((e - far, e - 1))
    return best, locs


def read_barcodes(primer_file):
    """Primers of a FASTA file as {name_fw: sequence, name_rc: reverse complement (upper case)}, the
    forward entries first (behaviour of the reference's modules/barcode_trimmer.py:15-23)."""
    with open(primer_file, "r") as handle:
        forward = [(name, seq.strip()) for name, (seq, _qual) in help_functions.readfq(handle)]
    table = dict((name + "_fw", seq) for name, seq in forward)
    table.update((name + "_rc", reverse_complement(seq.upper())) for name, seq in forward)
    return table


_UNIVERSAL_TAILS = (("1_F", "TTTCTGTTGGTGCTGATATTGC", "fw"), ("2_R", "ACTTGCCTGTCGCTCTATCTTC", "rc"))


def get_universal_tails():
    """The two ONT universal tails in both orientations (reference: modules/barcode_trimmer.py:25-31;
    the second tail is given as its reverse complement there too)."""
    table = {"%s_%s" % (name, strand): seq for name, seq, strand in _UNIVERSAL_TAILS}
    for name, seq, strand in _UNIVERSAL_TAILS:
        table["%s_%s" % (name, "rc" if strand == "fw" else "fw")] = reverse_complement(seq)
    return table


def find_barcode_locations(center, barcodes, primer_max_ed):
    """Reference: modules/barcode_trimmer.py:34-59 -> [(primer name, start, end, distance)] with
    the first location of every primer that is found."""
    hits = []
    for name, primer in barcodes.items():
        distance, where = find_locations(primer, center, primer_max_ed)
        logging.debug(f"{where} {distance}")
        if where:
            first_start, first_end = where[0]
            hits.append((name, first_start, first_end, distance))
    return hits


def _trimmed_span(center, barcodes, window, max_ed):
    """[a, b) of `center` that survives: behind the furthest end index of a primer hit inside the
    first `window` bases, in front of the earliest primer hit inside the last `window` bases."""
    head_hits = find_barcode_locations(center[:window], barcodes, max_ed)
    tail_hits = find_barcode_locations(center[-window:], barcodes, max_ed)
    a = max([0] + [end for _n, _s, end, _d in head_hits])
    b = len(center)
    if tail_hits:
        first = min([len(center)] + [start for _n, start, _e, _d in tail_hits])
        b = len(center) - (window - first)
    return a, b


def remove_barcodes(centers, barcodes, args):
    """Trims the primers off every consensus of `centers` ([n_reads, c_id, sequence, reads_path]) in
    place and tells whether anything changed (reference: modules/barcode_trimmer.py:62-104, including
    its conventions: the search window shrinks to half the sequence for short consensi, and a hit at
    the beginning cuts at the hit's inclusive end index)."""
    changed = False
    for entry in centers:
        center = entry[2]
        window = args.trim_window if 2 * args.trim_window <= len(center) else len(center) // 2
        a, b = _trimmed_span(center, barcodes, window, args.primer_max_ed)
        if a > 0 or b < len(center):
            entry[2] = center[a:b]
            changed = True
    return changed
