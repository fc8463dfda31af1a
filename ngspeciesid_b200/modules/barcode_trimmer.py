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
    """Reference: modules/barcode_trimmer.py:15-23."""
    barcodes = {acc + "_fw": seq.strip() for acc, (seq, _) in help_functions.readfq(open(primer_file, "r"))}
    for acc, seq in list(barcodes.items()):
        barcodes[acc[:-3] + "_rc"] = reverse_complement(seq.upper())
    return barcodes


def get_universal_tails():
    """Reference: modules/barcode_trimmer.py:25-31."""
    barcodes = {"1_F_fw": "TTTCTGTTGGTGCTGATATTGC", "2_R_rc": "ACTTGCCTGTCGCTCTATCTTC"}
    barcodes["1_F_rc"] = reverse_complement(barcodes["1_F_fw"])
    barcodes["2_R_fw"] = reverse_complement(barcodes["2_R_rc"])
    return barcodes


def find_barcode_locations(center, barcodes, primer_max_ed):
    """Reference: modules/barcode_trimmer.py:34-59 -> [(primer name, start, end, distance)] with
    the first location of every primer that is found."""
    all_locations = []
    for primer_acc, primer_seq in barcodes.items():
        ed, locations = find_locations(primer_seq, center, primer_max_ed)
        logging.debug(f"{locations} {ed}")
        if locations:
            all_locations.append((primer_acc, locations[0][0], locations[0][1], ed))
    return all_locations


def remove_barcodes(centers, barcodes, args):
    """Reference: modules/barcode_trimmer.py:62-104: cuts every consensus in `centers`
    ([n_reads, c_id, sequence, reads_path]) behind the last primer hit of its first `trim_window`
    bases and in front of the earliest hit of its last `trim_window` bases; returns whether any
    sequence changed. Like the reference, a hit at the beginning cuts at its (inclusive) end index."""
    centers_updated = False
    for i, (_nr_reads, _c_id, center, _reads_path) in enumerate(centers):
        trim_window = len(center) // 2 if 2 * args.trim_window > len(center) else args.trim_window
        begin = find_barcode_locations(center[:trim_window], barcodes, args.primer_max_ed)
        end = find_barcode_locations(center[-trim_window:], barcodes, args.primer_max_ed)
        cut_start = 0
        for _bc, _start, stop, _ed in begin:
            if stop > cut_start:
                cut_start = stop
        cut_end = len(center)
        if end:
            earliest_hit = len(center)
            for _bc, start, _stop, _ed in end:
                if start < earliest_hit:
                    earliest_hit = start
            cut_end = len(center) - (trim_window - earliest_hit)
        if cut_start > 0 or cut_end < len(center):
            centers[i][2] = center[cut_start:cut_end]
            centers_updated = True
    return centers_updated
