"""
Synthetic amplicon read generator (ONT / PacBio error profiles) used by bench.py and the tests.

Implements the workload definition of SURVEY.md section 8(d): S species templates derived from one
random template by substitutions + indels, reads drawn uniformly over species and strand, a
per-read error rate, indel-dominated errors with extra weight inside homopolymers, PHRED qualities
tied to the error rate with erroneous bases forced low. Everything is produced as flat numpy
arrays (ASCII bases, ASCII qualities, offsets) so that 10^6-read sets never become Python strings
unless a caller asks for them.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


class ReadSet(object):
    """Flat read container: seq/qual are uint8 ASCII, offsets has n+1 int64 entries."""

    def __init__(self, seq, qual, offsets, names=None, species=None, strand=None, templates=None):
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.qual = np.ascontiguousarray(qual, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.names = names
        self.species = species
        self.strand = strand
        self.templates = templates

    def __len__(self):
        return len(self.offsets) - 1

    def lengths(self):
        return np.diff(self.offsets)

    def read(self, i):
        a, b = self.offsets[i], self.offsets[i + 1]
        return self.seq[a:b].tobytes().decode(), self.qual[a:b].tobytes().decode()

    def name(self, i):
        if self.names is not None:
            return self.names[i]
        sp = int(self.species[i]) if self.species is not None else 0
        st = "+" if (self.strand is None or self.strand[i] == 0) else "-"
        return "read%d species=%d strand=%s" % (i, sp, st)

    def records(self):
        for i in range(len(self)):
            s, q = self.read(i)
            yield self.name(i), s, q

    def write_fastq(self, path):
        with open(path, "w") as f:
            for acc, s, q in self.records():
                f.write("@%s\n%s\n+\n%s\n" % (acc, s, q))

    def subset(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths()[idx]
        offs = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        seq = np.empty(offs[-1], dtype=np.uint8)
        qual = np.empty(offs[-1], dtype=np.uint8)
        for n, i in enumerate(idx):
            a, b = self.offsets[i], self.offsets[i + 1]
            seq[offs[n]:offs[n + 1]] = self.seq[a:b]
            qual[offs[n]:offs[n + 1]] = self.qual[a:b]
        names = [self.name(int(i)) for i in idx]
        return ReadSet(seq, qual, offs, names=names,
                       species=None if self.species is None else self.species[idx],
                       strand=None if self.strand is None else self.strand[idx],
                       templates=self.templates)


def _mutate(rng, tpl, sub_rate, indel_rate):
    out = []
    for b in tpl:
        r = rng.random()
        if r < indel_rate / 2:
            continue
        if r < indel_rate:
            out.append(b)
            out.append(_ACGT[rng.integers(4)])
            continue
        if rng.random() < sub_rate:
            choices = _ACGT[_ACGT != b]
            out.append(choices[rng.integers(3)])
        else:
            out.append(b)
    return np.array(out, dtype=np.uint8)


def make_templates(rng, n_species, len_lo, len_hi, independent=False):
    """First template uniform random; the others carry 8-20 % substitutions + 1 % indels
    relative to it (barcode-like divergence), or are independent when `independent`."""
    base = _ACGT[rng.integers(0, 4, size=int(rng.integers(len_lo, len_hi + 1)))]
    tpls = [base]
    for _ in range(1, n_species):
        if independent:
            tpls.append(_ACGT[rng.integers(0, 4, size=int(rng.integers(len_lo, len_hi + 1)))])
        else:
            tpls.append(_mutate(rng, base, rng.uniform(0.08, 0.20), 0.01))
    return tpls


_PROFILES = {
    # mean, sd, lo, hi of the per-read error rate; (del, ins, sub) split; hpol weight
    "ont": (0.08, 0.03, 0.02, 0.18, (0.40, 0.25, 0.35), 2.0),
    "pacbio": (0.02, 0.01, 0.005, 0.05, (0.45, 0.40, 0.15), 2.0),
}


def simulate_reads(n_reads, n_species=10, len_lo=700, len_hi=800, seed=1002, profile="ont",
                   independent=False, templates=None, per_read_len=None):
    """Returns a ReadSet of n_reads synthetic amplicon reads.

    per_read_len: optional (lo, hi); when given every read covers a random-length prefix of its
    template (used by the mixed-length PacBio configuration, templates then being >= hi long).
    Vectorised over all reads that share a (species, strand) template.
    """
    rng = np.random.default_rng(seed)
    if templates is None:
        templates = make_templates(rng, n_species, len_lo, len_hi, independent)
    n_species = len(templates)
    mean, sd, lo, hi, split, hw = _PROFILES[profile]
    species = rng.integers(0, n_species, size=n_reads)
    strand = rng.integers(0, 2, size=n_reads).astype(np.uint8)
    erate = np.clip(rng.normal(mean, sd, size=n_reads), lo, hi)
    p_del, p_ins, _p_sub = split

    def hp_weight(t):
        same_prev = np.zeros(len(t), dtype=bool)
        same_prev[1:] = t[1:] == t[:-1]
        same_next = np.zeros(len(t), dtype=bool)
        same_next[:-1] = t[1:] == t[:-1]
        return np.where(same_prev | same_next, hw, 1.0)

    lengths = np.zeros(n_reads, dtype=np.int64)
    chunks = {}                         # read index -> (bases, quals) views, filled per group
    group_out = []
    for s in range(n_species):
        for st in (0, 1):
            idx = np.nonzero((species == s) & (strand == st))[0]
            if len(idx) == 0:
                continue
            t = templates[s] if st == 0 else _COMP[templates[s][::-1]]
            w = hp_weight(t)
            L = len(t)
            n = len(idx)
            e = erate[idx][:, None]
            if per_read_len is not None:
                ln = rng.integers(per_read_len[0], min(per_read_len[1], L) + 1, size=n)
                live = np.arange(L)[None, :] < ln[:, None]
                wsum = np.where(live, w[None, :], 0.0).sum(axis=1)[:, None]
                lnf = ln[:, None].astype(np.float64)
            else:
                live = None
                wsum = w.sum()
                lnf = float(L)
            pe = np.minimum(e * w[None, :] * (lnf / wsum), 0.9)
            err = rng.random((n, L)) < pe
            kind = rng.random((n, L))
            is_del = err & (kind < p_del)
            is_ins = err & (kind >= p_del) & (kind < p_del + p_ins)
            is_sub = err & (kind >= p_del + p_ins)
            code = np.broadcast_to(np.searchsorted(_ACGT, t)[None, :], (n, L))
            shift = rng.integers(1, 4, size=(n, L))
            bases = _ACGT[np.where(is_sub, (code + shift) & 3, code)]
            qmean = -10.0 * np.log10(e)
            q = np.clip(np.rint(rng.normal(qmean, 4.0, size=(n, L))), 2, 50)
            low = is_sub | is_ins
            q = np.where(low, np.minimum(q, rng.integers(2, 9, size=(n, L))), q)
            reps = np.ones((n, L), dtype=np.int64)
            reps[is_del] = 0
            reps[is_ins] = 2
            if live is not None:
                reps[~live] = 0
            flat_reps = reps.ravel()
            out_b = np.repeat(bases.ravel(), flat_reps)
            out_q = np.repeat(q.ravel(), flat_reps)
            ins_flat = (is_ins if live is None else (is_ins & live)).ravel()
            nins = int(ins_flat.sum())
            if nins:
                pos = np.cumsum(flat_reps)[ins_flat] - 1
                out_b[pos] = _ACGT[rng.integers(0, 4, size=nins)]
                out_q[pos] = rng.integers(2, 9, size=nins)
            lens = reps.sum(axis=1)
            lengths[idx] = lens
            group_out.append((idx, lens, out_b, (out_q + 33).astype(np.uint8)))
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    seq = np.empty(offsets[-1], dtype=np.uint8)
    qual = np.empty(offsets[-1], dtype=np.uint8)
    for idx, lens, out_b, out_q in group_out:
        # destination index of every base of the group = start of its read + rank inside the read
        starts = offsets[idx]
        gofs = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=gofs[1:])
        dest = np.repeat(starts - gofs[:-1], lens) + np.arange(gofs[-1])
        seq[dest] = out_b
        qual[dest] = out_q
    return ReadSet(seq, qual, offsets, species=species, strand=strand, templates=templates)
