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

    per_read_len: optional (lo, hi); when given every read is a random-length prefix window of its
    template (used by the mixed-length PacBio configuration, templates then being >= hi long).
    """
    rng = np.random.default_rng(seed)
    if templates is None:
        templates = make_templates(rng, n_species, len_lo, len_hi, independent)
    n_species = len(templates)
    mean, sd, lo, hi, split, hw = _PROFILES[profile]
    species = rng.integers(0, n_species, size=n_reads)
    strand = rng.integers(0, 2, size=n_reads).astype(np.uint8)
    erate = np.clip(rng.normal(mean, sd, size=n_reads), lo, hi)

    fw = [t for t in templates]
    rc = [_COMP[t[::-1]] for t in templates]
    # homopolymer membership weight per template position
    def hp_weight(t):
        same_prev = np.zeros(len(t), dtype=bool)
        same_prev[1:] = t[1:] == t[:-1]
        same_next = np.zeros(len(t), dtype=bool)
        same_next[:-1] = t[1:] == t[:-1]
        return np.where(same_prev | same_next, hw, 1.0)
    fw_w = [hp_weight(t) for t in fw]
    rc_w = [hp_weight(t) for t in rc]

    seq_parts, qual_parts = [], []
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    p_del, p_ins, p_sub = split
    for i in range(n_reads):
        s = species[i]
        t, w = (fw[s], fw_w[s]) if strand[i] == 0 else (rc[s], rc_w[s])
        if per_read_len is not None:
            ln = int(rng.integers(per_read_len[0], min(per_read_len[1], len(t)) + 1))
            t, w = t[:ln], w[:ln]
        L = len(t)
        e = erate[i]
        # normalise so that the expected error fraction stays e
        pe = np.minimum(e * w * (L / w.sum()), 0.9)
        u = rng.random(L)
        err = u < pe
        kind = rng.random(L)
        is_del = err & (kind < p_del)
        is_ins = err & (kind >= p_del) & (kind < p_del + p_ins)
        is_sub = err & (kind >= p_del + p_ins)
        bases = t.copy()
        nsub = int(is_sub.sum())
        if nsub:
            shift = rng.integers(1, 4, size=nsub)
            code = np.searchsorted(_ACGT, bases[is_sub])  # A,C,G,T are sorted ASCII
            bases[is_sub] = _ACGT[(code + shift) & 3]
        # qualities
        qmean = -10.0 * np.log10(e)
        q = np.clip(np.rint(rng.normal(qmean, 4.0, size=L)), 2, 50)
        low = is_sub | is_ins
        q[low] = np.minimum(q[low], rng.integers(2, 9, size=int(low.sum())))
        reps = np.ones(L, dtype=np.int64)
        reps[is_del] = 0
        reps[is_ins] = 2
        out_b = np.repeat(bases, reps)
        out_q = np.repeat(q, reps)
        nins = int(is_ins.sum())
        if nins:
            # second copy of each inserted position becomes a random base with low quality
            pos = np.cumsum(reps)[is_ins] - 1
            out_b[pos] = _ACGT[rng.integers(0, 4, size=nins)]
            out_q[pos] = rng.integers(2, 9, size=nins)
        seq_parts.append(out_b)
        qual_parts.append((out_q + 33).astype(np.uint8))
        offsets[i + 1] = offsets[i] + len(out_b)
    seq = np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
    qual = np.concatenate(qual_parts) if qual_parts else np.zeros(0, np.uint8)
    return ReadSet(seq, qual, offsets, species=species, strand=strand, templates=templates)
