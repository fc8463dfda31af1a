"""Per-thread phases of the K1 stream kernel, compiled for the host, against the oracle. CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import cluster_oracle as oc


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libk1_stream_host.so")
    src = os.path.join(ROOT, "tests", "k1_stream_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.k1s_host_minimizers.restype = ctypes.c_int
    lib.k1s_host_minimizers.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return lib


def run(lib, seq, k, extra=0, garbage=0, tight=0):
    cap = len(seq) + 8
    code = np.zeros(cap, dtype=np.uint32)
    pos = np.zeros(cap, dtype=np.uint32)
    lc = ctypes.c_int(0)
    n = lib.k1s_host_minimizers(seq.encode(), len(seq), k, extra, garbage, tight, code.ctypes.data, pos.ctypes.data,
                                cap, ctypes.byref(lc))
    return n, lc.value, code[:max(n, 0)], pos[:max(n, 0)]


def decode(c, k):
    return "".join("ACGT"[(int(c) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def check(lib, seq, k, extra=0, garbage=0):
    w = k + 7
    seqc, _ = oc.hpol_compress(seq)
    n, lc, code, pos = run(lib, seq, k, extra, garbage)
    assert lc == len(seqc)
    if len(seqc) - w + 1 < 8:
        assert n == -1
        return
    exp = oc.minimizers(seqc, k, w)
    got = [(decode(c, k), int(p)) for c, p in zip(code, pos)]
    assert got == exp, (seq, k)


@pytest.mark.parametrize("k", [13, 12, 9, 5])
def test_random_reads(host, k):
    rng = np.random.default_rng(100 + k)
    for _ in range(300):
        L = int(rng.integers(20, 900))
        seq = "".join(rng.choice(list("ACGT"), size=L))
        check(host, seq, k, extra=int(rng.integers(0, 3)), garbage=int(rng.integers(0, 5)))


def test_low_complexity_and_ties(host):
    rng = np.random.default_rng(7)
    for _ in range(300):
        L = int(rng.integers(40, 700))
        kind = int(rng.integers(0, 4))
        if kind == 0:      # dinucleotide / short-period repeats: equal k-mers inside one window
            unit = "".join(rng.choice(list("ACGT"), size=int(rng.integers(2, 7))))
            seq = (unit * (L // len(unit) + 1))[:L]
        elif kind == 1:    # long homopolymers
            seq = "".join(ch * int(rng.integers(1, 40)) for ch in rng.choice(list("ACGT"), size=L // 6 + 4))
        elif kind == 2:    # two-letter alphabet
            seq = "".join(rng.choice(list("AC"), size=L))
        else:              # repeat with a few mutations
            unit = "".join(rng.choice(list("ACGT"), size=int(rng.integers(3, 9))))
            s = list((unit * (L // len(unit) + 1))[:L])
            for i in rng.integers(0, L, size=3):
                s[i] = "ACGT"[int(rng.integers(0, 4))]
            seq = "".join(s)
        for k in (13, 7):
            check(host, seq, k, extra=int(rng.integers(0, 2)), garbage=int(rng.integers(0, 3)))


def test_lengths_around_word_and_step_boundaries(host):
    rng = np.random.default_rng(11)
    base = "".join(rng.choice(list("ACGT"), size=1200))
    comp, _ = oc.hpol_compress(base)
    # feed already-compressed text so that the compressed length is exactly the raw length
    for L in list(range(27, 140)) + list(range(500, 600)) + [1023, 1024, 1025]:
        check(host, comp[:L], 13, extra=L % 2, garbage=L % 3)


def test_lut_entries(host):
    # every (previous base, 4 bases) combination through phase A on a 5-base input
    for prev in "ACGT":
        for a in "ACGT":
            for b in "ACGT":
                seq = prev + a + b + "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
                check(host, seq, 13)


def test_region_overflow_is_reported(host):
    """A read that compresses to more than 85 % of the sizing length must be refused (-4), never
    written past its region; one that fits the tight region must still be exact."""
    rng = np.random.default_rng(5)
    base = "".join(rng.choice(list("ACGT"), size=900))
    comp, _ = oc.hpol_compress(base)
    n, lc, _c, _p = run(host, comp[:600], 13, tight=1)       # compressed length == raw length
    assert n == -4 and lc == -1
    seq = "".join(ch * 2 for ch in comp[:300])               # compresses to 50 %
    n, lc, code, pos = run(host, seq, 13, tight=1)
    exp = oc.minimizers(comp[:300], 13, 20)
    assert lc == 300 and [(decode(c, 13), int(p)) for c, p in zip(code, pos)] == exp
