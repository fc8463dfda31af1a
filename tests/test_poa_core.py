"""The product's POA graph core (compiled for the host) against the independent oracle. CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import consensus_oracle as co


@pytest.fixture(scope="module")
def core():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpoa_core_host.so")
    src = os.path.join(ROOT, "tests", "poa_core_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.poa_core_host_consensus.restype = ctypes.c_int
    lib.poa_core_host_consensus.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                            ctypes.c_int] + [ctypes.c_int] * 5 + [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    return lib


def run_core(lib, seqs, quals, mode, m, x, g, trim):
    n = len(seqs)
    a = (ctypes.c_char_p * n)(*[s.encode() for s in seqs])
    q = (ctypes.c_char_p * n)(*[s.encode() for s in quals])
    cap = sum(len(s) for s in seqs) + 16
    out = ctypes.create_string_buffer(cap)
    r = lib.poa_core_host_consensus(a, q, n, mode, m, x, g, 1 if trim else 0, out, cap, 20000)
    assert r >= 0, r
    return out.value.decode()


def noisy_cluster(rng, L, n, e):
    tpl = "".join(rng.choice(list("ACGT"), size=L))
    seqs, quals = [], []
    for _ in range(n):
        s, q = [], []
        for ch in tpl:
            r = rng.random()
            if r < e * 0.4:
                continue
            if r < e * 0.65:
                s.append("ACGT"[rng.integers(4)]); q.append(chr(33 + int(rng.integers(2, 9))))
            if r < e:
                s.append("ACGT"[rng.integers(4)]); q.append(chr(33 + int(rng.integers(2, 9))))
            else:
                s.append(ch); q.append(chr(33 + int(rng.integers(8, 30))))
        seqs.append("".join(s)); quals.append("".join(q))
    return tpl, seqs, quals


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_local_quality_weighted_matches_oracle(core, seed):
    rng = np.random.default_rng(seed)
    tpl, seqs, quals = noisy_cluster(rng, int(rng.integers(60, 300)), int(rng.integers(2, 40)), 0.12)
    exp = co.poa_consensus(seqs, quals, mode=0, match=5, mismatch=-4, gap=-2)
    assert run_core(core, seqs, quals, 0, 5, -4, -2, False) == exp


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_global_window_with_backbone_matches_oracle(core, seed):
    rng = np.random.default_rng(seed)
    tpl, seqs, quals = noisy_cluster(rng, int(rng.integers(80, 250)), int(rng.integers(3, 30)), 0.1)
    seqs = [seqs[0]] + seqs
    quals = [""] + quals                     # backbone carries no weight
    exp = co.poa_consensus(seqs, quals, mode=1, match=3, mismatch=-5, gap=-4, trim=True)
    assert run_core(core, seqs, quals, 1, 3, -5, -4, True) == exp


def test_degenerate_inputs(core):
    for seqs, quals in ([["ACGT"], ["5555"]], [["A", "C", "G"], ["5", "5", "5"]],
                        [["ACGTACGT", "TTTTTTTT", "ACGTACGT"], ["55555555"] * 3]):
        exp = co.poa_consensus(seqs, quals, mode=0, match=5, mismatch=-4, gap=-2)
        assert run_core(core, seqs, quals, 0, 5, -4, -2, False) == exp


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_subgraph_view_matches_oracle(core, seed):
    """poa_subgraph_view (the part of a window graph racon aligns a non-spanning layer to) against the oracle's
    restatement of spoa Graph::subgraph: same nodes in the same order, on graphs both sides built themselves."""
    from oracle import cluster_oracle as oc
    olib = oc._lib()
    rng = np.random.default_rng(seed)
    tpl, seqs, quals = noisy_cluster(rng, int(rng.integers(120, 260)), int(rng.integers(8, 30)), 0.12)
    seqs, quals = [tpl] + seqs, [""] + quals                      # backbone first, weight 0, like a racon window
    n = len(seqs)
    a = (ctypes.c_char_p * n)(*[s.encode() for s in seqs])
    q = (ctypes.c_char_p * n)(*[s.encode() for s in quals])
    cap = 20000
    core.poa_core_host_subview.restype = ctypes.c_int
    core.poa_core_host_subview.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p)] + [ctypes.c_int] * 7 + \
                                          [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]
    olib.oracle_poa_subview.restype = ctypes.c_int
    olib.oracle_poa_subview.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p)] + [ctypes.c_int] * 7 + \
                                       [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L = len(tpl)
    for b, e in [(0, L - 1), (10, L - 20), (L // 3, 2 * L // 3), (5, 5), (L - 2, L - 1), (0, 0)]:
        got, want = (ctypes.c_int * cap)(), (ctypes.c_int * cap)()
        ng = core.poa_core_host_subview(a, q, n, 1, 3, -5, -4, b, e, got, cap, 20000)
        nw = olib.oracle_poa_subview(a, q, n, 1, 3, -5, -4, b, e, want, cap)
        assert ng == nw and ng >= e - b + 1
        assert list(got[:ng]) == list(want[:nw])
        assert set(range(b, e + 1)) <= set(got[:ng])          # the backbone stretch is in the view
        assert all(v >= b for v in got[:ng])
