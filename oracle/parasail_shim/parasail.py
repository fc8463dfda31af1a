"""
oracle/parasail_shim/parasail.py -- TEST INFRASTRUCTURE ONLY.

Stand-in for the third-party `parasail` package (pinned ==1.2.4 in the reference's setup.py:78,
absent from this image) exposing exactly what the reference touches:
  parasail.matrix_create(alphabet, match, mismatch)            modules/cluster.py:131
  parasail.sg_trace_scan_16 / sg_trace_scan_32(s1, s2, o, e, M) modules/cluster.py:132,135
  result.saturated, result.score, result.cigar.decode (bytes)  modules/cluster.py:133,140; consensus.py:73
The arithmetic is oracle/sg_align.c (see its header for the tie-break choices; PARITY UNPINNED
against real parasail). Put this directory first on sys.path to drive the unmodified reference.
"""
import ctypes
import os
import time

_here = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_here, "..", "_build", "liboracle.so"))
_lib.oracle_sg_align.restype = ctypes.c_int
_lib.oracle_sg_align.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                 ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]

CALLS = {"n": 0, "cells": 0, "seconds": 0.0}


class Matrix(object):
    def __init__(self, alphabet, match, mismatch):
        self.alphabet, self.match, self.mismatch = alphabet, match, mismatch


def matrix_create(alphabet, match, mismatch):
    return Matrix(alphabet, match, mismatch)


class _Cigar(object):
    def __init__(self, ops):
        out = []
        i, n = 0, len(ops)
        while i < n:
            j = i
            while j < n and ops[j] == ops[i]:
                j += 1
            out.append(b"%d%c" % (j - i, ops[i]))
            i = j
        self.decode = b"".join(out)


class _Result(object):
    def __init__(self, score, ops, end_query, end_ref):
        self.saturated = False
        self.score = score
        self.end_query = end_query
        self.end_ref = end_ref
        self.cigar = _Cigar(ops)


def _align(s1, s2, open_, ext, matrix):
    t0 = time.perf_counter()
    b1, b2 = s1.encode(), s2.encode()
    buf = ctypes.create_string_buffer(len(b1) + len(b2) + 2)
    score, ei, ej = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    n = _lib.oracle_sg_align(b1, len(b1), b2, len(b2), matrix.match, matrix.mismatch, open_, ext,
                             buf, ctypes.byref(score), ctypes.byref(ei), ctypes.byref(ej))
    if n < 0:
        raise MemoryError("oracle_sg_align")
    CALLS["n"] += 1
    CALLS["cells"] += len(b1) * len(b2)
    CALLS["seconds"] += time.perf_counter() - t0
    return _Result(score.value, buf.raw[:n], ei.value, ej.value)


def sg_trace_scan_16(s1, s2, open_, ext, matrix):
    return _align(s1, s2, open_, ext, matrix)


def sg_trace_scan_32(s1, s2, open_, ext, matrix):
    return _align(s1, s2, open_, ext, matrix)


sg_trace = sg_trace_scan_32
