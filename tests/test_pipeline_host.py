"""Host logic of multi_gpu.Pipeline that needs no GPU: double-buffered input (prefetch) and overlapped steps
(cluster_stream) on a minimal stand-in engine -- which engine holds which batch, what runs on which thread,
how errors surface."""
import threading
import time

import numpy as np
import pytest

from ngspeciesid_b200 import multi_gpu as M


class StubEngine(object):
    world = 1

    def __init__(self, name, fail_upload=False):
        self.name, self.fail_upload = name, fail_upload
        self.uploads, self.passes, self.threads = [], 0, set()

    def upload(self, seq, qual, off):
        if self.fail_upload:
            raise RuntimeError("upload failed on " + self.name)
        self.uploads.append(int(seq[0]))
        self.threads.add(threading.current_thread().name)

    def sync(self):
        pass

    def minimizers(self, k, w):
        pass

    def quality_stats(self):
        pass

    def cluster(self, k, w, max_gap, order, ranks, **kw):
        time.sleep(0.01)
        self.passes += 1
        self.threads.add(threading.current_thread().name)
        a = np.zeros(len(order), dtype=np.int32)
        a[0] = -1                                            # read 0 is the only representative
        return a, None, {"n_new_reps": 1}


ACCS = ["r%d_%d.0" % (i, 9 - i) for i in range(6)]
SCORES = [float(9 - i) for i in range(6)]


def batch(tag):
    return (np.full(4, tag, dtype=np.uint8), np.full(4, 33, dtype=np.uint8), np.array([0, 4], dtype=np.int64))


def test_prefetch_alternates_engines_and_keeps_roots():
    a, b = StubEngine("a"), StubEngine("b")
    p = M.Pipeline(a, StubEngine("mg"), None, None, alt=b)
    p.prefetch(batch(1))
    for s in range(4):
        roots = p.cluster(None, ACCS, SCORES, 100, 6, prefetched=True, then_prefetch=batch(s + 2) if s < 3 else None)
        assert list(roots) == [100] * 6                      # everything joins read 0 (global id 100)
        assert p.eng is (b if s % 2 == 0 else a) and p.alt is (a if s % 2 == 0 else b)
    assert b.uploads == [1, 3] and a.uploads == [2, 4]
    assert a.passes + b.passes == 4
    assert "MainThread" not in "".join(sorted(t for t in a.threads | b.threads if "Thread-" in t))
    with pytest.raises(RuntimeError):
        p.cluster(None, ACCS, SCORES, 0, 6, prefetched=True)  # nothing in flight


def test_prefetch_error_surfaces_in_cluster():
    p = M.Pipeline(StubEngine("a"), StubEngine("mg"), None, None, alt=StubEngine("b", fail_upload=True))
    p.prefetch(batch(1))
    with pytest.raises(RuntimeError, match="upload failed on b"):
        p.cluster(None, ACCS, SCORES, 0, 6, prefetched=True)


def test_cluster_stream_overlaps_and_reports_every_step():
    a, b = StubEngine("a"), StubEngine("b")
    p = M.Pipeline(a, StubEngine("mg"), None, None, alt=b)
    seen = []
    roots = p.cluster_stream(5, None, ACCS, SCORES, 7, 6, upload=batch(9),
                             on_step=lambda s, r, L: seen.append((s, L["eng"].name, list(r))))
    assert [s for s, _e, _r in seen] == [0, 1, 2, 3, 4]
    assert [e for _s, e, _r in seen] == ["a", "b", "a", "b", "a"]
    assert all(r == [7] * 6 for _s, _e, r in seen) and list(roots) == [7] * 6
    assert p.eng is a and p.alt is b                         # the engine of the last step holds its batch
    assert a.passes == 3 and b.passes == 2 and a.uploads == [9, 9, 9]
    assert all(t != "MainThread" for t in a.threads | b.threads)    # local passes ran on the second host thread
    assert p.phase["cluster_local"] > 0 and "wait_local_pass" in p.phase
    assert p.local_stats == {"n_new_reps": 1} and list(p.local_reps) == [0]


def test_cluster_stream_worker_error_reaches_the_caller():
    p = M.Pipeline(StubEngine("a", fail_upload=True), StubEngine("mg"), None, None, alt=StubEngine("b"))
    with pytest.raises(RuntimeError, match="upload failed on a"):
        p.cluster_stream(3, None, ACCS, SCORES, 0, 6, upload=batch(1))
    with pytest.raises(RuntimeError, match="alternate engine"):
        M.Pipeline(StubEngine("a"), StubEngine("mg"), None, None).cluster_stream(1, None, ACCS, SCORES, 0, 6)
