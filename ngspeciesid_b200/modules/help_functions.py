"""Host utilities with the reference's names (modules/help_functions.py): FASTA/FASTQ reader and
mkdir_p. Plain host I/O, not part of the accelerated path."""
import errno
import os


def readfq(fp):
    """Generator over (name, (seq, qual)) records with the contract of the reference's reader
    (modules/help_functions.py:13-42): the name is the whole header line, FASTA records and a
    FASTQ record whose quality is cut short by the end of the file yield qual None. Like the
    reference, every line loses its last character unseen, so the last line of a file without a
    final newline is one character short (golden vectors: tests/golden/readfq.json.gz)."""
    header = None
    lines = iter(fp)
    while True:
        if not header:
            for line in lines:
                if line[0] in ">@":
                    header = line[:-1]
                    break
        if not header:
            return
        name, parts, header = header[1:], [], None
        for line in lines:
            if line[0] in "@+>":
                header = line[:-1]
                break
            parts.append(line[:-1])
        seq = "".join(parts)
        if not header or header[0] != "+":
            yield name, (seq, None)
            if not header:
                return
            continue
        parts, got, complete = [], 0, False
        for line in lines:
            parts.append(line[:-1])
            got += len(line) - 1
            if got >= len(seq):
                complete = True
                break
        if not complete:
            yield name, (seq, None)
            return
        header = None
        yield name, (seq, "".join(parts))


class FastqArrays(object):
    """A FASTA/FASTQ file parsed by ngsid_fastq_parse (host C code in libngsid.so): concatenated
    sequences and qualities as uint8 arrays with offsets (the layout Engine.upload takes) and the
    names as spans of the file buffer -- no Python string per read until one is asked for."""

    def __init__(self, buf, seq, qual, seq_off, qual_off, name_off, name_len, has_qual):
        self.buf, self.seq, self.qual = buf, seq, qual
        self.seq_off, self.qual_off = seq_off, qual_off
        self.name_off, self.name_len, self.has_qual = name_off, name_len, has_qual

    def __len__(self):
        return len(self.name_off)

    def name(self, i):
        a = int(self.name_off[i])
        return bytes(self.buf[a:a + int(self.name_len[i])]).decode()

    def record(self, i):
        """(name, (seq, qual)) as readfq yields it."""
        s = self.seq[self.seq_off[i]:self.seq_off[i + 1]].tobytes().decode()
        q = self.qual[self.qual_off[i]:self.qual_off[i + 1]].tobytes().decode() if self.has_qual[i] else None
        return self.name(i), (s, q)

    def records(self):
        for i in range(len(self)):
            yield self.record(i)


def parse_fastq_bytes(data):
    """bytes / uint8 array of a whole FASTA/FASTQ file -> FastqArrays (same records as readfq on
    the file opened in text mode)."""
    import ctypes
    import numpy as np
    from .. import _lib
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = ctypes.c_int64(0)
    rc = lib.ngsid_fastq_parse(_lib.ptr(buf), len(buf), 0, None, None, None, None, None, None, None, ctypes.byref(n))
    if rc:
        raise _lib.NgsidError(rc, "ngsid_fastq_parse")
    cap = n.value
    seq = np.empty(max(1, len(buf)), dtype=np.uint8)
    qual = np.empty(max(1, len(buf)), dtype=np.uint8)
    name_off = np.zeros(cap, dtype=np.int64)
    name_len = np.zeros(cap, dtype=np.int32)
    seq_off = np.zeros(cap + 1, dtype=np.int64)
    qual_off = np.zeros(cap + 1, dtype=np.int64)
    has_qual = np.zeros(max(1, cap), dtype=np.uint8)
    rc = lib.ngsid_fastq_parse(_lib.ptr(buf), len(buf), cap, _lib.ptr(seq), _lib.ptr(qual), _lib.ptr(name_off),
                               _lib.ptr(name_len), _lib.ptr(seq_off), _lib.ptr(qual_off), _lib.ptr(has_qual),
                               ctypes.byref(n))
    if rc:
        raise _lib.NgsidError(rc, "ngsid_fastq_parse")
    return FastqArrays(buf, seq[:seq_off[cap]], qual[:qual_off[cap]], seq_off, qual_off, name_off, name_len,
                       has_qual[:cap])


def read_fastq_arrays(path):
    with open(path, "rb") as f:
        return parse_fastq_bytes(f.read())


def mkdir_p(path):
    try:
        os.makedirs(path)
    except OSError as exc:
        if not (exc.errno == errno.EEXIST and os.path.isdir(path)):
            raise
