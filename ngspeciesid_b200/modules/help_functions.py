"""Host utilities with the reference's names (modules/help_functions.py): FASTA/FASTQ reader and
mkdir_p. Plain host I/O, not part of the accelerated path."""
import errno
import os


def readfq(fp):
    """Generator over (name, (seq, qual)) records; the name is the whole header line, FASTA records
    yield qual None (same contract as modules/help_functions.py:13-42)."""
    header = None
    for line in fp:
        if line[:1] in ">@":
            header = line[:-1] if line.endswith("\n") else line
            break
    while header is not None:
        name, seq_lines, nxt = header[1:], [], None
        for line in fp:
            if line[:1] in "@+>":
                nxt = line[:-1] if line.endswith("\n") else line
                break
            seq_lines.append(line.rstrip("\n"))
        seq = "".join(seq_lines)
        if nxt is None or nxt[0] != "+":
            yield name, (seq, None)
            header = nxt
            if header is None:
                break
            continue
        qual_lines, got = [], 0
        header = None
        complete = False
        for line in fp:
            q = line.rstrip("\n")
            qual_lines.append(q)
            got += len(q)
            if got >= len(seq):
                complete = True
                break
        if not complete:
            yield name, (seq, None)
            break
        yield name, (seq, "".join(qual_lines))
        for line in fp:
            if line[:1] in ">@":
                header = line[:-1] if line.endswith("\n") else line
                break


def mkdir_p(path):
    try:
        os.makedirs(path)
    except OSError as exc:
        if not (exc.errno == errno.EEXIST and os.path.isdir(path)):
            raise
