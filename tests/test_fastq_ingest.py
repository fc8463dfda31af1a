"""FASTA/FASTQ ingest (SURVEY.md 8 f rank 3) against golden vectors produced by the REFERENCE's
readfq (tests/golden/make_fastq_golden.py): the Python mirror of the generator and the host C
parser behind ngsid_fastq_parse (CPU only: the function takes no context and no GPU)."""
import gzip
import hashlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from ngspeciesid_b200.modules import help_functions as hf

G = load_golden("readfq.json.gz")


def _via_file(text, tmp_path):
    p = os.path.join(str(tmp_path), "x.fq")
    with open(p, "wb") as f:
        f.write(text.encode("ascii"))
    return p


@pytest.mark.parametrize("case", G["cases"], ids=[c["tag"] for c in G["cases"]])
def test_readfq_generator_matches_reference(case, tmp_path):
    with open(_via_file(case["text"], tmp_path), "r") as fp:
        got = [[n, s, q] for n, (s, q) in hf.readfq(fp)]
    assert got == case["records"]


@pytest.mark.parametrize("case", G["cases"], ids=[c["tag"] for c in G["cases"]])
def test_c_parser_matches_reference(case, tmp_path):
    fa = hf.read_fastq_arrays(_via_file(case["text"], tmp_path))
    got = [[n, s, q] for n, (s, q) in fa.records()]
    assert got == case["records"]


def test_fixture_file_both_readers():
    with gzip.open(os.path.join(GOLDEN, "sample_h1.fastq.gz"), "rb") as f:
        data = f.read()
    g = G["sample_h1"]
    fa = hf.parse_fastq_bytes(data)
    recs = [[n, s, q] for n, (s, q) in fa.records()]
    py = [[n, s, q] for n, (s, q) in hf.readfq(io.StringIO(data.decode()))]
    assert recs == py
    assert len(recs) == g["n"] and recs[0] == g["first"] and recs[-1] == g["last"]
    assert sum(len(r[1]) for r in recs) == g["total_seq"]
    assert hashlib.sha1("\n".join(r[0] for r in recs).encode()).hexdigest() == g["names_sha"]
    # array layout: offsets are the cumulative lengths, every record of this file has a quality
    assert fa.has_qual.all() and (np.diff(fa.seq_off) == np.diff(fa.qual_off)).all()
    assert fa.seq_off[-1] == len(fa.seq) == g["total_seq"]


def test_c_parser_random_texts_match_generator():
    """Random line soups (headers, '+', blank lines, all three newline styles, with and without a
    final newline): the C state machine and the Python generator agree record by record."""
    rng = np.random.default_rng(7)
    pieces = ["@r", ">s", "+", "+x", "ACGT", "AC", "", "IIII", "!!", "@", ">", "G" * 9, "@@@@", "+I+I"]
    for _ in range(400):
        n = int(rng.integers(0, 14))
        text = ""
        for _j in range(n):
            text += pieces[int(rng.integers(len(pieces)))] + ["\n", "\r\n", "\r"][int(rng.integers(3))]
        if n and rng.random() < 0.4:
            text = text.rstrip("\r\n")
        exp = [[nm, s, q] for nm, (s, q) in hf.readfq(io.StringIO(text, newline=None))]
        got = [[nm, s, q] for nm, (s, q) in hf.parse_fastq_bytes(text.encode()).records()]
        assert got == exp, repr(text)
