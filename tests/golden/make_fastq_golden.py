#!/usr/bin/env python
"""Golden vectors for the FASTA/FASTQ reader (SURVEY.md 8 f rank 3): the REFERENCE's own
modules/help_functions.readfq (imported from /root/reference, runs in the build container only) on
crafted edge cases and on the committed fixture. Output: tests/golden/readfq.json.gz
    python tests/golden/make_fastq_golden.py
"""
import gzip
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from modules import help_functions as ref  # noqa: E402

CASES = {
    "plain4": "@r1 a b\nACGT\n+\nIIII\n@r2\nGGCC\n+r2\n!!!!\n",
    "no_final_newline": "@r1\nACGT\n+\nIIII\n@r2\nGGCCA\n+\n!!!!!",
    "multiline_fastq": "@r1\nACGT\nAC\n+\nIIII\nII\n@r2\nGG\n+\n!!\n",
    "quality_starts_with_at": "@r1\nACGT\n+\n@III\n@r2\nGG\n+\n@@\n",
    "fasta": ">s1 desc\nACGT\nACGT\n>s2\nGG\n",
    "fasta_then_fastq": ">s1\nACGT\n@r1\nAC\n+\nII\n",
    "fastq_then_fasta": "@r1\nAC\n+\nII\n>s1\nACGT\n",
    "blank_lines": "\n\n@r1\nACGT\n\n+\nIIII\n\n@r2\nGG\n+\n!!\n\n",
    "leading_junk": "junk line\nmore junk\n@r1\nACGT\n+\nIIII\n",
    "truncated_quality": "@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\n!!\n",
    "missing_quality": "@r1\nACGT\n+\n",
    "empty": "",
    "only_header": "@r1\n",
    "empty_sequence": "@r1\n\n+\n\n@r2\nAC\n+\nII\n",
    "crlf": "@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\n!!\r\n",
    "lone_cr": "@r1\rACGT\r+\rIIII\r@r2\rGG\r+\r!!\r",
    "mixed_newlines": "@r1\r\nACGT\n+\rIIII\r\n@r2\nGG\r\n+\n!!",
    "long_quality_line": "@r1\nAC\n+\nIIIIII\n@r2\nGG\n+\n!!\n",
    "plus_in_name": "@r+1 x\nACGT\n+r+1 x\nII+I\n",
    "header_only_at": "@\nACGT\n+\nIIII\n",
    "gt_inside_fastq": "@r1\nAC\n>weird\nGG\n+\nII\n",
    "two_plus_lines": "@r1\nAC\n+\n+I\n@r2\nG\n+\n!\n",
}


def run(text):
    """Through a real file opened the way the reference opens its input (text mode 'r':
    universal newlines), get_sorted_fastq_for_cluster.py:126, NGSpeciesID:54."""
    import tempfile
    with tempfile.NamedTemporaryFile("wb", suffix=".fq", delete=False) as f:
        f.write(text.encode("ascii"))
        path = f.name
    try:
        with open(path, "r") as fp:
            return [[name, seq, qual] for name, (seq, qual) in ref.readfq(fp)]
    finally:
        os.unlink(path)


def main():
    out = {"cases": []}
    for tag, text in CASES.items():
        out["cases"].append({"tag": tag, "text": text, "records": run(text)})
    with gzip.open(os.path.join(HERE, "sample_h1.fastq.gz"), "rt") as f:
        text = f.read()
    recs = run(text)
    out["sample_h1"] = {"n": len(recs), "first": recs[0], "last": recs[-1],
                        "total_seq": sum(len(r[1]) for r in recs),
                        "names_sha": __import__("hashlib").sha1("\n".join(r[0] for r in recs).encode()).hexdigest()}
    with gzip.open(os.path.join(HERE, "readfq.json.gz"), "wt") as f:
        json.dump(out, f)
    print(len(out["cases"]), "cases;", out["sample_h1"]["n"], "fixture records")


if __name__ == "__main__":
    main()
