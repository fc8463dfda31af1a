// FASTA/FASTQ ingest on the host (SURVEY.md section 8 f, rank 3): the record grammar of the
// reference's reader (modules/help_functions.py:13-42, lh3's readfq generator) over a whole file
// held in memory, written as a line state machine instead of a generator. Sequences and qualities
// come out concatenated (the layout ngsid_upload_reads takes), names as spans of the input buffer.
//
// Line model = Python text mode 'r' (how the reference opens its input,
// get_sorted_fastq_for_cluster.py:126): "\n", "\r\n" and a lone "\r" all end a line. The
// generator strips the last character of every line without looking at it (`l[:-1]`), so the last
// line of a file that does not end with a newline loses its last character, and a line counts
// len - 1 quality characters; both quirks are kept (golden vectors: tests/golden/readfq.json.gz).
#pragma once
#include <stdint.h>
#include <string.h>

namespace fastq_ingest {

struct Line {
    int64_t s, e;        // what the generator sees after l[:-1]
    uint8_t first;       // l[0] ('\n' for an empty line)
};

struct Reader {
    const uint8_t *b; int64_t n, p; bool has_cr;
    bool next(Line &L)
    {
        if (p >= n) return false;
        const int64_t s = p;
        int64_t q;
        if (!has_cr) {
            const void *f = memchr(b + p, '\n', (size_t)(n - p));
            q = f ? (int64_t)((const uint8_t *)f - b) : n;
        } else {
            q = p;
            while (q < n && b[q] != '\n' && b[q] != '\r') ++q;
        }
        if (q >= n) { L.s = s; L.e = n - 1; p = n; }                      // no terminator: l[:-1] drops a character
        else { L.s = s; L.e = q; p = q + 1; if (b[q] == '\r' && p < n && b[p] == '\n') ++p; }
        L.first = (q > s) ? b[s] : (uint8_t)'\n';
        return true;
    }
};

struct Out {
    int64_t cap, n;
    uint8_t *seq, *qual;
    int64_t *name_off; int32_t *name_len; int64_t *seq_off, *qual_off; uint8_t *has_qual;
    int64_t sp, qp;      // bytes written
};

// Returns the number of records in the buffer; writes the first `cap` of them.
static inline int64_t parse(const uint8_t *buf, int64_t len, Out &O)
{
    Reader R = {buf, len, 0, memchr(buf, '\r', (size_t)len) != nullptr};
    Line L, last = {0, 0, 0};
    bool have_last = false;
    O.n = 0; O.sp = 0; O.qp = 0;
    const bool wr = O.seq != nullptr;
    while (true) {
        if (!have_last) {
            while (R.next(L))
                if (L.first == '>' || L.first == '@') { last = L; have_last = L.e > L.s; break; }
        }
        if (!have_last) break;
        const int64_t rec = O.n;
        const bool store = wr && rec < O.cap;
        if (store) { O.name_off[rec] = last.s + 1; O.name_len[rec] = (int32_t)(last.e - last.s - 1); O.seq_off[rec] = O.sp; O.qual_off[rec] = O.qp; }
        have_last = false;
        int64_t seq_len = 0;
        while (R.next(L)) {
            if (L.first == '@' || L.first == '+' || L.first == '>') { last = L; have_last = L.e > L.s; break; }
            if (store) memcpy(O.seq + O.sp + seq_len, buf + L.s, (size_t)(L.e - L.s));
            seq_len += L.e - L.s;
        }
        if (store) O.sp += seq_len;
        if (!have_last || buf[last.s] != '+') {          // FASTA record
            if (store) O.has_qual[rec] = 0;
            O.n++;
            if (!have_last) break;
            continue;
        }
        int64_t got = 0;
        bool done = false;
        while (R.next(L)) {
            if (store) memcpy(O.qual + O.qp + got, buf + L.s, (size_t)(L.e - L.s));
            got += L.e - L.s;
            if (got >= seq_len) { done = true; break; }
        }
        have_last = false;
        if (store) { O.has_qual[rec] = done ? 1 : 0; if (done) O.qp += got; }
        O.n++;
        if (!done) break;                                 // end of file inside the quality: FASTA record, stop
    }
    if (wr) { const int64_t m = O.n < O.cap ? O.n : O.cap; O.seq_off[m] = O.sp; O.qual_off[m] = O.qp; }
    return O.n;
}

}  // namespace fastq_ingest
