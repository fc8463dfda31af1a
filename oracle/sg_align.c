/*
 * oracle/sg_align.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the semi-global affine alignment the reference obtains from the
 * third-party package parasail==1.2.4 (pinned in /root/reference/setup.py:78), which is
 * NOT vendored under /root/reference and is not installable here (no network).
 *
 * Reference call sites this stands in for:
 *   modules/cluster.py:131-135    parasail.matrix_create("ACGT", 2, -2);
 *                                 parasail.sg_trace_scan_16(s1, s2, open, ext, M) (+ _32 retry)
 *   modules/consensus.py:59-63    same call, open=3
 *   modules/cluster.py:138-168    CIGAR -> gapped strings -> k-window block statistic
 *   modules/help_functions.py:56-97  cigar_to_seq ('=','X' both; 'I' query only; 'D' ref only)
 *
 * PARITY UNPINNED: parasail's published algorithm (Daily 2016; serial `sg_trace` is the
 * specification its vectorised variants are verified against) is restated from its
 * documentation. No golden vector for parasail output exists in the reference (it has no
 * tests), so tie-breaking below is a documented choice, not a verified fact:
 *
 *   - s1 = query = DP rows i, s2 = reference = DP columns j.
 *   - free end gaps on both ends of both sequences: H[0][j] = H[i][0] = 0; the end cell is the
 *     best cell of the last column (scanned i ascending, strictly-greater replaces) and then
 *     the last row (j ascending, strictly-greater replaces).
 *   - a gap of length n costs open + (n-1)*ext.
 *   - 'D' state (consumes s2, horizontal):  D[i][j] = max(H[i][j-1]-open, D[i][j-1]-ext),
 *     'I' state (consumes s1, vertical):    I[i][j] = max(H[i-1][j]-open, I[i-1][j]-ext);
 *     a gap state records "opened from H" only when the open branch is STRICTLY better
 *     (ties extend).
 *   - H[i][j] = max(H[i-1][j-1] + sub, D, I); on ties H prefers the diagonal, then D, then I.
 *   - traceback starts in the H state at the end cell; when a sequence is exhausted the
 *     remainder of the other is emitted as end gaps; end gaps after the end cell are
 *     emitted too, so the CIGAR always covers both sequences completely.
 *   - characters outside "ACGT" score as mismatch against everything (including themselves).
 *
 * The CUDA kernel (ngspeciesid_b200/csrc) must reproduce exactly these choices; the tests
 * compare it with this file, and the golden fixtures under tests/golden were produced by
 * the reference's own Python driven by this aligner through oracle/parasail_shim.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define NEG_INF (INT_MIN / 2)

enum { T_DIAG = 1, T_D = 2, T_I = 4, T_D_OPEN = 8, T_I_OPEN = 16 };

/*
 * Tie-break variants for the sensitivity study (tests/test_aligner_pinning.py, scripts/tiebreak_sensitivity.py):
 * bit 0: H prefers I over D on ties (default D over I); bit 1: a gap state tied between open and extend
 * records "opened" (default: extend); bit 2: the end cell scans the last row before the last column
 * (default: column first); bit 3: H prefers the gap states over the diagonal on ties.
 * 0 = the documented choices above. Every variant is an optimal alignment: only co-optimal paths differ.
 */
static int g_tiebreak = 0;
void oracle_sg_set_tiebreak(int v) { g_tiebreak = v; }
int oracle_sg_get_tiebreak(void) { return g_tiebreak; }

static inline int sub_score(char a, char b, int match, int mismatch)
{
    int ok = (a == 'A' || a == 'C' || a == 'G' || a == 'T');
    return (ok && a == b) ? match : mismatch;
}

/*
 * Align s1 (rows) against s2 (columns). Writes the expanded CIGAR (one op char per alignment
 * column, in forward order, ops "=XID") into ops_out (capacity >= n1+n2+1) and returns the
 * number of columns; *score_out receives the alignment score. Returns -1 on allocation failure.
 */
int oracle_sg_align(const char *s1, int n1, const char *s2, int n2,
                    int match, int mismatch, int open, int ext,
                    char *ops_out, int *score_out, int *end_i_out, int *end_j_out)
{
    if (n1 <= 0 || n2 <= 0) {
        int c = 0;
        for (int i = 0; i < n1; ++i) ops_out[c++] = 'I';
        for (int j = 0; j < n2; ++j) ops_out[c++] = 'D';
        if (score_out) *score_out = 0;
        if (end_i_out) *end_i_out = n1 - 1;
        if (end_j_out) *end_j_out = n2 - 1;
        return c;
    }
    uint8_t *trace = (uint8_t *)malloc((size_t)n1 * (size_t)n2);
    int *H = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    int *I = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    if (!trace || !H || !I) { free(trace); free(H); free(I); return -1; }

    for (int j = 0; j <= n2; ++j) { H[j] = 0; I[j] = NEG_INF; }
    int best = NEG_INF, bi = -1, bj = -1;

    for (int i = 1; i <= n1; ++i) {
        int diag = H[0];      /* H[i-1][0] */
        int left = 0;         /* H[i][0]   */
        int D = NEG_INF;      /* D[i][0]   */
        H[0] = 0;
        uint8_t *trow = trace + (size_t)(i - 1) * (size_t)n2;
        for (int j = 1; j <= n2; ++j) {
            int up = H[j];    /* H[i-1][j] */
            uint8_t t = 0;
            int i_open = up - open, i_ext = I[j] - ext;
            int vi;
            const int open_on_tie = (g_tiebreak & 2) != 0;
            if (i_open > i_ext || (open_on_tie && i_open == i_ext)) { vi = i_open; t |= T_I_OPEN; } else vi = i_ext;
            I[j] = vi;
            int d_open = left - open, d_ext = D - ext;
            if (d_open > d_ext || (open_on_tie && d_open == d_ext)) { D = d_open; t |= T_D_OPEN; } else D = d_ext;
            int hd = diag + sub_score(s1[i - 1], s2[j - 1], match, mismatch);
            int h = hd;
            if (D > h) h = D;
            if (vi > h) h = vi;
            if (g_tiebreak & 8) {
                if ((g_tiebreak & 1) ? (h == vi) : (h == D)) t |= (g_tiebreak & 1) ? T_I : T_D;
                else if ((g_tiebreak & 1) ? (h == D) : (h == vi)) t |= (g_tiebreak & 1) ? T_D : T_I;
                else t |= T_DIAG;
            } else if (h == hd) t |= T_DIAG;
            else if (g_tiebreak & 1) { if (h == vi) t |= T_I; else t |= T_D; }
            else { if (h == D) t |= T_D; else t |= T_I; }
            trow[j - 1] = t;
            diag = up;
            left = h;
            H[j] = h;
        }
        /* last column of this row */
        if (!(g_tiebreak & 4) && H[n2] > best) { best = H[n2]; bi = i - 1; bj = n2 - 1; }
        if ((g_tiebreak & 4) && i < n1 && H[n2] > best) { best = H[n2]; bi = i - 1; bj = n2 - 1; }     /* provisional */
    }
    if (g_tiebreak & 4) {
        /* last row first (j ascending), then the last column (i ascending), strictly greater replaces */
        int b2 = NEG_INF, i2 = -1, j2 = -1;
        for (int j = 1; j <= n2; ++j)
            if (H[j] > b2) { b2 = H[j]; i2 = n1 - 1; j2 = j - 1; }
        if (best > b2) { /* a last-column cell of an earlier row is strictly better */ }
        else { best = b2; bi = i2; bj = j2; }
    } else {
        for (int j = 1; j <= n2; ++j)
            if (H[j] > best) { best = H[j]; bi = n1 - 1; bj = j - 1; }
    }

    /* traceback, ops collected in reverse */
    char *rev = (char *)malloc((size_t)n1 + (size_t)n2 + 2);
    if (!rev) { free(trace); free(H); free(I); return -1; }
    int c = 0;
    for (int j = n2 - 1; j > bj; --j) rev[c++] = 'D';
    for (int i = n1 - 1; i > bi; --i) rev[c++] = 'I';
    int i = bi, j = bj, where = T_DIAG;
    while (i >= 0 || j >= 0) {
        if (i < 0) { rev[c++] = 'D'; --j; continue; }
        if (j < 0) { rev[c++] = 'I'; --i; continue; }
        uint8_t t = trace[(size_t)i * (size_t)n2 + (size_t)j];
        if (where == T_DIAG) {
            if (t & T_DIAG) {
                /* '=' by character equality (what the reference's statistic compares, modules/cluster.py:147);
                   the SCORE of a base outside ACGT is a mismatch even against itself (sub_score) */
                rev[c++] = (s1[i] == s2[j]) ? '=' : 'X';
                --i; --j;
            } else if (t & T_D) where = T_D;
            else where = T_I;
        } else if (where == T_D) {
            rev[c++] = 'D';
            if (t & T_D_OPEN) where = T_DIAG;
            --j;
        } else {
            rev[c++] = 'I';
            if (t & T_I_OPEN) where = T_DIAG;
            --i;
        }
    }
    for (int x = 0; x < c; ++x) ops_out[x] = rev[c - 1 - x];
    if (score_out) *score_out = best;
    if (end_i_out) *end_i_out = bi;
    if (end_j_out) *end_j_out = bj;
    free(rev); free(trace); free(H); free(I);
    return c;
}

/*
 * Block statistic of modules/cluster.py:146-168 computed from the expanded ops:
 * match bit per column ('=' -> 1; 'X','I','D' -> 0; note the reference compares the gapped
 * strings, and a mismatch/gap column is never equal); first window = first min(k, n) columns,
 * then one window per further column; returns the number of windows holding >= match_id matches.
 */
int oracle_block_count(const char *ops, int n_cols, int k, int match_id)
{
    int cur = 0, count = 0;
    int first = n_cols < k ? n_cols : k;
    for (int c = 0; c < first; ++c) cur += (ops[c] == '=');
    if (cur >= match_id) ++count;
    for (int c = k; c < n_cols; ++c) {
        cur += (ops[c] == '=') - (ops[c - k] == '=');
        if (cur >= match_id) ++count;
    }
    return count;
}

/* Convenience: alignment + statistic in one call (what one K4 work item produces). */
int oracle_sg_block_align(const char *s1, int n1, const char *s2, int n2,
                          int open, int ext, int k, int match_id,
                          int *n_cols_out, int *score_out)
{
    char *ops = (char *)malloc((size_t)n1 + (size_t)n2 + 2);
    if (!ops) return -1;
    int n = oracle_sg_align(s1, n1, s2, n2, 2, -2, open, ext, ops, score_out, 0, 0);
    if (n < 0) { free(ops); return -1; }
    int cnt = oracle_block_count(ops, n, k, match_id);
    if (n_cols_out) *n_cols_out = n;
    free(ops);
    return cnt;
}
