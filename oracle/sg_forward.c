/*
 * oracle/sg_forward.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Second, traceback-free restatement of the block statistic of modules/cluster.py:130-169:
 * every DP state carries the "payload" (match bits of the last k alignment columns, a
 * saturating column counter, number of good windows so far) of the path the traceback of
 * oracle/sg_align.c would follow to reach it. All tie-breaks are local to a cell, so the
 * payload at the end cell equals the statistic of the traced path. The CUDA kernel K4 is a
 * wavefront parallelisation of exactly this recurrence; tests check this file against
 * sg_align.c on random and adversarial pairs so the formulation is verified without a GPU.
 */
#include <stdint.h>
#include <stdlib.h>
#include <limits.h>

#define NEG_INF (INT_MIN / 2)

typedef struct { uint32_t hist; int ncol; int cnt; } payload_t;

static inline payload_t push(payload_t p, int bit, int k, int m)
{
    uint32_t mask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    p.hist = ((p.hist << 1) | (uint32_t)bit) & mask;
    if (p.ncol < k) p.ncol++;
    if (p.ncol == k && __builtin_popcount(p.hist) >= m) p.cnt++;
    return p;
}

static inline payload_t lead(int n, int k, int m)
{
    payload_t p = {0u, 0, 0};
    for (int i = 0; i < n; ++i) p = push(p, 0, k, m);
    return p;
}

int oracle_sg_block_forward(const char *s1, int n1, const char *s2, int n2,
                            int open, int ext, int k, int m, int *score_out)
{
    int *H = malloc(sizeof(int) * (n2 + 1)), *I = malloc(sizeof(int) * (n2 + 1));
    payload_t *PH = malloc(sizeof(payload_t) * (n2 + 1)), *PI = malloc(sizeof(payload_t) * (n2 + 1));
    if (!H || !I || !PH || !PI) return -1;
    for (int j = 0; j <= n2; ++j) { H[j] = 0; I[j] = NEG_INF; PH[j] = lead(j, k, m); PI[j] = PH[j]; }
    int best = NEG_INF, bi = -1, bj = -1; payload_t bp = {0u, 0, 0};
    for (int i = 1; i <= n1; ++i) {
        int diag = H[0], left = 0, D = NEG_INF;
        payload_t pdiag = PH[0], pleft = lead(i, k, m), pD = pleft;
        H[0] = 0; PH[0] = pleft;
        for (int j = 1; j <= n2; ++j) {
            int up = H[j]; payload_t pup = PH[j];
            int io = up - open, ie = I[j] - ext, vi; payload_t pi;
            if (io > ie) { vi = io; pi = push(pup, 0, k, m); } else { vi = ie; pi = push(PI[j], 0, k, m); }
            I[j] = vi; PI[j] = pi;
            int dopen = left - open, dext = D - ext;
            if (dopen > dext) { D = dopen; pD = push(pleft, 0, k, m); } else { D = dext; pD = push(pD, 0, k, m); }
            char a = s1[i - 1], b = s2[j - 1];
            int match = ((a == 'A' || a == 'C' || a == 'G' || a == 'T') && a == b);
            int hd = diag + (match ? 2 : -2);
            int h = hd; if (D > h) h = D; if (vi > h) h = vi;
            payload_t ph;
            if (h == hd) ph = push(pdiag, match, k, m); else if (h == D) ph = pD; else ph = pi;
            diag = up; pdiag = pup; left = h; pleft = ph; H[j] = h; PH[j] = ph;
        }
        if (H[n2] > best) { best = H[n2]; bi = i - 1; bj = n2 - 1; bp = PH[n2]; }
    }
    for (int j = 1; j <= n2; ++j) if (H[j] > best) { best = H[j]; bi = n1 - 1; bj = j - 1; bp = PH[j]; }
    int trailing = (n2 - 1 - bj) + (n1 - 1 - bi);
    int ncols_lt_k = (bp.ncol + trailing) < k;   /* ncol saturates at k, so this is exact */
    for (int t = 0; t < trailing; ++t) bp = push(bp, 0, k, m);
    int cnt = bp.cnt;
    if (ncols_lt_k) cnt = (__builtin_popcount(bp.hist) >= m) ? 1 : 0;
    if (score_out) *score_out = best;
    free(H); free(I); free(PH); free(PI);
    return cnt;
}
