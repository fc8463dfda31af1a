// K4 (trace variant): semi-global affine alignment as a skewed-wavefront DP that writes 4 trace
// bits per cell, followed by a one-thread-per-pair traceback that turns the path into
//   (a) the k-column block statistic of cluster.parasail_block_alignment (modules/cluster.py:130-169),
//   (b) alignment identity = matches / columns (consensus.highest_aln_identity, consensus.py:129-145),
//   (c) per-window breaking points of the read on the target (racon's window layers).
// Same recurrence and tie-breaks as oracle/sg_align.c. Compared with the payload kernel of
// k4_align.cuh the DP needs ~3x fewer instructions per cell; the price is 0.5 byte/cell of trace
// that goes to HBM/L2 in fully coalesced 128-byte rows (one row per wavefront step and warp).
//
// Trace nibble of cell (i,j):  bits 0-1: source of H (0 diagonal match, 1 diagonal mismatch,
// 2 horizontal gap state D, 3 vertical gap state I); bit 2: D opened from H; bit 3: I opened from H.
// Word layout: trace[slot][pass][step t][lane] holds rows 8*lane..8*lane+7 of column j = t - lane.
#pragma once
#include "ngsid_internal.cuh"

#define K4T_RPL 8
#define K4T_MAXWIN 16

__host__ __device__ inline size_t k4t_smem_per_warp(int n2cap) { return (size_t)n2cap * 9; }
__host__ __device__ inline size_t k4t_trace_words(int n1, int n2)
{
    return (size_t)((n1 + 32 * K4T_RPL - 1) / (32 * K4T_RPL)) * (size_t)(n2 + 31) * 32;
}

struct K4TEnd { int32_t score, end_i, end_j, pad; };

// sequences come from the uploaded reads (index >= 0) or from a per-call auxiliary arena
// (index < 0 -> aux sequence -index-1), e.g. consensus targets
struct K4TSeqs {
    const uint8_t *seq; const int64_t *off;
    const uint8_t *aux; const int64_t *aoff;
    __device__ __forceinline__ const uint8_t *ptr(int r) const { return r >= 0 ? seq + off[r] : aux + aoff[-r - 1]; }
    __device__ __forceinline__ int len(int r) const {
        return r >= 0 ? (int)(off[r + 1] - off[r]) : (int)(aoff[-r] - aoff[-r - 1]);
    }
};

__global__ void __launch_bounds__(128)
k4t_dp_kernel(K4TSeqs Q,
              const int32_t *__restrict__ pa, const int32_t *__restrict__ pb,
              const int32_t *__restrict__ popen, int stride, int64_t pair0, int64_t n_pairs,
              int n2cap, uint32_t *__restrict__ trace, size_t slot_words, K4TEnd *__restrict__ ends)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int wid = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = (int)lane_id();
    int2 *bnd = reinterpret_cast<int2 *>(smem_raw + (size_t)wid * k4t_smem_per_warp(n2cap));
    uint8_t *s2s = reinterpret_cast<uint8_t *>(bnd + n2cap);

    for (int64_t sl = (int64_t)blockIdx.x * warps_per_block + wid; sl < n_pairs;
         sl += (int64_t)gridDim.x * warps_per_block) {
        const int64_t pr = pair0 + sl;
        const int ra = pa[pr * stride], rb = pb[pr * stride];
        const int open = popen[pr * stride];
        const uint8_t *s1 = Q.ptr(ra);
        const uint8_t *s2 = Q.ptr(rb);
        const int n1 = Q.len(ra);
        const int n2 = Q.len(rb);
        uint32_t *tr = trace + (size_t)sl * slot_words;
        for (int j = lane; j < n2; j += 32) s2s[j] = s2[j];
        __syncwarp();

        int bestv = NGSID_NEG_INF, besti = 0x7fffffff;
        int lrv = NGSID_NEG_INF, lrj = 0;
        const int rows_per_pass = 32 * K4T_RPL;
        const int npass = (n1 + rows_per_pass - 1) / rows_per_pass;
        const int nsteps = n2 + 31;
        for (int pass = 0; pass < npass; ++pass) {
            const int row0 = pass * rows_per_pass + lane * K4T_RPL;
            int H[K4T_RPL], D[K4T_RPL];
            uint32_t c1[K4T_RPL];
#pragma unroll
            for (int r = 0; r < K4T_RPL; ++r) {
                int i = row0 + r;
                c1[r] = (i < n1) ? (uint32_t)s1[i] : 0xffu;
                H[r] = 0;
                D[r] = NGSID_NEG_INF;
            }
            const int r_last = (n1 - 1) - row0;
            int oH = 0, oI = NGSID_NEG_INF;
            int dHp = 0;
            uint32_t *trp = tr + (size_t)pass * nsteps * 32 + lane;
            for (int t = 0; t < nsteps; ++t) {
                const int j = t - lane;
                int uH = __shfl_up_sync(NGSID_FULL_MASK, oH, 1);
                int uI = __shfl_up_sync(NGSID_FULL_MASK, oI, 1);
                if (j >= 0 && j < n2) {
                    if (lane == 0) {
                        if (pass == 0) { uH = 0; uI = NGSID_NEG_INF; }
                        else { int2 v = bnd[j]; uH = v.x; uI = v.y; }
                    }
                    const int sH = uH;
                    const uint32_t c2 = s2s[j];
                    int dH = dHp;
                    uint32_t word = 0;
#pragma unroll
                    for (int r = 0; r < K4T_RPL; ++r) {
                        const int io = uH - open, ie = uI - 1;
                        const int vI = max(io, ie);
                        const int dopn = H[r] - open, dext = D[r] - 1;
                        const int vD = max(dopn, dext);
                        const bool match = (c1[r] == c2);
                        const int hd = dH + (match ? 2 : -2);
                        const int h = max(hd, max(vD, vI));
                        uint32_t nib = (h == hd) ? (match ? 0u : 1u) : ((h == vD) ? 2u : 3u);
                        nib |= (dopn > dext) ? 4u : 0u;
                        nib |= (io > ie) ? 8u : 0u;
                        word |= nib << (4 * r);
                        dH = H[r];
                        H[r] = h; D[r] = vD;
                        uH = h; uI = vI;
                    }
                    trp[(size_t)t * 32] = word;
                    dHp = sH;
                    oH = uH; oI = uI;
                    if (j == n2 - 1) {
#pragma unroll
                        for (int r = 0; r < K4T_RPL; ++r)
                            if (row0 + r < n1 && H[r] > bestv) { bestv = H[r]; besti = row0 + r; }
                    }
                    if (r_last >= 0 && r_last < K4T_RPL) {
                        int hv = 0;
#pragma unroll
                        for (int r = 0; r < K4T_RPL; ++r) if (r == r_last) hv = H[r];
                        if (hv > lrv) { lrv = hv; lrj = j; }
                    }
                    if (lane == 31 && pass + 1 < npass) bnd[j] = make_int2(oH, oI);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
            int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int ov = __shfl_xor_sync(NGSID_FULL_MASK, lrv, d);
            int oj = __shfl_xor_sync(NGSID_FULL_MASK, lrj, d);
            if (ov > lrv) { lrv = ov; lrj = oj; }
        }
        if (lane == 0) {
            K4TEnd e;
            e.pad = 0;
            if (lrv > bestv) { e.score = lrv; e.end_i = n1 - 1; e.end_j = lrj; }
            else { e.score = bestv; e.end_i = besti; e.end_j = n2 - 1; }
            ends[sl] = e;
        }
        __syncwarp();
    }
}

// ---- traceback: one thread per pair -----------------------------------------------------------
struct K4TWindow { int32_t q_first, q_last, t_first, t_last; };   // q_last/t_last exclusive; -1 = none

struct K4TStat {
    uint32_t hist; int ncol; int cnt; uint32_t hm; int k, m;
    __device__ __forceinline__ void push(uint32_t bit) {
        hist = ((hist << 1) | bit) & hm;
        if (ncol < k) ncol++;
        if (ncol == k && __popc(hist) >= m) cnt++;
    }
};

// One WARP per pair. The path visits the 128-byte trace rows (one per wavefront step) in strictly
// descending order -- a diagonal or horizontal move goes to row t-1 of the same lane strip, a
// vertical move stays in the row or goes to row t-1 of the strip above -- so the warp streams the
// rows with coalesced loads, keeps K4T_PF rows in flight, and every lane follows the (uniform)
// state machine taking the word it needs by shuffle.
#define K4T_PF 8

__global__ void __launch_bounds__(128)
k4t_traceback_kernel(K4TSeqs Q, const int32_t *__restrict__ pa,
                     const int32_t *__restrict__ pb, const int32_t *__restrict__ pm, int stride,
                     int64_t pair0, int64_t n_pairs, int k, const uint32_t *__restrict__ trace,
                     size_t slot_words, const K4TEnd *__restrict__ ends,
                     int32_t *__restrict__ out_count, int32_t *__restrict__ out_score,
                     int32_t *__restrict__ out_match, int32_t *__restrict__ out_cols,
                     K4TWindow *__restrict__ out_win, int window)
{
    const int lane = (int)lane_id();
    const int64_t sl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sl >= n_pairs) return;
    const int64_t pr = pair0 + sl;
    const int ra = pa[pr * stride], rb = pb[pr * stride];
    const int n1 = Q.len(ra);
    const int n2 = Q.len(rb);
    const K4TEnd e = ends[sl];
    const uint32_t *tr = trace + (size_t)sl * slot_words;
    const int nsteps = n2 + 31;
    K4TStat S;
    S.hist = 0; S.ncol = 0; S.cnt = 0; S.k = k; S.m = pm ? pm[pr * stride] : 1;
    S.hm = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    int n_match = 0, n_cols = 0;
    // window breaking points: lane w keeps window w (K4T_MAXWIN <= 32)
    const int nwin = out_win ? min(K4T_MAXWIN, (n2 + window - 1) / window) : 0;
    K4TWindow mywin; mywin.q_first = mywin.q_last = mywin.t_first = mywin.t_last = -1;

    const int trailing = (n1 - 1 - e.end_i) + (n2 - 1 - e.end_j);
    for (int t = 0; t < trailing; ++t) S.push(0u);
    n_cols += trailing;
    int i = e.end_i, j = e.end_j, state = 0;            // 0 = H, 2 = D, 3 = I
    while (i >= 0 && j >= 0) {
        // rows of the current pass, from the current one downwards, K4T_PF at a time
        const int pass = i >> 8;
        const uint32_t *tp = tr + (size_t)pass * nsteps * 32;
        int t_cur = j + ((i & 255) >> 3);
        uint32_t buf[K4T_PF];
#pragma unroll
        for (int q = 0; q < K4T_PF; ++q) buf[q] = (t_cur - q >= 0) ? __ldcs(tp + (size_t)(t_cur - q) * 32 + lane) : 0u;
        int t_top = t_cur;                               // buf[q] holds row t_top - q
        while (i >= 0 && j >= 0 && (i >> 8) == pass) {
            const int strip = (i & 255) >> 3;
            const int t = j + strip;
            if (t_top - t >= K4T_PF) {                   // refill the window of rows
                t_top = t;
#pragma unroll
                for (int q = 0; q < K4T_PF; ++q) buf[q] = (t_top - q >= 0) ? __ldcs(tp + (size_t)(t_top - q) * 32 + lane) : 0u;
            }
            uint32_t mine = 0;
            const int sel = t_top - t;
#pragma unroll
            for (int q = 0; q < K4T_PF; ++q) if (q == sel) mine = buf[q];
            const uint32_t word = __shfl_sync(NGSID_FULL_MASK, mine, strip);
            const uint32_t nib = (word >> (4 * (i & 7))) & 15u;
            if (state == 0) {
                const uint32_t c = nib & 3u;
                if (c <= 1u) {
                    S.push(c == 0u ? 1u : 0u);
                    n_match += (c == 0u);
                    n_cols++;
                    if (nwin) {
                        const int w = j / window;
                        if (lane == w) {
                            if (mywin.q_last < 0) { mywin.q_last = i + 1; mywin.t_last = j + 1; }
                            mywin.q_first = i; mywin.t_first = j;
                        }
                    }
                    --i; --j;
                } else state = (int)c;
            } else if (state == 2) {
                S.push(0u); n_cols++;
                if (nib & 4u) state = 0;
                --j;
            } else {
                S.push(0u); n_cols++;
                if (nib & 8u) state = 0;
                --i;
            }
        }
    }
    const int leading = (i + 1) + (j + 1);
    for (int t = 0; t < leading; ++t) S.push(0u);
    n_cols += leading;
    int cnt = S.cnt;
    if (n_cols < k) cnt = (__popc(S.hist) >= S.m) ? 1 : 0;
    if (lane == 0) {
        if (out_count) out_count[pr] = cnt;
        if (out_score) out_score[pr] = e.score;
        if (out_match) out_match[pr] = n_match;
        if (out_cols) out_cols[pr] = n_cols;
    }
    if (out_win && lane < K4T_MAXWIN) {
        K4TWindow x; x.q_first = x.q_last = x.t_first = x.t_last = -1;
        out_win[pr * K4T_MAXWIN + lane] = (lane < nwin) ? mywin : x;
    }
}
