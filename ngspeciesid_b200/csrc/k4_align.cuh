// K4: semi-global affine alignment + k-column block statistic, without a trace table.
// Reference semantics: cluster.parasail_block_alignment (modules/cluster.py:130-169) =
// parasail sg_trace_scan_16(match 2, mismatch -2, open o, extend 1) + CIGAR expansion
// (help_functions.py:56-97) + sliding window of k alignment columns counting windows with
// >= match_id matches. Tie-breaks are those of oracle/sg_align.c (documented there).
//
// Formulation (verified on the CPU by oracle/sg_forward.c against the traceback oracle): every
// DP state carries a 32-bit payload describing the path the traceback would follow to reach it:
//   bits  0..14  match bits of the last k columns (or, while fewer than k columns exist, the
//                bits so far with a sentinel 1 above them)
//   bit   15     FULL: at least k columns seen
//   bits 16..31  number of good windows so far
// Because every tie-break is local to a cell, the payload of the end cell is the statistic of
// the traced path; no trace memory is written and no traceback runs.
//
// Parallelisation: one warp per pair. Lane l owns RPL consecutive rows of s1 and sweeps the
// columns of s2 one step behind lane l-1 (skewed wavefront); the bottom row of each lane is
// handed to the next lane by shuffle. Reads longer than 32*RPL rows take several passes; the
// bottom row of a pass waits in shared memory for the next pass.
#pragma once
#include "ngsid_internal.cuh"

#define K4_RPL 8
#define K4_FULL 0x8000u

struct K4Const { uint32_t hm; int k; int m; };

__device__ __forceinline__ uint32_t k4_push_fast(uint32_t p, uint32_t bit, const K4Const &c)
{
    uint32_t h = ((p << 1) | bit) & c.hm;
    uint32_t good = (__popc(h) >= c.m) ? 0x10000u : 0u;
    return ((p & 0xffff8000u) | h) + good;
}

__device__ __forceinline__ uint32_t k4_push_slow(uint32_t p, uint32_t bit, const K4Const &c)
{
    if (p & K4_FULL) return k4_push_fast(p, bit, c);
    uint32_t s = ((p & 0x7fffu) << 1) | bit;
    if ((s >> c.k) & 1u) {                       // the sentinel reached bit k: k columns seen
        uint32_t h = s & c.hm;
        uint32_t good = (__popc(h) >= c.m) ? 0x10000u : 0u;
        return ((p & 0xffff0000u) | K4_FULL | h) + good;
    }
    return (p & 0xffff0000u) | s;
}

// payload of a path that starts with n gap columns
__device__ __forceinline__ uint32_t k4_lead(int n, const K4Const &c)
{
    if (n < c.k) return 1u << n;
    uint32_t cnt = (c.m <= 0) ? (uint32_t)(n - c.k + 1) : 0u;
    return (cnt << 16) | K4_FULL;
}

template <bool SLOW>
__device__ __forceinline__ uint32_t k4_push(uint32_t p, uint32_t bit, const K4Const &c)
{
    return SLOW ? k4_push_slow(p, bit, c) : k4_push_fast(p, bit, c);
}

struct K4Rows {
    int H[K4_RPL], D[K4_RPL];
    uint32_t PH[K4_RPL], PD[K4_RPL];
    uint32_t c1[K4_RPL];
};

// One column step over the RPL rows of this lane. In: up values (row above the strip) and the
// diagonal value; out: the strip's bottom-row values for the lane below.
template <bool SLOW>
__device__ __forceinline__ void k4_column(K4Rows &R, uint32_t c2, int open, const K4Const &kc,
                                          int &uH, int &uI, uint32_t &uPH, uint32_t &uPI,
                                          int dH, uint32_t dP)
{
#pragma unroll
    for (int r = 0; r < K4_RPL; ++r) {
        // vertical gap state (consumes s1): opened from H above or extended
        int io = uH - open, ie = uI - 1;
        bool iopen = io > ie;
        int vI = iopen ? io : ie;
        uint32_t pI = k4_push<SLOW>(iopen ? uPH : uPI, 0u, kc);
        // horizontal gap state (consumes s2)
        int dopen = R.H[r] - open, dext = R.D[r] - 1;
        bool dop = dopen > dext;
        int vD = dop ? dopen : dext;
        uint32_t pD = k4_push<SLOW>(dop ? R.PH[r] : R.PD[r], 0u, kc);
        // diagonal
        bool match = (R.c1[r] == c2);
        int hd = dH + (match ? 2 : -2);
        int h = max(hd, max(vD, vI));
        uint32_t pH = (h == hd) ? k4_push<SLOW>(dP, match ? 1u : 0u, kc) : ((h == vD) ? pD : pI);
        // rotate: this row's old left value is the next row's diagonal
        dH = R.H[r]; dP = R.PH[r];
        R.H[r] = h; R.D[r] = vD; R.PH[r] = pH; R.PD[r] = pD;
        uH = h; uI = vI; uPH = pH; uPI = pI;
    }
}

__host__ __device__ inline size_t k4_smem_per_warp(int n2cap) { return (size_t)n2cap * 17; }

struct K4Pair { int32_t a, b, open, m; };

__global__ void __launch_bounds__(128)
k4_align_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ off,
                const int32_t *__restrict__ pa, const int32_t *__restrict__ pb,
                const int32_t *__restrict__ popen, const int32_t *__restrict__ pm,
                int stride, int64_t n_pairs, int k, int n2cap,
                int32_t *__restrict__ out_count, int32_t *__restrict__ out_score)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int wid = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = (int)lane_id();
    int4 *bnd = reinterpret_cast<int4 *>(smem_raw + (size_t)wid * k4_smem_per_warp(n2cap));
    uint8_t *s2s = reinterpret_cast<uint8_t *>(bnd + n2cap);

    for (int64_t pr = (int64_t)blockIdx.x * warps_per_block + wid; pr < n_pairs;
         pr += (int64_t)gridDim.x * warps_per_block) {
        const int ra = pa[pr * stride], rb = pb[pr * stride];
        const int open = popen[pr * stride];
        K4Const kc;
        kc.k = k; kc.m = pm[pr * stride]; kc.hm = (1u << k) - 1u;
        const uint8_t *s1 = seq + off[ra];
        const uint8_t *s2 = seq + off[rb];
        const int n1 = (int)(off[ra + 1] - off[ra]);
        const int n2 = (int)(off[rb + 1] - off[rb]);
        for (int j = lane; j < n2; j += 32) s2s[j] = k4_col_base(s2[j]);
        __syncwarp();

        // best cell of the last column (rows ascending, strictly greater replaces)
        int bestv = NGSID_NEG_INF, besti = 0x7fffffff; uint32_t bestP = 0;
        // best cell of the last row (columns ascending, strictly greater replaces)
        int lrv = NGSID_NEG_INF, lrj = 0; uint32_t lrP = 0;

        const int rows_per_pass = 32 * K4_RPL;
        const int npass = (n1 + rows_per_pass - 1) / rows_per_pass;
        for (int pass = 0; pass < npass; ++pass) {
            const int row0 = pass * rows_per_pass + lane * K4_RPL;   // 0-based first row
            K4Rows R;
#pragma unroll
            for (int r = 0; r < K4_RPL; ++r) {
                int i = row0 + r;
                R.c1[r] = (i < n1) ? k4_row_base(s1[i]) : 0xffu;
                R.H[r] = 0;
                R.D[r] = NGSID_NEG_INF;
                R.PH[r] = k4_lead(i + 1, kc);
                R.PD[r] = 0;
            }
            const int r_last = (n1 - 1) - row0;       // row slot holding the last row, if any
            int oH = 0, oI = NGSID_NEG_INF; uint32_t oPH = 0, oPI = 0;
            int dHp = 0; uint32_t dPp = k4_lead(row0, kc);
            const bool lane_slow = row0 < k;
            const int nsteps = n2 + 31;
            for (int t = 0; t < nsteps; ++t) {
                const int j = t - lane;
                int uH = __shfl_up_sync(NGSID_FULL_MASK, oH, 1);
                int uI = __shfl_up_sync(NGSID_FULL_MASK, oI, 1);
                uint32_t uPH = __shfl_up_sync(NGSID_FULL_MASK, oPH, 1);
                uint32_t uPI = __shfl_up_sync(NGSID_FULL_MASK, oPI, 1);
                const bool active = (j >= 0) && (j < n2);
                if (active) {
                    if (lane == 0) {
                        if (pass == 0) {
                            uH = 0; uI = NGSID_NEG_INF; uPH = k4_lead(j + 1, kc); uPI = 0;
                        } else {
                            int4 v = bnd[j];
                            uH = v.x; uI = v.y; uPH = (uint32_t)v.z; uPI = (uint32_t)v.w;
                        }
                    }
                    const int sH = uH; const uint32_t sP = uPH;       // next column's diagonal
                    const uint32_t c2 = s2s[j];
                    if (lane_slow && j < k) k4_column<true>(R, c2, open, kc, uH, uI, uPH, uPI, dHp, dPp);
                    else k4_column<false>(R, c2, open, kc, uH, uI, uPH, uPI, dHp, dPp);
                    dHp = sH; dPp = sP;
                    oH = uH; oI = uI; oPH = uPH; oPI = uPI;
                    if (j == n2 - 1) {
#pragma unroll
                        for (int r = 0; r < K4_RPL; ++r)
                            if (row0 + r < n1 && R.H[r] > bestv) { bestv = R.H[r]; besti = row0 + r; bestP = R.PH[r]; }
                    }
                    if (r_last >= 0 && r_last < K4_RPL) {
                        int hv = 0; uint32_t hp = 0;
#pragma unroll
                        for (int r = 0; r < K4_RPL; ++r) if (r == r_last) { hv = R.H[r]; hp = R.PH[r]; }
                        if (hv > lrv) { lrv = hv; lrj = j; lrP = hp; }
                    }
                    if (lane == 31 && pass + 1 < npass) bnd[j] = make_int4(oH, oI, (int)oPH, (int)oPI);
                }
            }
            __syncwarp();
        }
        // ---- end cell: last column first (smallest row on ties), then last row if strictly better
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
            int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
            uint32_t op = __shfl_xor_sync(NGSID_FULL_MASK, bestP, d);
            if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; bestP = op; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {   // only one lane holds a real last-row value
            int ov = __shfl_xor_sync(NGSID_FULL_MASK, lrv, d);
            int oj = __shfl_xor_sync(NGSID_FULL_MASK, lrj, d);
            uint32_t op = __shfl_xor_sync(NGSID_FULL_MASK, lrP, d);
            if (ov > lrv) { lrv = ov; lrj = oj; lrP = op; }
        }
        if (lane == 0) {
            int score = bestv, trailing = (n1 - 1) - besti; uint32_t P = bestP;
            if (lrv > bestv) { score = lrv; trailing = (n2 - 1) - lrj; P = lrP; }
            int npush = min(trailing, 2 * k + 2);
            for (int t = 0; t < npush; ++t) P = k4_push_slow(P, 0u, kc);
            int cnt = (int)(P >> 16);
            if (kc.m <= 0 && (P & K4_FULL)) cnt += trailing - npush;   // every further column counts
            if (!(P & K4_FULL)) {
                // fewer than k columns in total: a single window (cluster.py:147-153)
                uint32_t s = P & 0x7fffu;
                int ones = __popc(s) - 1;           // minus the sentinel
                cnt = (ones >= kc.m) ? 1 : 0;
            }
            out_count[pr] = cnt;
            if (out_score) out_score[pr] = score;
        }
        __syncwarp();
    }
}
