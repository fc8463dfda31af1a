// K4 (trace variant): semi-global affine alignment as a skewed-wavefront DP that writes 4 trace
// bits per cell, followed by a warp-per-pair traceback that turns the path into
//   (a) the k-column block statistic of cluster.parasail_block_alignment (modules/cluster.py:130-169),
//   (b) alignment identity = matches / columns (consensus.highest_aln_identity, consensus.py:129-145),
//   (c) per-window breaking points of the read on the target (racon's window layers).
// Same recurrence and tie-breaks as oracle/sg_align.c. The price of the short DP inner loop is
// 0.5 byte/cell of trace that goes to HBM/L2 in fully coalesced 128-byte rows (one row per
// wavefront step and strip of 256 rows).
//
// Trace nibble of cell (i,j):  bit 0 (a): H is not the diagonal; bit 1 (b): H is not the
// horizontal gap state D  (source of H = diagonal if !a, else D if !b, else the vertical gap state
// I -- the oracle's tie order; the traceback re-reads the two bases to tell match from mismatch);
// bit 2: D opened from H; bit 3: I opened from H.
// Word layout: trace[slot][pass][step t][lane] holds rows 8*lane..8*lane+7 of column j = t - lane.
//
// The DP kernel has two shapes. MULTI = false: one warp per pair, the strips of 256 rows ("passes")
// one after the other -- throughput shape for thousands of pairs. MULTI = true: one thread block
// per pair, warp w takes passes w, w+W, ... and starts a column as soon as the warp above has
// published its bottom row for it (shared-memory row + progress counter), so a single pair takes
// (n2 + 31 + 32 (W-1)) steps instead of npass * (n2 + 31) -- latency shape for the small rounds of
// the greedy clustering pass.
#pragma once
#include "ngsid_internal.cuh"

#define K4T_RPL 8
#define K4T_MAXWIN 16
#define K4T_MAXW 4                   // warps per pair in the MULTI shape

__host__ __device__ inline size_t k4t_smem_per_warp(int n2cap) { return (size_t)n2cap * 9; }
// MULTI shape, per block: W boundary rows (int2 per column) | target bases | progress + end records
__host__ __device__ inline size_t k4t_smem_multi(int n2cap, int W) { return (size_t)n2cap * (8 * W + 1) + 128; }
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

// One cell. H/D are the lane's row state, (uH, uI) come from the row above, dH is the diagonal.
// Written for the ALU pipe (half rate for compare / select / min-max, full rate for plain adds):
// fused add-max for the two gap states, one 3-input maximum for H, and every trace bit is ONE
// compare (result != the operand that loses ties) plus one predicated add into the trace word.
#define K4T_CELL(r)                                                                   \
    {                                                                                 \
        const int ie = uI - 1;                                                        \
        const int vI = __viaddmax_s32(uH, nopen, ie);             /* max(uH - open, uI - 1) */ \
        const int dext = D[r] - 1;                                                    \
        const int vD = __viaddmax_s32(H[r], nopen, dext);                             \
        int hd = dH + 2;                                                              \
        if (c1[r] != c2) hd = dH - 2;                                                 \
        const int h = __vimax3_s32(hd, vD, vI);                                       \
        if (h != hd) word += 1u << (4 * r);                       /* a: H is not the diagonal */ \
        if (h != vD) word += 2u << (4 * r);                       /* b: H is not D */ \
        if (vD != dext) word += 4u << (4 * r);                    /* D opened from H (open > extend) */ \
        if (vI != ie) word += 8u << (4 * r);                      /* I opened from H */ \
        dH = H[r];                                                                    \
        H[r] = h; D[r] = vD;                                                          \
        uH = h; uI = vI;                                                              \
    }

template <bool MULTI>
__global__ void __launch_bounds__(128)
k4t_dp_kernel(K4TSeqs Q,
              const int32_t *__restrict__ pa, const int32_t *__restrict__ pb,
              const int32_t *__restrict__ popen, int stride, int64_t pair0, int64_t n_pairs,
              int n2cap, uint32_t *__restrict__ trace, size_t slot_words, K4TEnd *__restrict__ ends)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int wid = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int lane = (int)lane_id();
    // MULTI: bnd[w] is the bottom row published by the warp running pass p with p % W == w
    int2 *bnd_base = MULTI ? reinterpret_cast<int2 *>(smem_raw)
                           : reinterpret_cast<int2 *>(smem_raw + (size_t)wid * k4t_smem_per_warp(n2cap));
    uint8_t *s2s = MULTI ? smem_raw + (size_t)n2cap * 8 * W : reinterpret_cast<uint8_t *>(bnd_base + n2cap);
    volatile int *prog = reinterpret_cast<volatile int *>(smem_raw + (size_t)n2cap * (8 * W + 1));   // MULTI only
    int *endbuf = const_cast<int *>(prog) + 8;                                                        // 4 ints per warp

    const int64_t first = MULTI ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * W + wid;
    const int64_t step_sl = MULTI ? (int64_t)gridDim.x : (int64_t)gridDim.x * W;
    for (int64_t sl = first; sl < n_pairs; sl += step_sl) {
        const int64_t pr = pair0 + sl;
        const int ra = pa[pr * stride], rb = pb[pr * stride];
        const int open = popen[pr * stride];
        const int nopen = -open;
        const uint8_t *s1 = Q.ptr(ra);
        const uint8_t *s2 = Q.ptr(rb);
        const int n1 = Q.len(ra);
        const int n2 = Q.len(rb);
        uint32_t *tr = trace + (size_t)sl * slot_words;
        if (MULTI) {
            for (int j = threadIdx.x; j < n2; j += blockDim.x) s2s[j] = k4_col_base(s2[j]);
            if (threadIdx.x < K4T_MAXW) prog[threadIdx.x] = 0;
            __syncthreads();
        } else {
            for (int j = lane; j < n2; j += 32) s2s[j] = k4_col_base(s2[j]);
            __syncwarp();
        }

        int bestv = NGSID_NEG_INF, besti = 0x7fffffff;
        int lrv = NGSID_NEG_INF, lrj = 0;
        const int rows_per_pass = 32 * K4T_RPL;
        const int npass = (n1 + rows_per_pass - 1) / rows_per_pass;
        const int nsteps = n2 + 31;
        for (int pass = MULTI ? wid : 0; pass < npass; pass += MULTI ? W : 1) {
            const int row0 = pass * rows_per_pass + lane * K4T_RPL;
            int2 *bnd_out = MULTI ? bnd_base + (size_t)(pass % W) * n2cap : bnd_base;
            const int2 *bnd_in = MULTI ? bnd_base + (size_t)((pass + W - 1) % W) * n2cap : bnd_base;
            volatile int *prog_in = prog + ((pass + W - 1) % W);
            volatile int *prog_out = prog + (pass % W);
            const int prog_base = pass * (n2 + 1);          // progress values of pass p: p (n2+1) + columns done
            int H[K4T_RPL], D[K4T_RPL];
            uint32_t c1[K4T_RPL];
#pragma unroll
            for (int r = 0; r < K4T_RPL; ++r) {
                int i = row0 + r;
                c1[r] = (i < n1) ? k4_row_base(s1[i]) : 0xffu;
                H[r] = 0;
                D[r] = NGSID_NEG_INF;
            }
            const int r_last = (n1 - 1) - row0;
            int oH = 0, oI = NGSID_NEG_INF;
            int dHp = 0;
            uint32_t *trp = tr + (size_t)pass * nsteps * 32 + lane;
            for (int t = 0; t < nsteps; ++t) {
                const int j = t - lane;
                if (MULTI && pass > 0 && t < n2) {
                    // column t of the pass above must be published before lane 0 reads it
                    const int need = prog_base - (n2 + 1) + t + 1;
                    while (*prog_in < need) { }
                    __threadfence_block();
                }
                int uH = __shfl_up_sync(NGSID_FULL_MASK, oH, 1);
                int uI = __shfl_up_sync(NGSID_FULL_MASK, oI, 1);
                if (j >= 0 && j < n2) {
                    if (lane == 0) {
                        if (pass == 0) { uH = 0; uI = NGSID_NEG_INF; }
                        else { int2 v = bnd_in[j]; uH = v.x; uI = v.y; }
                    }
                    const int sH = uH;
                    const uint32_t c2 = s2s[j];
                    int dH = dHp;
                    uint32_t word = 0;
                    K4T_CELL(0) K4T_CELL(1) K4T_CELL(2) K4T_CELL(3)
                    K4T_CELL(4) K4T_CELL(5) K4T_CELL(6) K4T_CELL(7)
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
                    if (lane == 31 && pass + 1 < npass) {
                        bnd_out[j] = make_int2(oH, oI);
                        if (MULTI) {
                            __threadfence_block();
                            *prog_out = prog_base + j + 1;
                        }
                    }
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
        if (MULTI) {
            // combine the warps: the last-column maximum over all strips (smallest row on ties);
            // the last row lives in exactly one strip
            if (lane == 0) {
                endbuf[4 * wid + 0] = bestv; endbuf[4 * wid + 1] = besti;
                endbuf[4 * wid + 2] = lrv;   endbuf[4 * wid + 3] = lrj;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < W; ++w) {
                    const int ov = endbuf[4 * w], oi = endbuf[4 * w + 1];
                    if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
                    if (endbuf[4 * w + 2] > lrv) { lrv = endbuf[4 * w + 2]; lrj = endbuf[4 * w + 3]; }
                }
            }
        }
        if (MULTI ? threadIdx.x == 0 : lane == 0) {
            K4TEnd e;
            e.pad = 0;
            if (lrv > bestv) { e.score = lrv; e.end_i = n1 - 1; e.end_j = lrj; }
            else { e.score = bestv; e.end_i = besti; e.end_j = n2 - 1; }
            ends[sl] = e;
        }
        if (MULTI) __syncthreads(); else __syncwarp();
    }
}

// ---- traceback: one warp per pair ---------------------------------------------------------------
struct K4TWindow { int32_t q_first, q_last, t_first, t_last; };   // q_last/t_last exclusive; -1 = none

struct K4TStat {
    uint32_t hist; int ncol; int cnt; uint32_t hm; int k, m;
    __device__ __forceinline__ void push(uint32_t bit) {
        hist = ((hist << 1) | bit) & hm;
        if (ncol < k) ncol++;
        if (ncol == k && __popc(hist) >= m) cnt++;
    }
};

// The path visits the 128-byte trace rows (one per wavefront step) of a strip in strictly
// descending order -- a diagonal or horizontal move goes to row t-1 of the same lane strip, a
// vertical move stays in the row or goes to row t-1 of the strip above. The warp therefore streams
// the rows through a shared-memory ring with cp.async, two batches of K4T_BATCH rows in flight, and
// every lane follows the same (uniform) state machine, reading the word it needs as a broadcast.
#define K4T_BATCH 16
#define K4T_RING (2 * K4T_BATCH)

// per warp: ring | both sequences
__host__ __device__ inline size_t k4t_tb_smem_per_warp(int ncap) { return (size_t)K4T_RING * 128 + 2 * (size_t)ncap; }

__device__ __forceinline__ void k4t_issue_batch(const uint32_t *tp, uint32_t ring_s, int t_hi, int lane)
{
    // rows t_hi, t_hi-1, ..., t_hi-K4T_BATCH+1 (those >= 0) into ring slot (t & (K4T_RING-1))
#pragma unroll
    for (int q = 0; q < K4T_BATCH; ++q) {
        const int t = t_hi - q;
        if (t >= 0) {
            const uint32_t dst = ring_s + (uint32_t)(((t & (K4T_RING - 1)) * 32 + lane) * 4);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(tp + (size_t)t * 32 + lane) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(128)
k4t_traceback_kernel(K4TSeqs Q, const int32_t *__restrict__ pa,
                     const int32_t *__restrict__ pb, const int32_t *__restrict__ pm, int stride,
                     int64_t pair0, int64_t n_pairs, int k, const uint32_t *__restrict__ trace,
                     size_t slot_words, const K4TEnd *__restrict__ ends,
                     int32_t *__restrict__ out_count, int32_t *__restrict__ out_score,
                     int32_t *__restrict__ out_match, int32_t *__restrict__ out_cols,
                     K4TWindow *__restrict__ out_win, int window, int ncap)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = (int)lane_id();
    const int wid = threadIdx.x >> 5;
    const int64_t sl = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    if (sl >= n_pairs) return;
    uint8_t *base = smem_raw + (size_t)wid * k4t_tb_smem_per_warp(ncap);
    const uint32_t *ring = reinterpret_cast<const uint32_t *>(base);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(base);
    uint8_t *q1 = base + K4T_RING * 128, *q2 = q1 + ncap;
    const int64_t pr = pair0 + sl;
    const int ra = pa[pr * stride], rb = pb[pr * stride];
    const int n1 = Q.len(ra);
    const int n2 = Q.len(rb);
    {
        const uint8_t *s1 = Q.ptr(ra), *s2 = Q.ptr(rb);
        for (int x = lane; x < n1; x += 32) q1[x] = s1[x];
        for (int x = lane; x < n2; x += 32) q2[x] = s2[x];
    }
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
    __syncwarp();
    while (i >= 0 && j >= 0) {
        // rows of the current strip of 256 rows, from the current one downwards
        const int pass = i >> 8;
        const uint32_t *tp = tr + (size_t)pass * nsteps * 32;
        const int t_start = j + ((i & 255) >> 3);
        __syncwarp();                                    // every lane is done with the previous strip's rows
        k4t_issue_batch(tp, ring_s, t_start, lane);
        k4t_issue_batch(tp, ring_s, t_start - K4T_BATCH, lane);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        int cur = 0;                                     // batch the walk is in
        while (i >= 0 && j >= 0 && (i >> 8) == pass) {
            const int strip = (i & 255) >> 3;
            const int t = j + strip;
            const int b = (t_start - t) >> 4;            // K4T_BATCH == 16
            if (b != cur) {
                // batch `cur` is consumed: its slots take batch cur + 2; batch cur + 1 must have landed
                __syncwarp();
                k4t_issue_batch(tp, ring_s, t_start - (cur + 2) * K4T_BATCH, lane);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                cur = b;
            }
            const uint32_t word = ring[(t & (K4T_RING - 1)) * 32 + strip];
            const uint32_t nib = (word >> (4 * (i & 7))) & 15u;
            if (state == 0) {
                if ((nib & 3u) == 3u) state = 3;
                else if (nib & 1u) state = 2;
                else {
                    const uint32_t mt = (q1[i] == q2[j]) ? 1u : 0u;
                    S.push(mt);
                    n_match += (int)mt;
                    n_cols++;
                    if (nwin) {
                        const int w = j / window;
                        if (lane == w) {
                            if (mywin.q_last < 0) { mywin.q_last = i + 1; mywin.t_last = j + 1; }
                            mywin.q_first = i; mywin.t_first = j;
                        }
                    }
                    --i; --j;
                }
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
        asm volatile("cp.async.wait_group 0;" ::: "memory");
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

// ---- traceback: one THREAD per pair (bulk launches) ---------------------------------------------
// The warp kernel above spends a whole warp on one sequential walk (~65 warp-instructions per step,
// 15 % of K4's instructions); with tens of thousands of pairs per launch the walks are better run
// 32 to a warp: every step is one dependent 4-byte load of the trace word (L2 / HBM latency, hidden
// by the number of pairs) plus the same state machine. Same results as the warp kernel (tested on
// the same pairs); no window breaking points (consensus uses the warp kernel).
__global__ void __launch_bounds__(128)
k4t_traceback_thread_kernel(K4TSeqs Q, const int32_t *__restrict__ pa,
                            const int32_t *__restrict__ pb, const int32_t *__restrict__ pm, int stride,
                            int64_t pair0, int64_t n_pairs, int k, const uint32_t *__restrict__ trace,
                            size_t slot_words, const K4TEnd *__restrict__ ends,
                            int32_t *__restrict__ out_count, int32_t *__restrict__ out_score,
                            int32_t *__restrict__ out_match, int32_t *__restrict__ out_cols)
{
    const int64_t sl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= n_pairs) return;
    const int64_t pr = pair0 + sl;
    const int ra = pa[pr * stride], rb = pb[pr * stride];
    const int n1 = Q.len(ra), n2 = Q.len(rb);
    const uint8_t *s1 = Q.ptr(ra), *s2 = Q.ptr(rb);
    const K4TEnd e = ends[sl];
    const uint32_t *tr = trace + (size_t)sl * slot_words;
    const size_t nsteps = (size_t)n2 + 31;
    K4TStat S;
    S.hist = 0; S.ncol = 0; S.cnt = 0; S.k = k; S.m = pm ? pm[pr * stride] : 1;
    S.hm = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    int n_match = 0, n_cols = 0;
    const int trailing = (n1 - 1 - e.end_i) + (n2 - 1 - e.end_j);
    for (int t = 0; t < trailing; ++t) S.push(0u);
    n_cols += trailing;
    int i = e.end_i, j = e.end_j, state = 0;
    size_t cur_idx = (size_t)-1;
    uint32_t word = 0;
    while (i >= 0 && j >= 0) {
        const int strip = (i & 255) >> 3;
        const size_t idx = ((size_t)(i >> 8) * nsteps + (size_t)(j + strip)) * 32 + (size_t)strip;
        if (idx != cur_idx) { word = __ldg(tr + idx); cur_idx = idx; }
        const uint32_t nib = (word >> (4 * (i & 7))) & 15u;
        if (state == 0) {
            if ((nib & 3u) == 3u) state = 3;
            else if (nib & 1u) state = 2;
            else {
                const uint32_t mt = (__ldg(s1 + i) == __ldg(s2 + j)) ? 1u : 0u;
                S.push(mt);
                n_match += (int)mt;
                n_cols++;
                --i; --j;
            }
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
    const int leading = (i + 1) + (j + 1);
    for (int t = 0; t < leading; ++t) S.push(0u);
    n_cols += leading;
    int cnt = S.cnt;
    if (n_cols < k) cnt = (__popc(S.hist) >= S.m) ? 1 : 0;
    if (out_count) out_count[pr] = cnt;
    if (out_score) out_score[pr] = e.score;
    if (out_match) out_match[pr] = n_match;
    if (out_cols) out_cols[pr] = n_cols;
}

