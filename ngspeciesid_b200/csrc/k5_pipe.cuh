// K5, row-pipelined shape: the same DP, direction bytes and traceback as k5_rows.cuh (the structures,
// direction codes and the traceback come from there), with the warps of a CTA laid out the other way
// round.
//
// k5_rows.cuh gives every warp a tile of 128 columns and lets it walk ALL graph rows; the warps form a
// pipeline over the column tiles, and a job takes  V x (time one warp needs for a row)  -- ~335
// instructions issued by one warp, 1 000-1 400 cycles. Here warp w takes the rows w+1, w+1+NW, ... and
// walks the column chunks of its row left to right; row i+1 (next warp) follows row i one chunk
// behind. A job now takes  V x (time of ONE CHUNK)  as long as NW chunks cover a row's own time: the
// per-row overhead (row record, predecessor list, bookkeeping) is paid by a warp that is off the
// critical path for NW - 1 rows out of NW.
//
// Hand-over between rows: the H values of the last RR rows live in a shared-memory ring
// ring[row mod RR][column]; a warp publishes (row << 8 | chunks done) after the ring stores of a chunk
// (block-level fence), and the warp of row i+1 polls that word of its left neighbour before it touches
// chunk c. Every row waits for the row before it whether or not that row is a predecessor, so the rows
// complete in order: when row i is at chunk c, every earlier row has finished chunk c, and every row
// <= i - NW has finished altogether. That makes one poll per chunk sufficient for all predecessors,
// makes a ring of RR >= R + NW - 1 rows safe (R = the host's near-predecessor distance; rows that a
// successor at distance >= R reads are also written to the global matrix, K5R_FLAG_STORE), and keeps
// far-row reads from global memory ordered behind their writes.
#pragma once
#include "k5_rows.cuh"

#define K5P_STATE 32                 // ints reserved for the per-warp progress words

static inline size_t k5p_smem_bytes(int ld, int rr)
{
    // progress words | ring rr x ld ints | layer bases shifted by one column (ld bytes) -- the rings double
    // as traceback tiles afterwards (32 x 32 bytes + 32 row records)
    return (size_t)K5P_STATE * 4 + (size_t)rr * ld * 4 + (size_t)ld + 16;
}

template <int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) k5p_layer_kernel(K5RArgs A)
{
    extern __shared__ __align__(16) int k5p_smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const K5RJob J = A.jobs[blockIdx.x];
    const int V = J.V, L = J.L, g = J.gap, ld = A.ld, R = A.ring, RRM = A.rr - 1;
    volatile int *state = k5p_smem;
    int *ring = k5p_smem + K5P_STATE;
    uint8_t *lay = reinterpret_cast<uint8_t *>(ring + (size_t)A.rr * ld);    // lay[c] = base of column c (c >= 1), lay[0] = 0
    __shared__ int s_best[32][3];
    const uint8_t *s = A.layers + J.seq_off;
    int32_t *Hg = A.H + J.mat_off;
    uint8_t *Dg = A.DIR + J.mat_off;
    const uint4 *rows = A.meta + J.meta_off;
    const int32_t *ovf = A.ovf + J.ovf_off;
    if (tid < K5P_STATE) state[tid] = 0;
    for (int c = tid; c < ld; c += blockDim.x) lay[c] = (c >= 1 && c <= L) ? s[c - 1] : (uint8_t)0;   // 0 never matches
    __syncthreads();

    const int NEG8 = -(1 << 30);
    const int m8 = J.match << 8, x8 = J.mismatch << 8, g8 = g << 8;
    const int nch = (L + 1 + K5R_TILE - 1) / K5R_TILE;
    int bestv = 0, besti = 0, bestj = 0;                     // local mode: best cell
    int sinkv = POA_NEG, sinki = 0x7fffffff;                 // global mode: best sink row at column L
    int err = 0, far_rows = 0;
    const long long t_start = clock64();
    const volatile int *lstate = state + ((w + NWARPS - 1) % NWARPS);

    uint4 m = (w + 1 <= V) ? __ldg(rows + w) : make_uint4(0, 0, 0, 0);
    for (int i = w + 1; i <= V; i += NWARPS) {
        const uint4 m_n = (i + NWARPS <= V) ? __ldg(rows + i + NWARPS - 1) : make_uint4(0, 0, 0, 0);
        const int np = (int)((m.x >> 8) & 255u);
        const uint32_t letter = m.x & 255u;
        const bool inl = k5r_is_inline(m, np);
        const int ne = np == 0 ? 1 : np;
        if (np > K5R_MAXE) err = 7;
        int carry8 = NEG8;                                   // H8[i][c0 - 1] of lane 0: last column of the previous chunk
        int *myrow = ring + (size_t)(i & RRM) * ld;
        uint8_t *dirp = Dg + (size_t)i * ld;
        for (int ch = 0; ch < nch; ++ch) {
            const int c0 = ch * K5R_TILE + lane * K5R_CPL;
            // ---- the row before this one has to be past this chunk (all earlier rows are, then)
            if (i > 1) {
                int st = *lstate;
                while (!((st >> 8) > i - 1 || ((st >> 8) == i - 1 && (st & 255) > ch))) st = *lstate;
                __threadfence_block();
            }
            const uint32_t sq4 = *reinterpret_cast<const uint32_t *>(lay + c0);
            int key[K5R_CPL], sc8[K5R_CPL], gc[K5R_CPL];
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                key[t] = (int)0x80000000;
                sc8[t] = ((letter == ((sq4 >> (8 * t)) & 255u)) ? m8 : x8) + 255;
                gc[t] = g * (c0 + t);
            }
            int plist = 0;                                   // long lists: 32 predecessors per coalesced load
            for (int u = 0; u < ne; ++u) {
                int p;
                if (!inl) {
                    if ((u & 31) == 0) plist = (u + lane < np) ? __ldg(ovf + m.y + u + lane) : 0;
                    p = __shfl_sync(NGSID_FULL_MASK, plist, u & 31);
                } else p = (np == 0) ? 0 : k5r_inline_pred(m, u);
                int hv[K5R_CPL], left;
                if (p == 0) {                                // virtual row 0
#pragma unroll
                    for (int t = 0; t < K5R_CPL; ++t) hv[t] = (MODE ? gc[t] : 0) << 8;
                    left = c0 > 0 ? ((MODE ? g * (c0 - 1) : 0) << 8) : NEG8;
                } else if (i - p < R) {
                    const int *prow = ring + (size_t)(p & RRM) * ld;
                    const int4 v4 = *reinterpret_cast<const int4 *>(prow + c0);
                    hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                    left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                    if (lane == 0) left = c0 > 0 ? prow[c0 - 1] : NEG8;
                } else {
                    ++far_rows;
                    const int4 v4 = __ldcg(reinterpret_cast<const int4 *>(Hg + (size_t)p * ld + c0));
                    hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                    left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                    if (lane == 0) left = c0 > 0 ? __ldcg(Hg + (size_t)p * ld + c0 - 1) : NEG8;
                }
                // keys (H << 8) | (255 - direction code): see k5_rows.cuh
                const int cu_ = g8 + 135 - u;
                key[0] = __vimax3_s32(key[0], left + (sc8[0] - u), hv[0] + cu_);
                key[1] = __vimax3_s32(key[1], hv[0] + (sc8[1] - u), hv[1] + cu_);
                key[2] = __vimax3_s32(key[2], hv[1] + (sc8[2] - u), hv[2] + cu_);
                key[3] = __vimax3_s32(key[3], hv[2] + (sc8[3] - u), hv[3] + cu_);
            }
            // ---- values before the horizontal move, then the max-plus scan along the chunk
            int hM[K5R_CPL], run[K5R_CPL];
            int acc = POA_NEG;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                if (!MODE && key[t] < 256) key[t] = 0;       // local: score <= 0 -> 0 and STOP (code 255)
                hM[t] = key[t] >> 8;
                acc = max(acc, hM[t] - gc[t]);
                run[t] = acc;
            }
            const int carryX = (lane == 0 && ch > 0) ? (carry8 >> 8) - g * (c0 - 1) : POA_NEG;
            int incl = max(acc, carryX);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(NGSID_FULL_MASK, incl, d);
                if (lane >= d) incl = max(incl, o);
            }
            int excl = __shfl_up_sync(NGSID_FULL_MASK, incl, 1);
            if (lane == 0) excl = carryX;
            int Hv[K5R_CPL];
            uint32_t codes = 0;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                const int h = max(run[t], excl) + gc[t];
                const uint32_t code = (h > hM[t]) ? 1u : ((uint32_t)key[t] & 255u);      // horizontal: code 254
                codes |= code << (8 * t);
                Hv[t] = h;
            }
            const int4 h8 = make_int4(Hv[0] << 8, Hv[1] << 8, Hv[2] << 8, Hv[3] << 8);
            *reinterpret_cast<int4 *>(myrow + c0) = h8;
            if (m.x & K5R_FLAG_STORE) __stcg(reinterpret_cast<int4 *>(Hg + (size_t)i * ld + c0), h8);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) state[w] = (i << 8) | (ch + 1);
            // ---- off the chain
            carry8 = __shfl_sync(NGSID_FULL_MASK, h8.w, 31);
            *reinterpret_cast<uint32_t *>(dirp + c0) = ~codes;
            const int rowmax = max(max(Hv[0], Hv[1]), max(Hv[2], Hv[3]));
            if (MODE) {
                const int tl = L - c0;
                if (tl >= 0 && tl < K5R_CPL && (m.x & K5R_FLAG_SINK)) {
                    const int h = tl == 0 ? Hv[0] : (tl == 1 ? Hv[1] : (tl == 2 ? Hv[2] : Hv[3]));
                    if (h > sinkv || (h == sinkv && i < sinki)) { sinkv = h; sinki = i; }
                }
            } else if (rowmax > bestv || (rowmax == bestv && rowmax > 0 && i < besti)) {
                // rows reach a lane in increasing order per warp only: keep the smallest row on ties
                // (columns of a lane increase with the chunk, so the first hit of a row is its smallest column)
                const int jj = c0 + (Hv[0] == rowmax ? 0 : (Hv[1] == rowmax ? 1 : (Hv[2] == rowmax ? 2 : 3)));
                if (rowmax > bestv || i < besti || (i == besti && jj < bestj)) { bestv = rowmax; besti = i; bestj = jj; }
            }
        }
        m = m_n;
    }
    // ---- end cell: best over lanes, then over warps (score, then smallest row, then smallest column)
    if (MODE) { bestv = sinkv; besti = sinki; bestj = L; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
        const int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
        const int oj = __shfl_xor_sync(NGSID_FULL_MASK, bestj, d);
        if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
        err |= __shfl_xor_sync(NGSID_FULL_MASK, err, d);
        far_rows += __shfl_xor_sync(NGSID_FULL_MASK, far_rows, d);
    }
    if (lane == 0) { s_best[w][0] = bestv; s_best[w][1] = besti; s_best[w][2] = bestj; if (err) A.out[blockIdx.x * 16 + 2] = err; }
    __threadfence();
    __syncthreads();
    if (w != 0) return;
    const long long t_tb = clock64();
    if (lane == 0) { A.out[blockIdx.x * 16 + 3] = (int)((t_tb - t_start) >> 10); A.out[blockIdx.x * 16 + 5] = far_rows >> 5; }
    int bv = s_best[0][0], bi_ = s_best[0][1], bj_ = s_best[0][2];
    for (int x = 1; x < NWARPS; ++x) {
        const int ov = s_best[x][0], oi = s_best[x][1], oj = s_best[x][2];
        if (ov > bv || (ov == bv && (oi < bi_ || (oi == bi_ && oj < bj_)))) { bv = ov; bi_ = oi; bj_ = oj; }
    }
    k5r_traceback<MODE>(A, J, reinterpret_cast<uint8_t *>(ring), bv, bi_, bj_, t_tb);
}
