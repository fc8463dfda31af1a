// K5: partial-order-alignment DP + traceback of ONE layer of many jobs per launch.
// Replaces the DP inside spoa (consensus.run_spoa, modules/consensus.py:83-92: local, 5 / -4 / gap -2)
// and inside racon's window consensus (consensus.run_racon, modules/consensus.py:107-126: global,
// 3 / -5 / -4); graph semantics (node / edge lists, topological order, consensus) in poa_core.cuh.
//
// Division of labour: the graph of a job lives on the HOST -- adding an alignment and
// spoa's depth-first topological re-sort are O(V + E) pointer chasing, 20-40 us on a CPU core and
// 1-2 ms on a single GPU thread -- while every DP cell and the traceback run here. Per layer the host
// sends the graph rows in topological order (letter + predecessor rows) and gets the alignment path.
//
// Shape: one CTA per job. Warp w owns the columns [128 w, 128 w + 128) of the DP matrix, lane l four
// consecutive ones, and walks the graph rows top to bottom:
//   * a row needs its predecessor rows at columns j-1 and j: the previous row stays in registers,
//     the last K5R ring rows in a per-warp shared-memory ring, older ones come from the matrix in
//     global memory (only rows that some far successor reads are written there);
//   * the horizontal move H[i][j] = max(M[i][j], H[i][j-1] + g) (linear gap) is a max-plus prefix
//     scan: sequential over a lane's four columns, 5 shuffle steps across the warp, and a carry
//     from the warp to the left, which is therefore always one row ahead: the warps of a CTA form a
//     pipeline over the rows (progress counters + a 64-row ring of boundary values in shared
//     memory, no CTA-wide barrier per row);
//   * every cell stores one direction byte with the tie order of poa_traceback (diagonal in-edges
//     in list order, then vertical ones, then horizontal), so the traceback never re-derives a move.
// Traceback: warp 0 fetches 32 x 32 tiles of direction bytes (+ the rows' predecessor lists) into
// shared memory and lane 0 walks inside the tile; one global round trip per ~20 path steps.
#pragma once
#include "ngsid_internal.cuh"
#include "poa_core.cuh"

#define K5R_CPL 4                    // columns per lane
#define K5R_TILE (32 * K5R_CPL)      // columns per warp
#define K5R_MAXW 32                  // warps per CTA: layers up to 4095 bases
#define K5R_EDGE 64                  // boundary values kept per warp
#define K5R_DIAG 0                   // direction byte: 0..119 diagonal through in-edge u
#define K5R_UP 120                   // 120..239 vertical through in-edge u - 120
#define K5R_LEFT 254
#define K5R_STOP 255
#define K5R_MAXE 119

struct K5RJob {
    int32_t V, L;                    // graph rows, layer length
    int32_t mode;                    // 0 local, 1 global
    int32_t match, mismatch, gap;
    int64_t seq_off;                 // layer bases in the layer arena
    int64_t meta_off;                // first row record (uint4 units)
    int64_t ovf_off;                 // predecessor lists of rows with more than 3 in-edges (int32 units)
    int64_t mat_off;                 // first cell of this job's matrices (cells: (V + 1) x ld)
    int64_t path_off;                // first path entry (int2 units), capacity V + L + 2
};

struct K5RArgs {
    const K5RJob *jobs;
    const uint8_t *layers;           // bases of the layers of this step
    const uint4 *meta;               // per row: {letter | np << 8 | flags << 16, p0, p1, p2}  (np <= 3)
                                     //          {.., ovf index, -, -}                          (np > 3)
    const int32_t *ovf;
    int32_t *H;                      // matrices (only rows flagged 0x2 are written)
    uint8_t *DIR;
    int2 *path;                      // (matrix row or -1, layer position or -1), reverse order
    int32_t *out;                    // per job: n_path, best score, err, -
    int ld;                          // row stride of H and DIR (multiple of 128)
    int ring;                        // rows per shared-memory ring (power of two, <= 16)
};

#define K5R_FLAG_SINK 0x10000u
#define K5R_FLAG_STORE 0x20000u

__device__ __forceinline__ int k5r_ldvol(const volatile int *p) { return *p; }

template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k5r_layer_kernel(K5RArgs A)
{
    extern __shared__ __align__(16) int k5r_smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, NW = blockDim.x >> 5;
    const K5RJob J = A.jobs[blockIdx.x];
    const int V = J.V, L = J.L, g = J.gap, ld = A.ld, R = A.ring, RM = R - 1;
    // shared: progress[NW] | edge[NW][64] | ring[NW][R][32] int4 | traceback tiles
    volatile int *progress = k5r_smem;
    volatile int *edge = k5r_smem + K5R_MAXW;
    int4 *ringb = reinterpret_cast<int4 *>(k5r_smem + K5R_MAXW + K5R_MAXW * K5R_EDGE);
    int4 *myring = ringb + (size_t)w * R * 32;
    __shared__ int s_best[K5R_MAXW][3];
    if (tid < K5R_MAXW) progress[tid] = 0;
    __syncthreads();

    const int c0 = w * K5R_TILE + lane * K5R_CPL;            // first column of this lane
    const bool live_warp = w * K5R_TILE <= L;
    const uint8_t *s = A.layers + J.seq_off;
    int32_t *Hg = A.H + J.mat_off;
    uint8_t *Dg = A.DIR + J.mat_off;
    const uint4 *rows = A.meta + J.meta_off;
    const int32_t *ovf = A.ovf + J.ovf_off;

    int bestv = 0, besti = 0, bestj = 0;                     // local mode: best cell
    int sinkv = POA_NEG, sinki = 0x7fffffff;                 // global mode: best sink row at column L
    int err = 0;
    if (live_warp) {
        uint32_t sq[K5R_CPL];
        int gc[K5R_CPL];
#pragma unroll
        for (int t = 0; t < K5R_CPL; ++t) {
            const int c = c0 + t;
            sq[t] = (c >= 1 && c <= L) ? s[c - 1] : 0u;
            gc[t] = g * c;
        }
        auto virt = [&](int c) { return c < 0 ? POA_NEG : (MODE ? g * c : 0); };
        int prevH[K5R_CPL];
#pragma unroll
        for (int t = 0; t < K5R_CPL; ++t) prevH[t] = virt(c0 + t);
        int prevL = virt(c0 - 1);                            // lane 0: H[i-1][c0 - 1]
        int avail = 0;                                       // rows the warp to the left has finished
        uint4 mblk = make_uint4(0, 0, 0, 0);
        for (int i = 1; i <= V; ++i) {
            const int bi = (i - 1) & 31;
            if (bi == 0) mblk = (i - 1 + lane < V) ? __ldg(rows + (i - 1 + lane)) : make_uint4(0, 0, 0, 0);
            uint4 m;
            m.x = __shfl_sync(NGSID_FULL_MASK, mblk.x, bi); m.y = __shfl_sync(NGSID_FULL_MASK, mblk.y, bi);
            m.z = __shfl_sync(NGSID_FULL_MASK, mblk.z, bi); m.w = __shfl_sync(NGSID_FULL_MASK, mblk.w, bi);
            const int np = (int)((m.x >> 8) & 255u);
            const uint32_t letter = m.x & 255u;
            // ---- the warp to the left has to be past this row; its boundary value is the carry
            int edgeIn = POA_NEG;
            if (w > 0) {
                if (avail < i) {
                    if (lane == 0) { while ((avail = k5r_ldvol(progress + (w - 1))) < i) { } }
                    avail = __shfl_sync(NGSID_FULL_MASK, avail, 0);
                    __threadfence_block();
                }
                if (lane == 0) edgeIn = edge[(w - 1) * K5R_EDGE + (i & (K5R_EDGE - 1))];
            }
            // ---- do not run more than 48 rows ahead of the warp to the right (it still reads the ring of
            // boundary values for its near predecessors)
            if (w + 1 < NW && (w + 1) * K5R_TILE <= L && i > 48) {
                if (lane == 0) { while (k5r_ldvol(progress + (w + 1)) < i - 48) { } }
                __syncwarp();
            }
            int bd[K5R_CPL], bu[K5R_CPL];
            uint32_t dd[K5R_CPL], du[K5R_CPL];
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) { bd[t] = POA_NEG; bu[t] = POA_NEG; dd[t] = K5R_STOP; du[t] = K5R_UP; }
            const int ne = np > 3 ? np : (np == 0 ? 1 : np);
            if (np > K5R_MAXE) err = 7;
            for (int u = 0; u < ne; ++u) {
                int p;
                if (np > 3) p = __ldg(ovf + m.y + u);
                else p = (np == 0) ? 0 : (u == 0 ? (int)m.y : (u == 1 ? (int)m.z : (int)m.w));
                int hv[K5R_CPL], left;
                if (p == i - 1) {
#pragma unroll
                    for (int t = 0; t < K5R_CPL; ++t) hv[t] = prevH[t];
                    left = __shfl_up_sync(NGSID_FULL_MASK, prevH[K5R_CPL - 1], 1);
                    if (lane == 0) left = prevL;
                } else if (p == 0) {
#pragma unroll
                    for (int t = 0; t < K5R_CPL; ++t) hv[t] = virt(c0 + t);
                    left = virt(c0 - 1);
                } else if (i - p < R) {
                    const int4 v4 = myring[(size_t)(p & RM) * 32 + lane];
                    hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                    left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                    if (lane == 0) left = (w > 0) ? edge[(w - 1) * K5R_EDGE + (p & (K5R_EDGE - 1))] : POA_NEG;
                } else {
                    const int4 v4 = __ldcg(reinterpret_cast<const int4 *>(Hg + (size_t)p * ld + c0));
                    hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                    left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                    if (lane == 0) left = (c0 > 0) ? __ldcg(Hg + (size_t)p * ld + c0 - 1) : POA_NEG;
                }
#pragma unroll
                for (int t = 0; t < K5R_CPL; ++t) {
                    const int dg = (t == 0) ? left : hv[t - 1];
                    if (dg > bd[t]) { bd[t] = dg; dd[t] = K5R_DIAG + u; }
                    if (hv[t] > bu[t]) { bu[t] = hv[t]; du[t] = K5R_UP + u; }
                }
            }
            // ---- cell values before the horizontal move, then the max-plus scan along the row
            int Mv[K5R_CPL], run[K5R_CPL];
            uint32_t dm[K5R_CPL];
            int acc = POA_NEG;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                const int sc = (letter == sq[t]) ? J.match : J.mismatch;
                int h = bd[t] + sc;
                uint32_t d = dd[t];
                if (bu[t] + g > h) { h = bu[t] + g; d = du[t]; }
                if (!MODE && h <= 0) { h = 0; d = K5R_STOP; }
                Mv[t] = h; dm[t] = d;
                acc = max(acc, h - gc[t]);
                run[t] = acc;
            }
            const int carryX = (lane == 0) ? ((w > 0) ? edgeIn - g * (c0 - 1) : POA_NEG) : POA_NEG;
            int incl = (lane == 0) ? max(acc, carryX) : acc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(NGSID_FULL_MASK, incl, d);
                if (lane >= d) incl = max(incl, o);
            }
            int excl = __shfl_up_sync(NGSID_FULL_MASK, incl, 1);
            if (lane == 0) excl = carryX;
            int Hv[K5R_CPL];
            uint32_t dirw = 0;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                const int h = max(run[t], excl) + gc[t];
                uint32_t d = dm[t];
                if (h > Mv[t]) d = K5R_LEFT;
                Hv[t] = h;
                dirw |= d << (8 * t);
                const int c = c0 + t;
                if (MODE) {
                    if (c == L && (m.x & K5R_FLAG_SINK) && h > sinkv) { sinkv = h; sinki = i; }
                } else if (c <= L && h > bestv) { bestv = h; besti = i; bestj = c; }
            }
            *reinterpret_cast<uint32_t *>(Dg + (size_t)i * ld + c0) = dirw;
            myring[(size_t)(i & RM) * 32 + lane] = make_int4(Hv[0], Hv[1], Hv[2], Hv[3]);
            if (m.x & K5R_FLAG_STORE) __stcg(reinterpret_cast<int4 *>(Hg + (size_t)i * ld + c0), make_int4(Hv[0], Hv[1], Hv[2], Hv[3]));
            if (lane == 31) edge[w * K5R_EDGE + (i & (K5R_EDGE - 1))] = Hv[K5R_CPL - 1];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) progress[w] = i;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) prevH[t] = Hv[t];
            prevL = edgeIn;
        }
    }
    // ---- end cell
    if (MODE) { bestv = sinkv; besti = sinki; bestj = L; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
        const int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
        const int oj = __shfl_xor_sync(NGSID_FULL_MASK, bestj, d);
        if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
        err |= __shfl_xor_sync(NGSID_FULL_MASK, err, d);
    }
    if (lane == 0) { s_best[w][0] = bestv; s_best[w][1] = besti; s_best[w][2] = bestj; if (err) A.out[blockIdx.x * 4 + 2] = err; }
    __threadfence();
    __syncthreads();
    if (w != 0) return;
    // ---- traceback by warp 0: 32 x 32 tiles of direction bytes + row records in shared memory
    int bv = s_best[0][0], bi_ = s_best[0][1], bj_ = s_best[0][2];
    for (int x = 1; x < NW; ++x) {
        const int ov = s_best[x][0], oi = s_best[x][1], oj = s_best[x][2];
        if (ov > bv || (ov == bv && (oi < bi_ || (oi == bi_ && oj < bj_)))) { bv = ov; bi_ = oi; bj_ = oj; }
    }
    uint8_t *tdir = reinterpret_cast<uint8_t *>(k5r_smem + K5R_MAXW + K5R_MAXW * K5R_EDGE);    // the rings are free now
    uint4 *tmeta = reinterpret_cast<uint4 *>(tdir + 32 * 32);
    int2 *path = A.path + J.path_off;
    int n = 0;
    int i = bi_, j = bj_;
    bool go = !(MODE == 0 && bv <= 0) && V > 0 && i >= 1 && i <= V;
    if (MODE && V > 0 && (i < 1 || i > V)) { if (lane == 0) A.out[blockIdx.x * 4 + 2] = 8; go = false; }    // no sink reached
    while (go) {
        if (i == 0) {
            if (MODE) { if (lane == 0) while (j > 0) { path[n++] = make_int2(-1, j - 1); --j; } }
            break;
        }
        const int top = i, j0 = j & ~31;
        {
            const int r = top - lane;
            uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0), mt = make_uint4(0, 0, 0, 0);
            if (r >= 1) {
                const uint4 *p = reinterpret_cast<const uint4 *>(Dg + (size_t)r * ld + j0);
                a = __ldcg(p); b = __ldcg(p + 1);
                mt = __ldg(rows + (r - 1));
            }
            uint4 *td = reinterpret_cast<uint4 *>(tdir + lane * 32);
            td[0] = a; td[1] = b;
            tmeta[lane] = mt;
        }
        __syncwarp();
        int stop = 0;
        if (lane == 0) {
            while (i >= 1 && top - i < 32 && j >= j0) {
                const int d = tdir[(top - i) * 32 + (j - j0)];
                if (d == K5R_STOP) { stop = 1; break; }
                if (d == K5R_LEFT) { path[n++] = make_int2(-1, j - 1); --j; continue; }
                const uint4 mt = tmeta[top - i];
                const int np = (int)((mt.x >> 8) & 255u);
                const int u = d >= K5R_UP ? d - K5R_UP : d;
                int pr;
                if (np == 0) pr = 0;
                else if (np > 3) pr = __ldg(ovf + mt.y + u);
                else pr = u == 0 ? (int)mt.y : (u == 1 ? (int)mt.z : (int)mt.w);
                if (d >= K5R_UP) { path[n++] = make_int2(i, -1); i = pr; }
                else { path[n++] = make_int2(i, j - 1); i = pr; --j; }
                if (!MODE && i == 0) { stop = 1; break; }
            }
            if (MODE && i == 0 && j == 0) stop = 1;
        }
        stop = __shfl_sync(NGSID_FULL_MASK, stop, 0);
        i = __shfl_sync(NGSID_FULL_MASK, i, 0);
        j = __shfl_sync(NGSID_FULL_MASK, j, 0);
        n = __shfl_sync(NGSID_FULL_MASK, n, 0);
        __syncwarp();
        if (stop) break;
    }
    n = __shfl_sync(NGSID_FULL_MASK, n, 0);
    if (lane == 0) { A.out[blockIdx.x * 4] = n; A.out[blockIdx.x * 4 + 1] = bv; }
}

static inline size_t k5r_smem_bytes(int n_warps, int ring)
{
    const size_t fixed = (size_t)(K5R_MAXW + K5R_MAXW * K5R_EDGE) * 4;
    const size_t rings = (size_t)n_warps * ring * 32 * 16;
    return fixed + std::max<size_t>(rings, 32 * 32 + 32 * 16);
}
