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
                                     //          {.. | PACK16, six 16-bit rows}                  (np <= 6, V < 65535)
                                     //          {.., ovf index, -, -}                          (longer lists)
    const int32_t *ovf;
    const uint2 *plan;               // per row: x = chain u | virtual u << 8 | prefetched << 16 | generic << 24 (u = 255: none)
                                     //          y = distance A | u A << 8 | distance B << 16 | u B << 24
    int32_t *H;                      // matrices (only rows flagged 0x2 are written)
    uint8_t *DIR;
    int2 *path;                      // (matrix row or -1, layer position or -1), reverse order
    int32_t *out;                    // per job (8 ints): n_path, best score, err, DP kilocycles, traceback kilocycles, rows with a far predecessor, kilocycles warp 1 waited
    int ld;                          // row stride of H and DIR (multiple of 128)
    int ring;                        // rows per shared-memory ring (power of two, <= 16)
};

#define K5R_FLAG_SINK 0x10000u
#define K5R_FLAG_STORE 0x20000u
#define K5R_FLAG_PACK16 0x40000u     // up to 6 predecessor rows as 16-bit fields of .y .z .w (graphs below 65 536 rows)
#define K5R_INLINE32 3
#define K5R_INLINE16 6

// predecessor u of a row record whose list is inline
__device__ __forceinline__ int k5r_inline_pred(const uint4 &m, int u)
{
    if (m.x & K5R_FLAG_PACK16) {
        const uint32_t wsel = u < 2 ? m.y : (u < 4 ? m.z : m.w);
        return (int)((wsel >> ((u & 1) * 16)) & 0xffffu);
    }
    return u == 0 ? (int)m.y : (u == 1 ? (int)m.z : (int)m.w);
}
__device__ __forceinline__ bool k5r_is_inline(const uint4 &m, int np)
{
    return np <= ((m.x & K5R_FLAG_PACK16) ? K5R_INLINE16 : K5R_INLINE32);
}

__device__ __forceinline__ int k5r_ldvol(const volatile int *p) { return *p; }

template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k5r_layer_kernel(K5RArgs A)
{
    extern __shared__ __align__(16) int k5r_smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, NW = blockDim.x >> 5;
    const K5RJob J = A.jobs[blockIdx.x];
    const int V = J.V, L = J.L, g = J.gap, ld = A.ld, R = A.ring, RM = R - 1;
    // shared: progress[32] | edge[32][64] {value, row tag} | ring[NW][R][32] int4 | row records | traceback tiles
    volatile int *progress = k5r_smem;
    volatile unsigned long long *edge = reinterpret_cast<volatile unsigned long long *>(k5r_smem + K5R_MAXW);
    int4 *ringb = reinterpret_cast<int4 *>(k5r_smem + K5R_MAXW + 2 * K5R_MAXW * K5R_EDGE);
    int4 *myring = ringb + (size_t)w * R * 32;
    __shared__ int s_best[K5R_MAXW][3];
    if (tid < K5R_MAXW) progress[tid] = 0;
    for (int x = tid; x < K5R_MAXW * K5R_EDGE; x += blockDim.x) edge[x] = 0ull;      // tag 0 = no row
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
    const long long t_start = clock64();
    long long t_wait = 0, t_bp = 0, t_pred = 0, t_scan = 0, t_store = 0, t_mark = 0;
    int far_rows = 0;
    if (live_warp) {
        // Cell values travel as keys  (H << 8) | (255 - direction code): one signed maximum picks the
        // larger score and, on equal scores, the smaller code -- diagonal in-edges in list order
        // (codes 0..119), then vertical ones (120..239), exactly poa_traceback's tie order -- so a
        // predecessor costs two adds and one 3-input maximum per column. H << 8 is what the ring, the
        // boundary values and the global matrix hold.
        //
        // The rows form one dependency chain (row i needs row i-1), so the loop is organised around its
        // latency: everything a row needs EXCEPT the previous row -- its record, its plan, the ring
        // rows of its other near predecessors, their boundary values -- is fetched one row ahead
        // (those rows are final by then); the chain itself is shuffle(prev) -> max3 -> scan -> H.
        // The host classifies every row (plan word): rows with at most the previous row, the virtual
        // row and two more near predecessors take this path, the rest the generic loop.
        const int NEG8 = -(1 << 30);
        uint32_t sq[K5R_CPL];
        int gc[K5R_CPL], vr[K5R_CPL];
#pragma unroll
        for (int t = 0; t < K5R_CPL; ++t) {
            const int c = c0 + t;
            sq[t] = (c >= 1 && c <= L) ? s[c - 1] : 0u;      // 0 never matches: cells right of column L only lose score
            gc[t] = g * c;
            vr[t] = (MODE ? g * c : 0) << 8;                 // virtual row 0
        }
        const int vleft = c0 > 0 ? ((MODE ? g * (c0 - 1) : 0) << 8) : NEG8;
        const int m8 = J.match << 8, x8 = J.mismatch << 8, g8 = g << 8;
        int prevH[K5R_CPL];
#pragma unroll
        for (int t = 0; t < K5R_CPL; ++t) prevH[t] = vr[t];
        int prevL = vleft;                                   // lane 0: H8[i-1][c0 - 1]
        uint4 *mrow = reinterpret_cast<uint4 *>(ringb) + (size_t)NW * R * 32 + (size_t)w * 32;
        uint2 *mplan = reinterpret_cast<uint2 *>(reinterpret_cast<uint4 *>(ringb) + (size_t)NW * R * 32 + (size_t)NW * 32) + (size_t)w * 32;
        const uint2 *plans = A.plan + J.meta_off;
        const volatile unsigned long long *ledge = edge + (w > 0 ? w - 1 : 0) * K5R_EDGE;
        uint8_t *dirp = Dg + (size_t)ld + c0;               // row 1
        int32_t *hgp = Hg + (size_t)ld + c0;
        // block of 32 row records + plans: rows 1..32
        mrow[lane] = (lane < V) ? __ldg(rows + lane) : make_uint4(0, 0, 0, 0);
        mplan[lane] = (lane < V) ? __ldg(plans + lane) : make_uint2(0, 0);
        __syncwarp();
        uint4 m = mrow[0];
        uint2 pl = mplan[0];
        int4 pa = make_int4(0, 0, 0, 0), pb = make_int4(0, 0, 0, 0);   // prefetched near predecessors of the current row
        int la = NEG8, lb = NEG8;
        for (int i = 1; i <= V; ++i) {
            // ---- fetch ahead for row i + 1 (record, plan, ring rows of its near predecessors other than row i)
            uint4 m_n = make_uint4(0, 0, 0, 0);
            uint2 pl_n = make_uint2(0, 0);
            int4 pa_n = make_int4(0, 0, 0, 0), pb_n = make_int4(0, 0, 0, 0);
            int la_n = NEG8, lb_n = NEG8;
            if (i < V) {
                if ((i & 31) == 0) {                         // next block of 32 rows (the current row is in registers)
                    __syncwarp();
                    mrow[lane] = (i + lane < V) ? __ldg(rows + i + lane) : make_uint4(0, 0, 0, 0);
                    mplan[lane] = (i + lane < V) ? __ldg(plans + i + lane) : make_uint2(0, 0);
                    __syncwarp();
                }
                m_n = mrow[i & 31];
                pl_n = mplan[i & 31];
                const int npre = (int)((pl_n.x >> 16) & 255u);
                if (!(pl_n.x >> 24) && npre > 0) {
                    const int ra = i + 1 - (int)(pl_n.y & 255u);
                    pa_n = myring[(size_t)(ra & RM) * 32 + lane];
                    const int ea = (w > 0) ? (int)(uint32_t)ledge[ra & (K5R_EDGE - 1)] : NEG8;
                    la_n = __shfl_up_sync(NGSID_FULL_MASK, pa_n.w, 1);
                    if (lane == 0) la_n = ea;
                    if (npre > 1) {
                        const int rb = i + 1 - (int)((pl_n.y >> 16) & 255u);
                        pb_n = myring[(size_t)(rb & RM) * 32 + lane];
                        const int eb = (w > 0) ? (int)(uint32_t)ledge[rb & (K5R_EDGE - 1)] : NEG8;
                        lb_n = __shfl_up_sync(NGSID_FULL_MASK, pb_n.w, 1);
                        if (lane == 0) lb_n = eb;
                    }
                }
            }
            const int np = (int)((m.x >> 8) & 255u);
            const uint32_t letter = m.x & 255u;
            // ---- the warp to the left has to be past this row; its boundary value is the carry
            // (an entry of the boundary ring is one 64-bit word {value, row}: no fence between value and
            // flag. Every lane polls the same word, so the branch is warp-uniform: a one-lane spin loop
            // leaves the warp diverged and every later shuffle takes the slow divergent path.)
            int edgeIn = NEG8;
            if (w > 0) {
                const volatile unsigned long long *slot = ledge + (i & (K5R_EDGE - 1));
                unsigned long long e = *slot;
                while ((int)(e >> 32) != i) e = *slot;
                edgeIn = (int)(uint32_t)e;
            }
            // ---- do not run more than 40 rows ahead of the warp to the right (it still reads the ring of
            // boundary values for its near predecessors)
            if (w + 1 < NW && (w + 1) * K5R_TILE <= L && i > 40 && (i & 7) == 0) {
                while (k5r_ldvol(progress + (w + 1)) < i - 32) { }
            }
            int key[K5R_CPL], sc8[K5R_CPL];
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) { key[t] = (int)0x80000000; sc8[t] = ((letter == sq[t]) ? m8 : x8) + 255; }
#define K5R_TAKE(HV0, HV1, HV2, HV3, LEFT, U)                                                            \
            {                                                                                             \
                const int cu_ = g8 + 135 - (U);                                                           \
                key[0] = __vimax3_s32(key[0], (LEFT) + (sc8[0] - (U)), (HV0) + cu_);                      \
                key[1] = __vimax3_s32(key[1], (HV0) + (sc8[1] - (U)), (HV1) + cu_);                       \
                key[2] = __vimax3_s32(key[2], (HV1) + (sc8[2] - (U)), (HV2) + cu_);                       \
                key[3] = __vimax3_s32(key[3], (HV2) + (sc8[3] - (U)), (HV3) + cu_);                       \
            }
            if (!(pl.x >> 24)) {
                const int cu = (int)(pl.x & 255u), vu = (int)((pl.x >> 8) & 255u), npre = (int)((pl.x >> 16) & 255u);
                if (cu != 255) {
                    int left = __shfl_up_sync(NGSID_FULL_MASK, prevH[K5R_CPL - 1], 1);
                    if (lane == 0) left = prevL;
                    K5R_TAKE(prevH[0], prevH[1], prevH[2], prevH[3], left, cu)
                }
                if (vu != 255) K5R_TAKE(vr[0], vr[1], vr[2], vr[3], vleft, vu)
                if (npre > 0) K5R_TAKE(pa.x, pa.y, pa.z, pa.w, la, (int)((pl.y >> 8) & 255u))
                if (npre > 1) K5R_TAKE(pb.x, pb.y, pb.z, pb.w, lb, (int)((pl.y >> 24) & 255u))
            } else {
                // generic: any number of predecessors, near or far
                const bool inl = k5r_is_inline(m, np);
                const int ne = np == 0 ? 1 : np;
                if (np > K5R_MAXE) err = 7;
                int plist = 0;                               // long lists: 32 predecessors per coalesced load
                for (int u = 0; u < ne; ++u) {
                    int p;
                    if (!inl) {
                        if ((u & 31) == 0) plist = (u + (int)lane < np) ? __ldg(ovf + m.y + u + lane) : 0;
                        p = __shfl_sync(NGSID_FULL_MASK, plist, u & 31);
                    } else p = (np == 0) ? 0 : k5r_inline_pred(m, u);
                    int hv[K5R_CPL], left;
                    if (p == i - 1) {
#pragma unroll
                        for (int t = 0; t < K5R_CPL; ++t) hv[t] = prevH[t];
                        left = __shfl_up_sync(NGSID_FULL_MASK, prevH[K5R_CPL - 1], 1);
                        if (lane == 0) left = prevL;
                    } else if (p == 0) {
#pragma unroll
                        for (int t = 0; t < K5R_CPL; ++t) hv[t] = vr[t];
                        left = vleft;
                    } else if (i - p < R) {
                        const int4 v4 = myring[(size_t)(p & RM) * 32 + lane];
                        hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                        left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                        const int el = (w > 0) ? (int)(uint32_t)ledge[p & (K5R_EDGE - 1)] : NEG8;
                        if (lane == 0) left = el;
                    } else {
                        ++far_rows;
                        const int4 v4 = __ldcg(reinterpret_cast<const int4 *>(Hg + (size_t)p * ld + c0));
                        hv[0] = v4.x; hv[1] = v4.y; hv[2] = v4.z; hv[3] = v4.w;
                        left = __shfl_up_sync(NGSID_FULL_MASK, v4.w, 1);
                        const int el = (w > 0) ? __ldcg(Hg + (size_t)p * ld + w * K5R_TILE - 1) : NEG8;
                        if (lane == 0) left = el;
                    }
                    K5R_TAKE(hv[0], hv[1], hv[2], hv[3], left, u)
                }
            }
#undef K5R_TAKE
            // ---- values before the horizontal move, then the max-plus scan along the row
            int hM[K5R_CPL], run[K5R_CPL];
            int acc = POA_NEG;
#pragma unroll
            for (int t = 0; t < K5R_CPL; ++t) {
                if (!MODE && key[t] < 256) key[t] = 0;       // local: score <= 0 -> 0 and STOP (code 255)
                hM[t] = key[t] >> 8;
                acc = max(acc, hM[t] - gc[t]);
                run[t] = acc;
            }
            const int carryX = (lane == 0 && w > 0) ? (edgeIn >> 8) - g * (c0 - 1) : POA_NEG;
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
            // the chain continues with h8; everything below is off it
            if (lane == 31) edge[w * K5R_EDGE + (i & (K5R_EDGE - 1))] = ((unsigned long long)(uint32_t)i << 32) | (uint32_t)h8.w;
            myring[(size_t)(i & RM) * 32 + lane] = h8;
            const int rowmax = max(max(Hv[0], Hv[1]), max(Hv[2], Hv[3]));
            if (MODE) {
                const int tl = L - c0;
                if (tl >= 0 && tl < K5R_CPL && (m.x & K5R_FLAG_SINK)) {
                    const int h = tl == 0 ? Hv[0] : (tl == 1 ? Hv[1] : (tl == 2 ? Hv[2] : Hv[3]));
                    if (h > sinkv) { sinkv = h; sinki = i; }
                }
            } else if (rowmax > bestv) {
                bestv = rowmax; besti = i;
                bestj = c0 + (Hv[0] == rowmax ? 0 : (Hv[1] == rowmax ? 1 : (Hv[2] == rowmax ? 2 : 3)));
            }
            *reinterpret_cast<uint32_t *>(dirp) = ~codes;
            if (m.x & K5R_FLAG_STORE) {                      // read back from global memory by a far successor (rare)
                __stcg(reinterpret_cast<int4 *>(hgp), h8);
                __threadfence_block();
            }
            if (lane == 0) progress[w] = i;
            dirp += ld; hgp += ld;
            prevH[0] = h8.x; prevH[1] = h8.y; prevH[2] = h8.z; prevH[3] = h8.w;
            prevL = edgeIn;
            m = m_n; pl = pl_n; pa = pa_n; pb = pb_n; la = la_n; lb = lb_n;
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
    if (lane == 0) { s_best[w][0] = bestv; s_best[w][1] = besti; s_best[w][2] = bestj; if (err) A.out[blockIdx.x * 16 + 2] = err; }
    if (w == 1 && lane == 0) {
        A.out[blockIdx.x * 16 + 6] = (int)(t_wait >> 10); A.out[blockIdx.x * 16 + 8] = (int)(t_bp >> 10);
        A.out[blockIdx.x * 16 + 9] = (int)(t_pred >> 10); A.out[blockIdx.x * 16 + 10] = (int)(t_scan >> 10);
        A.out[blockIdx.x * 16 + 11] = (int)(t_store >> 10);
    }
    __threadfence();
    __syncthreads();
    if (w != 0) return;
    const long long t_tb = clock64();
    if (lane == 0) { A.out[blockIdx.x * 16 + 3] = (int)((t_tb - t_start) >> 10); A.out[blockIdx.x * 16 + 5] = far_rows; }
    // ---- traceback by warp 0: 32 x 32 tiles of direction bytes + row records in shared memory
    int bv = s_best[0][0], bi_ = s_best[0][1], bj_ = s_best[0][2];
    for (int x = 1; x < NW; ++x) {
        const int ov = s_best[x][0], oi = s_best[x][1], oj = s_best[x][2];
        if (ov > bv || (ov == bv && (oi < bi_ || (oi == bi_ && oj < bj_)))) { bv = ov; bi_ = oi; bj_ = oj; }
    }
    uint8_t *tdir = reinterpret_cast<uint8_t *>(ringb);          // the rings are free now
    uint4 *tmeta = reinterpret_cast<uint4 *>(tdir + 32 * 32);
    int2 *path = A.path + J.path_off;
    int n = 0;
    int i = bi_, j = bj_;
    bool go = !(MODE == 0 && bv <= 0) && V > 0 && i >= 1 && i <= V;
    if (MODE && V > 0 && (i < 1 || i > V)) { if (lane == 0) A.out[blockIdx.x * 16 + 2] = 8; go = false; }    // no sink reached
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
                else if (!k5r_is_inline(mt, np)) pr = __ldg(ovf + mt.y + u);
                else pr = k5r_inline_pred(mt, u);
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
    if (lane == 0) {
        A.out[blockIdx.x * 16] = n; A.out[blockIdx.x * 16 + 1] = bv;
        A.out[blockIdx.x * 16 + 4] = (int)((clock64() - t_tb) >> 10);
    }
}

static inline size_t k5r_smem_bytes(int n_warps, int ring)
{
    const size_t fixed = (size_t)(K5R_MAXW + 2 * K5R_MAXW * K5R_EDGE) * 4;
    const size_t rings = (size_t)n_warps * ring * 32 * 16 + (size_t)n_warps * 32 * 24;     // H rings + row-record and plan blocks
    return fixed + std::max<size_t>(rings, 32 * 32 + 32 * 16);
}
