// K5 (wavefront shape): partial-order-alignment consensus, one thread block per job.
// Replaces the spoa / racon window arithmetic behind consensus.run_spoa (modules/consensus.py:83-92)
// and consensus.run_racon (modules/consensus.py:107-126); graph semantics in poa_core.cuh.
//
// DP of one layer (V graph rows in topological order x L+1 columns):
//   * thread t owns the rows [t*RPT, (t+1)*RPT) and in step s works on column j = s - t of each of
//     them, top to bottom: a cell needs its predecessors' columns j-1 and j, and every predecessor
//     row has a smaller rank, so it either sits higher in the same thread (done earlier in this
//     step) or in an earlier thread (done in an earlier step). One __syncthreads per step; the
//     linear gap makes the horizontal move a plain dependency on the row's own previous column.
//   * the last D columns of every row live in a shared-memory ring (hist[r][j mod D]); a
//     predecessor d threads back is D-2 >= d columns ahead at most, so its columns j-1 and j are
//     still in the ring and not the slot it writes in this step. Rarer, more distant predecessors
//     (and rows with more than 4 of them) read the matrix in global memory; only the rows that are
//     read that way write it (scattered 4-byte stores of every cell cost 10x the whole DP).
//   * every cell stores one direction byte (which in-edge, diagonal / vertical, or horizontal) with
//     the tie order of poa_traceback (diagonal in-edges in list order, then vertical, then
//     horizontal), so the traceback is one byte per step instead of re-deriving the move.
// Graph update stays sequential on thread 0 (poa_core.cuh). The topological order is spoa's
// depth-first re-sort after every layer (poa_topo_sort, order_mode 0 of the oracle); thread 0 runs
// it on shared-memory copies of the in-edge and aligned-node lists (the DP ring is free at that
// point), which the block copies in and out in parallel.
#pragma once
#include "ngsid_internal.cuh"
#include "poa_core.cuh"
#include "k5_poa.cuh"

#define K5W_THREADS 512
#define K5W_DIAG 0           // direction byte: 0..119 diagonal through in-edge u
#define K5W_UP 120           // 120..239 vertical through in-edge u - 120
#define K5W_LEFT 254
#define K5W_STOP 255
#define K5W_MAXE 119
#define K5W_FAST 12          // in-edges a row keeps in its shared-memory metadata

struct K5WArgs {
    K5Args a;                // job description, graph arena, H matrix, outputs (as the row kernel)
    uint8_t *dir; size_t dir_bytes;          // per slot: (Vcap+1) x (Lmax+1) direction bytes
    int smem_words;          // dynamic shared memory available for hist + meta (32-bit words)
};

__device__ __forceinline__ uint32_t k5w_opaque(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
template <class T>
__device__ __forceinline__ T *k5w_opaque_ptr(T *p)
{
    unsigned long long v = (unsigned long long)p, r;
    asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(v));
    return (T *)r;
}
__device__ __forceinline__ uint4 k5w_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(K5W_THREADS, 1) k5w_poa_kernel(K5WArgs W)
{
    const K5Args &A = W.a;
    __shared__ PoaGraph G;
    __shared__ int s_best[K5W_THREADS / 32][3];
    __shared__ int s_go, s_D, s_naln;
    extern __shared__ __align__(16) int k5w_smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t *H = A.H + (size_t)blockIdx.x * A.h_words;
    uint8_t *DIR = W.dir + (size_t)blockIdx.x * W.dir_bytes;
    int32_t *prank = reinterpret_cast<int32_t *>(A.rmeta_all) + (size_t)blockIdx.x * A.Vcap * K5W_FAST;   // rows of far in-edges
    uint8_t *sseq = reinterpret_cast<uint8_t *>(k5w_smem);            // layer bases (Lmax + 8 bytes)
    const int seq_words = ((A.Lmax + 8 + 15) / 16) * 4;
    int *dyn = k5w_smem + seq_words;
    const int dyn_words = W.smem_words - seq_words - 4;

    for (int64_t job = blockIdx.x; job < A.n_jobs; job += gridDim.x) {
        if (tid == 0) poa_graph_bind(G, A.arena + (size_t)blockIdx.x * A.graph_bytes, A.Vcap, A.Ecap, A.Acap, A.Scap, A.Lmax);
        long long cyc_dp = 0, cyc_tb = 0, cyc_add = 0, cyc_cons = 0, t0 = 0;
        __syncthreads();
        for (int64_t li = A.job_off[job]; li < A.job_off[job + 1]; ++li) {
            const int src = A.layer_src[li], lb = A.layer_begin[li], L = A.layer_len[li];
            const uint8_t *s = (src >= 0 ? A.seq + A.off[src] : A.aux + A.aoff[-src - 1]) + lb;
            const uint8_t *q = src >= 0 ? A.qual + A.off[src] + lb : nullptr;
            const int V = G.V;
            if (V == 0 || L == 0) {
                if (tid == 0) poa_add_alignment(G, 0, s, q, L, 0);
                __syncthreads();
                continue;
            }
            if (tid == 0) t0 = clock64();
            const int g = A.g;
            constexpr int mode = MODE;                          // 0 local (draft), 1 global (polishing windows)
            const size_t ld = (size_t)L + 1;
            const int RPT = ((V + K5W_THREADS - 1) / K5W_THREADS) | 1;      // odd: lanes of a warp hit distinct banks
            // ring rows: V graph rows | one virtual-source row per thread | one row of -infinity.
            // ring depth: the deepest power of two (8, 4, 2) whose ring + metadata fit
            const int NR = V + K5W_THREADS + 1;
            int D = 8;
            while (D > 2 && (size_t)NR * (size_t)(D + 1) + (size_t)V * 8 + 16 > (size_t)dyn_words) D >>= 1;
            if ((size_t)NR * (size_t)(D + 1) + (size_t)V * 8 + 16 > (size_t)dyn_words || NR >= 0xffff) {
                if (tid == 0) G.err = 6;                        // graph too large for this shape
                __syncthreads();
                break;
            }
            const int DS = D + 1;                               // row stride of the ring (odd)
            int *hist = dyn;                                    // NR x DS
            uint4 *smeta = reinterpret_cast<uint4 *>(dyn + (((size_t)NR * DS + 3) & ~(size_t)3));   // 2 x uint4 per row
            for (int j = tid; j < L; j += K5W_THREADS) sseq[j] = s[j];
            for (int x = tid; x < DS; x += K5W_THREADS) hist[(size_t)(V + K5W_THREADS) * DS + x] = POA_NEG;
            // ---- row metadata (32 bytes): letter | in-edges << 8 | flags, then up to 12 in-edges
            // as 16-bit ring rows: the predecessor's row, or this thread's virtual-source row, or
            // the -infinity row for unused slots; 0xffff = not reachable through the ring -> global
            // matrix (its row number goes to prank)
            for (int r = tid; r < V; r += K5W_THREADS) {
                const int v = G.order[r];
                const int tr = r / RPT;
                const uint32_t none = (uint32_t)(V + K5W_THREADS);
                uint32_t pw[K5W_FAST];
#pragma unroll
                for (int u = 0; u < K5W_FAST; ++u) pw[u] = none;
                int np = 0;
                bool has_far = false;
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e], ++np) {
                    const int prow = G.rank[G.e_from[e]];       // 0-based row of the predecessor
                    const bool near = prow < r && (tr - prow / RPT) <= D - 2;
#pragma unroll
                    for (int u = 0; u < K5W_FAST; ++u) if (u == np) pw[u] = near ? (uint32_t)prow : 0xffffu;
                    if (np < K5W_FAST && !near) { prank[(size_t)r * K5W_FAST + np] = prow + 1; has_far = true; }
                }
                if (np == 0) { np = 1; pw[0] = (uint32_t)(V + tr); }        // source node: the virtual row 0
                if (np > K5W_MAXE) { G.err = 7; }
                const uint32_t info = (uint32_t)G.letter[v] | ((uint32_t)min(np, 255) << 8) | (np > K5W_FAST ? 0x10000u : 0u) |
                                      (G.out_head[v] < 0 ? 0x40000u : 0u) | (has_far ? 0x80000u : 0u);
                smeta[2 * r] = make_uint4(info, pw[0] | (pw[1] << 16), pw[2] | (pw[3] << 16), pw[4] | (pw[5] << 16));
                smeta[2 * r + 1] = make_uint4(pw[6] | (pw[7] << 16), pw[8] | (pw[9] << 16), pw[10] | (pw[11] << 16), 0u);
            }
            __syncthreads();
            // rows that some other row reads through the global matrix keep writing it (bit 17)
            for (int r = tid; r < V; r += K5W_THREADS) {
                const uint4 m = smeta[2 * r], m2 = smeta[2 * r + 1];
                const bool slow = (m.x & 0x10000u) != 0;
                const int np = (int)((m.x >> 8) & 255u);
                const uint32_t pw[6] = {m.y, m.z, m.w, m2.x, m2.y, m2.z};
                bool any_far = slow;
#pragma unroll
                for (int u = 0; u < K5W_FAST; ++u) any_far |= (u < np) && ((pw[u >> 1] >> (16 * (u & 1))) & 0xffffu) == 0xffffu;
                if (any_far) {
                    const int v = G.order[r];
                    int u = 0;
                    for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e], ++u) {
                        bool far = slow;
#pragma unroll
                        for (int k = 0; k < K5W_FAST; ++k) if (k == u) far |= ((pw[k >> 1] >> (16 * (k & 1))) & 0xffffu) == 0xffffu;
                        if (far) atomicOr(&smeta[2 * G.rank[G.e_from[e]]].x, 0x20000u);
                    }
                }
            }
            __syncthreads();
            int bestv = 0, besti = 0, bestj = 0;
            int sinkv = POA_NEG, sinki = 0x7fffffff;            // global mode: best sink row at column L
            const int r_begin = tid * RPT, r_end = min(V, r_begin + RPT);
            const int n_steps = L + 1 + (V + RPT - 1) / RPT - 1;      // the last thread that owns rows finishes column L
            const int dmask = D - 1;
            // hot loop: 32-bit shared-window addresses, running pointers per row
            // (made opaque: otherwise the compiler rebuilds the shared-window base from the CTA id
            // next to every access instead of keeping it in a register)
            const uint32_t s_hist = k5w_opaque((uint32_t)__cvta_generic_to_shared(hist));
            const uint32_t s_meta = k5w_opaque((uint32_t)__cvta_generic_to_shared(smeta));
            const uint32_t row_b = k5w_opaque(4u * (uint32_t)DS);            // bytes per ring row
            const uint32_t a_virt = s_hist + (uint32_t)(V + tid) * row_b;
            // scoring constants and output bases in registers for the whole layer (otherwise every
            // cell re-reads them from the constant bank and rebuilds the slot bases from blockIdx)
            const int sc_m = (int)k5w_opaque((uint32_t)A.m), sc_x = (int)k5w_opaque((uint32_t)A.x);
            const int gq = (int)k5w_opaque((uint32_t)g);
            uint8_t *const DIRq = k5w_opaque_ptr(DIR);
            int32_t *const Hq = k5w_opaque_ptr(H);
            for (int st = 0; st < n_steps; ++st) {
                const int j = st - tid;
                if (j >= 0 && j <= L && r_begin < r_end) {
                    const uint32_t cj = (j >= 1) ? sseq[j - 1] : 0xffffu;
                    const uint32_t so = 4u * (uint32_t)(j & dmask), so1 = 4u * (uint32_t)((j - 1) & dmask);
                    k1s_sts32(a_virt + so, (uint32_t)(mode ? j * gq : 0));          // virtual row 0, column j
                    uint32_t a_hist = s_hist + (uint32_t)r_begin * row_b;
                    uint32_t a_meta = s_meta + 32u * (uint32_t)r_begin;
                    const size_t goff0 = (size_t)(r_begin + 1) * ld + (size_t)j;
                    uint8_t *dirp = DIRq + goff0;
                    int32_t *hout = Hq + goff0;
                    for (int r = r_begin; r < r_end; ++r, a_hist += row_b, a_meta += 32u, dirp += ld, hout += ld) {
                        const uint4 m = k5w_lds128(a_meta);
                        const int np = (int)((m.x >> 8) & 255u);
                        const int sc = ((m.x & 255u) == cj) ? sc_m : sc_x;
                        int best = POA_NEG, bdir = K5W_STOP;
                        int bup = POA_NEG, udir = 0;
                        // in-edge u of this row: values of the predecessor row at columns j-1 and j
#define K5W_PRED(idx, u)                                                                            \
                        {                                                                           \
                            int pa, pb;                                                             \
                            if ((idx) != 0xffffu) {                                                 \
                                const uint32_t hp = s_hist + (idx) * row_b;                         \
                                pa = (int)k1s_lds32(hp + so1); pb = (int)k1s_lds32(hp + so);        \
                            } else {                                                                \
                                const int32_t *hp = Hq + (size_t)prank[(size_t)r * K5W_FAST + (u)] * ld + j; \
                                pa = (j >= 1) ? hp[-1] : POA_NEG;                                   \
                                pb = hp[0];                                                         \
                            }                                                                       \
                            if (pa > best) { best = pa; bdir = K5W_DIAG + (u); }                    \
                            if (pb > bup) { bup = pb; udir = K5W_UP + (u); }                        \
                        }
                        // all in-edges in the ring (the common case): no per-edge test for far rows
#define K5W_PREDN(idx, u)                                                                           \
                        {                                                                           \
                            const uint32_t hp = s_hist + (idx) * row_b;                             \
                            const int pa = (int)k1s_lds32(hp + so1), pb = (int)k1s_lds32(hp + so);  \
                            if (pa > best) { best = pa; bdir = K5W_DIAG + (u); }                    \
                            if (pb > bup) { bup = pb; udir = K5W_UP + (u); }                        \
                        }
                        if (!(m.x & 0x90000u)) {
                            K5W_PREDN(m.y & 0xffffu, 0)
                            K5W_PREDN(m.y >> 16, 1)                          // unused slots point at the -infinity row
                            if (np > 2) {
                                K5W_PREDN(m.z & 0xffffu, 2)
                                K5W_PREDN(m.z >> 16, 3)
                                if (np > 4) {
                                    K5W_PREDN(m.w & 0xffffu, 4)
                                    K5W_PREDN(m.w >> 16, 5)
                                    if (np > 6) {
                                        const uint4 m2 = k5w_lds128(a_meta + 16u);
                                        K5W_PREDN(m2.x & 0xffffu, 6)
                                        K5W_PREDN(m2.x >> 16, 7)
                                        if (np > 8) {
                                            K5W_PREDN(m2.y & 0xffffu, 8)
                                            K5W_PREDN(m2.y >> 16, 9)
                                            K5W_PREDN(m2.z & 0xffffu, 10)
                                            K5W_PREDN(m2.z >> 16, 11)
                                        }
                                    }
                                }
                            }
                        } else if (!(m.x & 0x10000u)) {
                            K5W_PRED(m.y & 0xffffu, 0)
                            K5W_PRED(m.y >> 16, 1)                           // unused slots point at the -infinity row
                            if (np > 2) {
                                K5W_PRED(m.z & 0xffffu, 2)
                                K5W_PRED(m.z >> 16, 3)
                                if (np > 4) {
                                    K5W_PRED(m.w & 0xffffu, 4)
                                    K5W_PRED(m.w >> 16, 5)
                                    if (np > 6) {
                                        const uint4 m2 = k5w_lds128(a_meta + 16u);
                                        K5W_PRED(m2.x & 0xffffu, 6)
                                        K5W_PRED(m2.x >> 16, 7)
                                        if (np > 8) {
                                            K5W_PRED(m2.y & 0xffffu, 8)
                                            K5W_PRED(m2.y >> 16, 9)
                                            K5W_PRED(m2.z & 0xffffu, 10)
                                            K5W_PRED(m2.z >> 16, 11)
                                        }
                                    }
                                }
                            }
                        } else {
                            const int v = G.order[r];
                            int u = 0;
                            for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e], ++u) {
                                const int32_t *hp = H + (size_t)(G.rank[G.e_from[e]] + 1) * ld + j;
                                const int pa = (j >= 1) ? hp[-1] : POA_NEG;
                                const int pb = hp[0];
                                if (pa > best) { best = pa; bdir = K5W_DIAG + u; }
                                if (pb > bup) { bup = pb; udir = K5W_UP + u; }
                            }
                        }
#undef K5W_PRED
#undef K5W_PREDN
                        int h;
                        if (j >= 1) {
                            h = best + sc;                                   // diagonal (first maximal in-edge)
                            if (bup + gq > h) { h = bup + gq; bdir = udir; }   // vertical only if strictly better
                            const int left = (int)k1s_lds32(a_hist + so1) + gq;
                            if (left > h) { h = left; bdir = K5W_LEFT; }
                            if (!mode && h <= 0) { h = 0; bdir = K5W_STOP; }
                        } else {
                            // column 0: only vertical moves (global) or the free start (local)
                            if (mode) { h = bup + gq; bdir = udir; } else { h = 0; bdir = K5W_STOP; }
                        }
                        k1s_sts32(a_hist + so, (uint32_t)h);
                        if (m.x & 0x20000u) *hout = h;
                        *dirp = (uint8_t)bdir;
                        if (mode) {
                            if (j == L && (m.x & 0x40000u) && h > sinkv) { sinkv = h; sinki = r + 1; }
                        } else if (h >= bestv) {                 // rare: only cells on or next to the best path
                            if (h > bestv || (h > 0 && (r + 1 < besti || (r + 1 == besti && j < bestj)))) {
                                bestv = h; besti = r + 1; bestj = j;
                            }
                        }
                    }
                }
                __syncthreads();
            }
            // ---- end cell
            if (mode) { bestv = sinkv; besti = sinki; bestj = L; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
                const int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
                const int oj = __shfl_xor_sync(NGSID_FULL_MASK, bestj, d);
                if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
            }
            if (lane == 0) { s_best[wid][0] = bestv; s_best[wid][1] = besti; s_best[wid][2] = bestj; }
            __syncthreads();
            if (tid == 0) {
                const long long t1 = clock64();
                cyc_dp += t1 - t0;
                int bv = s_best[0][0], bi = s_best[0][1], bj = s_best[0][2];
                for (int w = 1; w < K5W_THREADS / 32; ++w) {
                    const int ov = s_best[w][0], oi = s_best[w][1], oj = s_best[w][2];
                    if (ov > bv || (ov == bv && (oi < bi || (oi == bi && oj < bj)))) { bv = ov; bi = oi; bj = oj; }
                }
                // ---- traceback over the direction bytes (alignment in reverse order, as poa_traceback)
                int n = 0;
                if (!(mode == 0 && bv <= 0)) {
                    int i = bi, j = bj;
                    while (mode == 0 ? i != 0 : (i != 0 || j != 0)) {
                        if (i == 0) { G.aln_node[n] = -1; G.aln_pos[n++] = j - 1; --j; continue; }   // global: along the virtual row
                        const int d = DIR[(size_t)i * ld + j];
                        if (d == K5W_STOP) break;
                        if (d == K5W_LEFT) { G.aln_node[n] = -1; G.aln_pos[n++] = j - 1; --j; continue; }
                        const int v = G.order[i - 1];
                        int u = d >= K5W_UP ? d - K5W_UP : d;
                        int pr = 0;
                        if (G.in_head[v] >= 0) {
                            int e = G.in_head[v];
                            while (u-- > 0) e = G.e_next_in[e];
                            pr = G.rank[G.e_from[e]] + 1;
                        }
                        G.aln_node[n] = v;
                        if (d >= K5W_UP) { G.aln_pos[n++] = -1; i = pr; }
                        else { G.aln_pos[n++] = j - 1; i = pr; --j; }
                    }
                }
                const long long t2 = clock64();
                cyc_tb += t2 - t1;
                poa_add_alignment(G, n, s, q, L, 2);             // graph only; the order follows below
                cyc_add += clock64() - t2;
                s_go = (G.err == 0 && G.V + A.Lmax + 2 < G.Vcap) ? 1 : 0;
                if (!s_go && G.err == 0) G.err = 1;
            }
            __syncthreads();
            // ---- topological re-sort (spoa's depth-first order) on shared-memory copies
            {
                const int V2 = G.V, E2 = G.E, A2 = G.A;
                const int by = (2 * V2 + 3) / 4;                           // mark + check bytes, in words
                const int fixed = 4 * V2 + 2 * E2 + 2 * A2 + by;
                const bool fits = G.err == 0 && fixed + 2 * V2 + 64 <= W.smem_words;
                int *w_in_head = k5w_smem, *w_al_head = w_in_head + V2, *w_order = w_al_head + V2, *w_rank = w_order + V2;
                int *w_enext = w_rank + V2, *w_efrom = w_enext + E2, *w_alnext = w_efrom + E2, *w_alnode = w_alnext + A2;
                uint8_t *w_mark = reinterpret_cast<uint8_t *>(w_alnode + A2);
                int *w_stack = w_alnode + A2 + by;
                long long t4 = 0;
                if (tid == 0) t4 = clock64();
                if (fits) {
                    for (int x = tid; x < V2; x += K5W_THREADS) { w_in_head[x] = G.in_head[x]; w_al_head[x] = G.al_head[x]; }
                    for (int x = tid; x < E2; x += K5W_THREADS) { w_enext[x] = G.e_next_in[x]; w_efrom[x] = G.e_from[x]; }
                    for (int x = tid; x < A2; x += K5W_THREADS) { w_alnext[x] = G.al_next[x]; w_alnode[x] = G.al_node[x]; }
                    __syncthreads();
                    if (tid == 0) {
                        PoaGraph Gs = G;
                        Gs.in_head = w_in_head; Gs.al_head = w_al_head; Gs.order = w_order; Gs.rank = w_rank;
                        Gs.e_next_in = w_enext; Gs.e_from = w_efrom; Gs.al_next = w_alnext; Gs.al_node = w_alnode;
                        Gs.mark = w_mark; Gs.check = w_mark + V2; Gs.stack = w_stack;
                        Gs.Scap = W.smem_words - fixed;
                        poa_topo_sort(Gs);
                        s_D = Gs.err;                                      // 4: the shared-memory stack was too small
                    }
                    __syncthreads();
                    if (s_D == 0) {
                        for (int x = tid; x < V2; x += K5W_THREADS) { G.order[x] = w_order[x]; G.rank[x] = w_rank[x]; }
                    }
                }
                if (tid == 0) {
                    if (G.err == 0 && (!fits || s_D != 0)) poa_topo_sort(G);   // global-memory fallback
                    cyc_cons += clock64() - t4;                                // reported with the consensus slot
                }
                __syncthreads();
            }
            if (!s_go) break;
        }
        if (tid == 0) {
            int len = -1;
            const long long t3 = clock64();
            if (G.err == 0) len = poa_consensus(G, A.trim, A.out + (size_t)job * A.out_stride, (int)A.out_stride);
            cyc_cons += clock64() - t3;
            if (A.cycles) { A.cycles[job * 4] = cyc_dp; A.cycles[job * 4 + 1] = cyc_tb; A.cycles[job * 4 + 2] = cyc_add; A.cycles[job * 4 + 3] = cyc_cons; }
            A.out_len[job] = len;
            if (A.out_nodes) A.out_nodes[job] = G.V;
            if (G.err) atomicMax(A.err, G.err);
            else if (len < 0) atomicMax(A.err, 5);
        }
        __syncthreads();
    }
    (void)s_D; (void)s_naln;
}
