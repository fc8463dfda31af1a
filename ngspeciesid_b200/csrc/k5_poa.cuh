// K5: partial-order-alignment consensus, one CTA per job (a cluster draft or a polishing window).
// Replaces the spoa / racon window arithmetic behind consensus.run_spoa (modules/consensus.py:83-92)
// and consensus.run_racon (modules/consensus.py:107-126).
//
// Layers of a job are added in order. For each layer the DP over (graph nodes in topological
// order) x (layer bases) is computed row by row by the whole CTA: every thread owns a contiguous
// chunk of columns, takes the maximum over the predecessor rows (diagonal and vertical moves) and
// the horizontal gap chain H[r][j] = max(A[j], H[r][j-1] + g) is resolved with a block-wide
// prefix maximum of A[j] - g*j. Traceback, graph update, topological sort and the final heaviest
// bundle are sequential and run on thread 0 (poa_core.cuh). The DP matrix lives in global memory
// (L2 resident for amplicon-sized jobs), one arena per CTA slot.
#pragma once
#include "ngsid_internal.cuh"
#include "poa_core.cuh"

#define K5_THREADS 256
#define K5_MAXC 16                 // columns per thread: layers up to 4096 bases

struct K5Args {
    int64_t n_jobs;
    const int64_t *job_off;
    const int32_t *layer_src, *layer_begin, *layer_len;
    const uint8_t *seq, *qual; const int64_t *off;
    const uint8_t *aux; const int64_t *aoff;
    int mode, m, x, g, trim;
    int order_mode;             // 0: spoa's DFS re-sort after every layer, 1: path insertion
    uint8_t *arena; size_t graph_bytes;
    int Vcap, Ecap, Acap, Scap, Lmax, ring_rows;
    int32_t *H; size_t h_words;
    int4 *rmeta_all; uint32_t *rinfo_all;       // per slot: Vcap entries each
    uint8_t *out; int64_t out_stride; int32_t *out_len; int32_t *out_nodes; int32_t *err;
    long long *cycles;          // optional: per job {dp, traceback, graph update + sort, consensus} clock64 sums
};

template <int CMAX>
__global__ void __launch_bounds__(K5_THREADS) k5_poa_kernel(K5Args A)
{
    __shared__ int4 s_meta[K5_THREADS];
    __shared__ uint32_t s_info[K5_THREADS];
    __shared__ PoaGraph G;
    __shared__ int s_wmax[2][K5_THREADS / 32];
    extern __shared__ __align__(16) int rowbuf[];        // ring of ring_rows DP rows of Lmax+2 ints
    __shared__ int s_best[K5_THREADS / 32][3];
    __shared__ int s_bi, s_bj, s_go;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t *H = A.H + (size_t)blockIdx.x * A.h_words;
    int4 *rmeta = A.rmeta_all + (size_t)blockIdx.x * A.Vcap;
    uint32_t *rinfo = A.rinfo_all + (size_t)blockIdx.x * A.Vcap;

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
                if (tid == 0) poa_add_alignment(G, 0, s, q, L, A.order_mode);
                __syncthreads();
                continue;
            }
            const size_t ld = (size_t)L + 1;
            const int g = A.g;
            if (tid == 0) t0 = clock64();
            // ---- per-row metadata (letter + up to 4 predecessor rows), built in parallel
            for (int r = tid; r < V; r += K5_THREADS) {
                const int v = G.order[r];
                int4 pm = make_int4(-1, -1, -1, -1);
                int np = 0;
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e], ++np) {
                    const int pr = G.rank[G.e_from[e]] + 1;
                    if (np == 0) pm.x = pr; else if (np == 1) pm.y = pr; else if (np == 2) pm.z = pr; else if (np == 3) pm.w = pr;
                }
                if (np == 0) pm.x = 0;                      // source node: virtual row 0
                rmeta[r] = pm;
                rinfo[r] = ((uint32_t)np << 8) | G.letter[v];
            }
            for (int j = tid; j <= L; j += K5_THREADS) { const int h = A.mode ? j * g : 0; __stcg(&H[j], h); rowbuf[j] = h; }
            const int C = (L + K5_THREADS - 1) / K5_THREADS;          // <= CMAX by construction
            const int j0 = 1 + tid * C, j1 = min(L, j0 + C - 1);
            int bestv = 0, besti = 0, bestj = 0;
            __syncthreads();
            // ring of the most recent DP rows in shared memory: matrix row q lives in slot q % R
            const int R = A.ring_rows, rstride = A.Lmax + 2;
            uint8_t sreg[CMAX];
#pragma unroll
            for (int t = 0; t < CMAX; ++t) sreg[t] = (j0 + t <= j1) ? s[j0 + t - 1] : 0;
            for (int rb = 0; rb < V; rb += K5_THREADS) {
                __syncthreads();
                if (rb + tid < V) { s_meta[tid] = rmeta[rb + tid]; s_info[tid] = rinfo[rb + tid]; }
                __syncthreads();
                const int rend = min(V, rb + K5_THREADS);
                for (int r = rb; r < rend; ++r) {
                    const int4 pm = s_meta[r - rb];
                    const uint32_t info = s_info[r - rb];
                    const uint8_t c = (uint8_t)(info & 255u);
                    const int np = (int)(info >> 8);
                    const int mrow = r + 1;
                    const size_t row = (size_t)mrow * ld;
                    int *curb = rowbuf + (size_t)(mrow % R) * rstride;
                    int h0 = 0;
                    int run, val[CMAX];
                    if (np <= 4) {
                        const int prs[4] = {pm.x, pm.y, pm.z, pm.w};
                        // gather: every predecessor value first (ring or L2), then the maxima
                        int pa[4][CMAX + 1];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int pr = prs[u];
                            if (pr >= 0) {
                                if (mrow - pr < R) {
                                    const int *pb = rowbuf + (size_t)(pr % R) * rstride;
#pragma unroll
                                    for (int t = 0; t <= CMAX; ++t) pa[u][t] = (j0 + t - 1 <= j1) ? pb[j0 + t - 1] : POA_NEG;
                                    if (A.mode && tid == 0) h0 = (u == 0) ? pb[0] : max(h0, pb[0]);
                                } else {
                                    const int32_t *hp = H + (size_t)pr * ld;
#pragma unroll
                                    for (int t = 0; t <= CMAX; ++t) pa[u][t] = (j0 + t - 1 <= j1) ? __ldcg(hp + j0 + t - 1) : POA_NEG;
                                    if (A.mode && tid == 0) { const int z = __ldcg(hp); h0 = (u == 0) ? z : max(h0, z); }
                                }
                            } else {
#pragma unroll
                                for (int t = 0; t <= CMAX; ++t) pa[u][t] = POA_NEG;
                            }
                        }
                        if (A.mode) h0 += g;
                        run = (tid == 0) ? h0 : POA_NEG;
#pragma unroll
                        for (int t = 0; t < CMAX; ++t) {
                            const int j = j0 + t;
                            val[t] = POA_NEG;
                            if (j <= j1) {
                                const int sc = (c == sreg[t]) ? A.m : A.x;
                                int dg = max(max(pa[0][t], pa[1][t]), max(pa[2][t], pa[3][t]));
                                int up = max(max(pa[0][t + 1], pa[1][t + 1]), max(pa[2][t + 1], pa[3][t + 1]));
                                int h = max(dg + sc, up + g);
                                if (!A.mode) h = max(h, 0);
                                run = max(run, h - g * j);
                                val[t] = run;
                            }
                        }
                    } else {
                        const int v = G.order[r];
                        if (A.mode) {
                            int p = POA_NEG;
                            for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) p = max(p, __ldcg(&H[(size_t)(G.rank[G.e_from[e]] + 1) * ld]));
                            h0 = p + g;
                        }
                        run = (tid == 0) ? h0 : POA_NEG;
#pragma unroll
                        for (int t = 0; t < CMAX; ++t) {
                            const int j = j0 + t;
                            val[t] = POA_NEG;
                            if (j <= j1) {
                                const int sc = (c == sreg[t]) ? A.m : A.x;
                                int h = POA_NEG;
                                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                                    const int32_t *hp = H + (size_t)(G.rank[G.e_from[e]] + 1) * ld + j;
                                    h = max(h, max(__ldcg(hp - 1) + sc, __ldcg(hp) + g));
                                }
                                if (!A.mode) h = max(h, 0);
                                run = max(run, h - g * j);
                                val[t] = run;
                            }
                        }
                    }
                    int incl = run;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        int o = __shfl_up_sync(NGSID_FULL_MASK, incl, d);
                        if (lane >= d) incl = max(incl, o);
                    }
                    if (lane == 31) s_wmax[r & 1][wid] = incl;
                    __syncthreads();
                    int carry = __shfl_up_sync(NGSID_FULL_MASK, incl, 1);
                    if (lane == 0) carry = POA_NEG;
                    {
                        int wt = (lane < K5_THREADS / 32) ? s_wmax[r & 1][lane] : POA_NEG;
#pragma unroll
                        for (int d = 1; d < K5_THREADS / 32; d <<= 1) {
                            int o = __shfl_up_sync(NGSID_FULL_MASK, wt, d);
                            if (lane >= d) wt = max(wt, o);
                        }
                        int prevw = __shfl_sync(NGSID_FULL_MASK, wt, max(wid - 1, 0));
                        if (wid > 0) carry = max(carry, prevw);
                    }
#pragma unroll
                    for (int t = 0; t < CMAX; ++t) {
                        const int j = j0 + t;
                        if (j <= j1) {
                            const int hv = max(val[t], carry) + g * j;
                            curb[j] = hv;
                            __stcg(&H[row + j], hv);
                            if (!A.mode && hv > bestv) { bestv = hv; besti = r + 1; bestj = j; }
                        }
                    }
                    if (tid == 0) { curb[0] = h0; __stcg(&H[row], h0); }
                    __syncthreads();
                }
            }
            // ---- end cell
            if (!A.mode) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    int ov = __shfl_xor_sync(NGSID_FULL_MASK, bestv, d);
                    int oi = __shfl_xor_sync(NGSID_FULL_MASK, besti, d);
                    int oj = __shfl_xor_sync(NGSID_FULL_MASK, bestj, d);
                    if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
                }
                if (lane == 0) { s_best[wid][0] = bestv; s_best[wid][1] = besti; s_best[wid][2] = bestj; }
                __syncthreads();
            }
            if (tid == 0) {
                long long t1 = clock64();
                cyc_dp += t1 - t0;
                int bv, bi = 0, bj = 0;
                if (!A.mode) {
                    bv = 0;
                    for (int w = 0; w < K5_THREADS / 32; ++w) {
                        const int ov = s_best[w][0], oi = s_best[w][1], oj = s_best[w][2];
                        if (ov > bv || (ov == bv && ov > 0 && (oi < bi || (oi == bi && oj < bj)))) { bv = ov; bi = oi; bj = oj; }
                    }
                } else {
                    bv = POA_NEG;
                    for (int r = 0; r < V; ++r) {
                        const int v = G.order[r];
                        if (G.out_head[v] < 0) {
                            const int hv = __ldcg(&H[(size_t)(r + 1) * ld + L]);
                            if (hv > bv) { bv = hv; bi = r + 1; bj = L; }
                        }
                    }
                }
                int n_aln = 0;
                if (!(A.mode == 0 && bv == 0)) n_aln = poa_traceback(G, H, ld, s, A.mode, A.m, A.x, g, bi, bj);
                long long t2 = clock64();
                cyc_tb += t2 - t1;
                poa_add_alignment(G, n_aln, s, q, L, A.order_mode);
                cyc_add += clock64() - t2;
                s_go = (G.err == 0 && G.V + A.Lmax + 2 < G.Vcap) ? 1 : 0;
                if (!s_go && G.err == 0) G.err = 1;
            }
            __syncthreads();
            if (!s_go) break;
        }
        if (tid == 0) {
            int len = -1;
            long long t3 = clock64();
            if (G.err == 0) len = poa_consensus(G, A.trim, A.out + (size_t)job * A.out_stride, (int)A.out_stride);
            cyc_cons = clock64() - t3;
            if (A.cycles) { A.cycles[job * 4] = cyc_dp; A.cycles[job * 4 + 1] = cyc_tb; A.cycles[job * 4 + 2] = cyc_add; A.cycles[job * 4 + 3] = cyc_cons; }
            A.out_len[job] = len;
            if (A.out_nodes) A.out_nodes[job] = G.V;
            if (G.err) atomicMax(A.err, G.err);
            else if (len < 0) atomicMax(A.err, 5);
        }
        __syncthreads();
    }
}
