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

struct K5Args {
    int64_t n_jobs;
    const int64_t *job_off;
    const int32_t *layer_src, *layer_begin, *layer_len;
    const uint8_t *seq, *qual; const int64_t *off;
    const uint8_t *aux; const int64_t *aoff;
    int mode, m, x, g, trim;
    uint8_t *arena; size_t graph_bytes;
    int Vcap, Ecap, Acap, Scap, Lmax;
    int32_t *H; size_t h_words;
    uint8_t *out; int64_t out_stride; int32_t *out_len; int32_t *out_nodes; int32_t *err;
};

__global__ void __launch_bounds__(K5_THREADS) k5_poa_kernel(K5Args A)
{
    __shared__ PoaGraph G;
    __shared__ int s_wmax[K5_THREADS / 32];
    __shared__ int s_best[K5_THREADS / 32][3];
    __shared__ int s_bi, s_bj, s_go;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t *H = A.H + (size_t)blockIdx.x * A.h_words;

    for (int64_t job = blockIdx.x; job < A.n_jobs; job += gridDim.x) {
        if (tid == 0) poa_graph_bind(G, A.arena + (size_t)blockIdx.x * A.graph_bytes, A.Vcap, A.Ecap, A.Acap, A.Scap, A.Lmax);
        __syncthreads();
        for (int64_t li = A.job_off[job]; li < A.job_off[job + 1]; ++li) {
            const int src = A.layer_src[li], lb = A.layer_begin[li], L = A.layer_len[li];
            const uint8_t *s = (src >= 0 ? A.seq + A.off[src] : A.aux + A.aoff[-src - 1]) + lb;
            const uint8_t *q = src >= 0 ? A.qual + A.off[src] + lb : nullptr;
            const int V = G.V;
            if (V == 0 || L == 0) {
                if (tid == 0) poa_add_alignment(G, 0, s, q, L);
                __syncthreads();
                continue;
            }
            const size_t ld = (size_t)L + 1;
            const int g = A.g;
            for (int j = tid; j <= L; j += K5_THREADS) H[j] = A.mode ? j * g : 0;
            const int C = (L + K5_THREADS - 1) / K5_THREADS;
            const int j0 = 1 + tid * C, j1 = min(L, j0 + C - 1);
            int bestv = 0, besti = 0, bestj = 0;
            __syncthreads();
            for (int r = 0; r < V; ++r) {
                const int v = G.order[r];
                const uint8_t c = G.letter[v];
                const size_t row = (size_t)(r + 1) * ld;
                const int eh = G.in_head[v];
                // column 0
                int h0 = 0;
                if (A.mode) {
                    int p = POA_NEG;
                    if (eh < 0) p = 0;
                    for (int e = eh; e >= 0; e = G.e_next_in[e]) p = max(p, H[(size_t)(G.rank[G.e_from[e]] + 1) * ld]);
                    h0 = p + g;
                }
                // A[j] - g*j over this thread's chunk, running prefix maximum
                int run = (tid == 0) ? h0 : POA_NEG;
                for (int j = j0; j <= j1; ++j) {
                    const int sc = (c == s[j - 1]) ? A.m : A.x;
                    int h = POA_NEG;
                    if (eh < 0) h = max(H[j - 1] + sc, H[j] + g);
                    for (int e = eh; e >= 0; e = G.e_next_in[e]) {
                        const size_t pr = (size_t)(G.rank[G.e_from[e]] + 1) * ld;
                        h = max(h, max(H[pr + j - 1] + sc, H[pr + j] + g));
                    }
                    if (!A.mode) h = max(h, 0);
                    run = max(run, h - g * j);
                    H[row + j] = run;                      // provisional: chunk-local prefix maximum
                }
                // exclusive prefix maximum of the chunk maxima across the block
                int incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int o = __shfl_up_sync(NGSID_FULL_MASK, incl, d);
                    if (lane >= d) incl = max(incl, o);
                }
                if (lane == 31) s_wmax[wid] = incl;
                __syncthreads();
                int carry = __shfl_up_sync(NGSID_FULL_MASK, incl, 1);
                if (lane == 0) carry = POA_NEG;
                for (int w = 0; w < wid; ++w) carry = max(carry, s_wmax[w]);
                for (int j = j0; j <= j1; ++j) {
                    const int hv = max(H[row + j], carry) + g * j;
                    H[row + j] = hv;
                    if (!A.mode && hv > bestv) { bestv = hv; besti = r + 1; bestj = j; }
                }
                if (tid == 0) H[row] = h0;
                __syncthreads();
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
                            const int hv = H[(size_t)(r + 1) * ld + L];
                            if (hv > bv) { bv = hv; bi = r + 1; bj = L; }
                        }
                    }
                }
                int n_aln = 0;
                if (!(A.mode == 0 && bv == 0)) n_aln = poa_traceback(G, H, ld, s, A.mode, A.m, A.x, g, bi, bj);
                poa_add_alignment(G, n_aln, s, q, L);
                s_go = (G.err == 0 && G.V + A.Lmax + 2 < G.Vcap) ? 1 : 0;
                if (!s_go && G.err == 0) G.err = 1;
            }
            __syncthreads();
            if (!s_go) break;
        }
        if (tid == 0) {
            int len = -1;
            if (G.err == 0) len = poa_consensus(G, A.trim, A.out + (size_t)job * A.out_stride, (int)A.out_stride);
            A.out_len[job] = len;
            if (A.out_nodes) A.out_nodes[job] = G.V;
            if (G.err) atomicMax(A.err, G.err);
            else if (len < 0) atomicMax(A.err, 5);
        }
        __syncthreads();
    }
}
