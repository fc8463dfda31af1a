// K5 graph core: partial-order-alignment graph on flat, pre-sized arrays. The sequential parts of
// POA (graph update, topological sort, traceback, heaviest bundle) run on one thread of the CTA
// that owns the job; they are written as host/device code so that tests/ can also compile them for
// the CPU and check them against oracle/poa_oracle.cpp without a GPU (tests/poa_core_host.cpp).
// Semantics: spoa's published algorithm as restated in oracle/poa_oracle.cpp (reference call
// sites modules/consensus.py:83-92 and :107-126).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define POA_HD __host__ __device__ __forceinline__
#else
#define POA_HD inline
#endif

#define POA_NEG (-(1 << 28))

// DP matrix reads: the kernel writes the matrix with L1-bypassing stores, so reads bypass L1 too
#if defined(__CUDA_ARCH__)
#define POA_LDH(p) __ldcg(p)
#else
#define POA_LDH(p) (*(p))
#endif

struct PoaGraph {
    int32_t Vcap, Ecap, Acap, Scap;
    int32_t V, E, A, n_seqs, err;           // err: 1 node, 2 edge, 3 aligned-list, 4 stack overflow
    uint8_t *letter;
    int32_t *cover;
    int32_t *in_head, *in_tail, *out_head, *out_tail;       // edge lists per node (insertion order)
    int32_t *e_from, *e_to, *e_w, *e_next_in, *e_next_out;
    int32_t *al_head, *al_tail, *al_node, *al_next;         // aligned-node lists per node
    int32_t *order, *rank;
    uint8_t *mark, *check;
    int32_t *stack;
    int64_t *score;
    int32_t *pred;
    int32_t *aln_node, *aln_pos;                            // alignment, reverse order
};

// bytes needed for one graph with the given capacities (all arrays 16-byte aligned)
POA_HD size_t poa_align16(size_t x) { return (x + 15) & ~(size_t)15; }

POA_HD size_t poa_graph_bytes(int Vcap, int Ecap, int Acap, int Scap, int Lmax)
{
    size_t b = 0;
    b += poa_align16((size_t)Vcap);                       // letter
    b += poa_align16((size_t)Vcap * 4) * 10;              // cover,in_head,in_tail,out_head,out_tail,al_head,al_tail,order,rank,pred
    b += poa_align16((size_t)Ecap * 4) * 5;               // e_*
    b += poa_align16((size_t)Acap * 4) * 2;               // al_node, al_next
    b += poa_align16((size_t)Vcap) * 2;                   // mark, check
    b += poa_align16((size_t)Scap * 4);                   // stack
    b += poa_align16((size_t)Vcap * 8);                   // score
    b += poa_align16((size_t)(Vcap + Lmax + 8) * 4) * 2;  // aln
    return b;
}

POA_HD void poa_graph_bind(PoaGraph &G, uint8_t *mem, int Vcap, int Ecap, int Acap, int Scap, int Lmax)
{
    G.Vcap = Vcap; G.Ecap = Ecap; G.Acap = Acap; G.Scap = Scap;
    G.V = G.E = G.A = G.n_seqs = G.err = 0;
    uint8_t *p = mem;
#define POA_TAKE(field, type, count) G.field = (type *)p; p += poa_align16((size_t)(count) * sizeof(type));
    POA_TAKE(letter, uint8_t, Vcap)
    POA_TAKE(cover, int32_t, Vcap)
    POA_TAKE(in_head, int32_t, Vcap) POA_TAKE(in_tail, int32_t, Vcap)
    POA_TAKE(out_head, int32_t, Vcap) POA_TAKE(out_tail, int32_t, Vcap)
    POA_TAKE(al_head, int32_t, Vcap) POA_TAKE(al_tail, int32_t, Vcap)
    POA_TAKE(order, int32_t, Vcap) POA_TAKE(rank, int32_t, Vcap) POA_TAKE(pred, int32_t, Vcap)
    POA_TAKE(e_from, int32_t, Ecap) POA_TAKE(e_to, int32_t, Ecap) POA_TAKE(e_w, int32_t, Ecap)
    POA_TAKE(e_next_in, int32_t, Ecap) POA_TAKE(e_next_out, int32_t, Ecap)
    POA_TAKE(al_node, int32_t, Acap) POA_TAKE(al_next, int32_t, Acap)
    POA_TAKE(mark, uint8_t, Vcap) POA_TAKE(check, uint8_t, Vcap)
    POA_TAKE(stack, int32_t, Scap)
    POA_TAKE(score, int64_t, Vcap)
    POA_TAKE(aln_node, int32_t, Vcap + Lmax + 8) POA_TAKE(aln_pos, int32_t, Vcap + Lmax + 8)
#undef POA_TAKE
}

POA_HD int poa_add_node(PoaGraph &G, uint8_t c)
{
    if (G.V >= G.Vcap) { G.err = 1; return G.Vcap - 1; }
    const int v = G.V++;
    G.letter[v] = c; G.cover[v] = 0;
    G.in_head[v] = G.in_tail[v] = G.out_head[v] = G.out_tail[v] = -1;
    G.al_head[v] = G.al_tail[v] = -1;
    return v;
}

POA_HD void poa_add_edge(PoaGraph &G, int a, int b, int w)
{
    for (int e = G.out_head[a]; e >= 0; e = G.e_next_out[e])
        if (G.e_to[e] == b) { G.e_w[e] += w; return; }
    if (G.E >= G.Ecap) { G.err = 2; return; }
    const int e = G.E++;
    G.e_from[e] = a; G.e_to[e] = b; G.e_w[e] = w; G.e_next_in[e] = G.e_next_out[e] = -1;
    if (G.out_tail[a] < 0) G.out_head[a] = e; else G.e_next_out[G.out_tail[a]] = e;
    G.out_tail[a] = e;
    if (G.in_tail[b] < 0) G.in_head[b] = e; else G.e_next_in[G.in_tail[b]] = e;
    G.in_tail[b] = e;
}

POA_HD void poa_add_aligned(PoaGraph &G, int v, int other)
{
    if (G.A >= G.Acap) { G.err = 3; return; }
    const int a = G.A++;
    G.al_node[a] = other; G.al_next[a] = -1;
    if (G.al_tail[v] < 0) G.al_head[v] = a; else G.al_next[G.al_tail[v]] = a;
    G.al_tail[v] = a;
}

// per-base weight of a layer: quality - 33, or 0 for layers without qualities (window backbone)
POA_HD int poa_weight(const uint8_t *qual, int t) { return qual ? (int)qual[t] - 33 : 0; }

POA_HD int poa_add_chain(PoaGraph &G, const uint8_t *s, const uint8_t *q, int b, int e)
{
    if (b >= e) return -1;
    const int first = poa_add_node(G, s[b]);
    G.cover[first]++;
    int prev = first;
    for (int i = b + 1; i < e; ++i) {
        const int v = poa_add_node(G, s[i]);
        G.cover[v]++;
        poa_add_edge(G, prev, v, poa_weight(q, i - 1) + poa_weight(q, i));
        prev = v;
    }
    return first;
}

// depth-first topological order that keeps the members of an aligned group adjacent
POA_HD void poa_topo_sort(PoaGraph &G)
{
    const int n = G.V;
    int n_order = 0, sp = 0;
    for (int i = 0; i < n; ++i) { G.mark[i] = 0; G.check[i] = 1; }
    for (int i = 0; i < n; ++i) {
        if (G.mark[i]) continue;
        G.stack[sp++] = i;
        while (sp > 0) {
            const int v = G.stack[sp - 1];
            bool ok = true;
            const int mv = G.mark[v];
            // mark 1 = v has pushed its unfinished predecessors / aligned nodes before and is on top
            // again: everything above it has been popped, i.e. finished (the graph is acyclic, aligned
            // nodes included), so the second scan of spoa's loop would find nothing -- skip it
            if (mv == 0) {
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                    const int u = G.e_from[e];
                    if (G.mark[u] != 2) {
                        if (sp >= G.Scap) { G.err = 4; return; }
                        G.stack[sp++] = u; ok = false;
                    }
                }
                if (G.check[v]) {
                    for (int a = G.al_head[v]; a >= 0; a = G.al_next[a]) {
                        const int u = G.al_node[a];
                        if (G.mark[u] != 2) {
                            if (sp >= G.Scap) { G.err = 4; return; }
                            G.stack[sp++] = u; G.check[u] = 0; ok = false;
                        }
                    }
                }
            }
            if (mv != 2) {
                if (ok) {
                    G.mark[v] = 2;
                    if (G.check[v]) {
                        G.order[n_order++] = v;
                        for (int a = G.al_head[v]; a >= 0; a = G.al_next[a]) G.order[n_order++] = G.al_node[a];
                    }
                } else G.mark[v] = 1;
            }
            if (ok) --sp;
        }
    }
    for (int r = 0; r < n; ++r) G.rank[G.order[r]] = r;
}

// The part of the graph a racon window aligns a layer to when the layer does not span the window (racon
// window.cpp generate_consensus -> spoa Graph::subgraph(begin, end)): the nodes reached from backbone node
// `end` by walking in-edges and aligned nodes while the node id stays >= `begin` (the backbone is the first
// sequence of the graph, so a backbone node's id is its window position), re-sorted by the depth-first
// rule of poa_topo_sort restricted to them. member[v] = 1 for the nodes of the view, order[0..n) their
// topological order, rank[v] = position in it (-1 outside). Returns n, or -1 on stack overflow.
POA_HD int poa_subgraph_view(PoaGraph &G, int begin, int end, uint8_t *member, int32_t *order, int32_t *rank)
{
    const int n = G.V;
    int sp = 0, n_order = 0;
    for (int i = 0; i < n; ++i) { member[i] = 0; rank[i] = -1; G.mark[i] = 0; G.check[i] = 1; }
    G.stack[sp++] = end;
    while (sp > 0) {
        const int v = G.stack[--sp];
        if (member[v] || v < begin) continue;
        for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) { if (sp >= G.Scap) { G.err = 4; return -1; } G.stack[sp++] = G.e_from[e]; }
        for (int a = G.al_head[v]; a >= 0; a = G.al_next[a]) { if (sp >= G.Scap) { G.err = 4; return -1; } G.stack[sp++] = G.al_node[a]; }
        member[v] = 1;
    }
    for (int i = 0; i < n; ++i) {
        if (!member[i] || G.mark[i]) continue;
        G.stack[sp++] = i;
        while (sp > 0) {
            const int v = G.stack[sp - 1];
            bool ok = true;
            const int mv = G.mark[v];
            if (mv == 0) {
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                    const int u = G.e_from[e];
                    if (member[u] && G.mark[u] != 2) { if (sp >= G.Scap) { G.err = 4; return -1; } G.stack[sp++] = u; ok = false; }
                }
                if (G.check[v]) {
                    for (int a = G.al_head[v]; a >= 0; a = G.al_next[a]) {
                        const int u = G.al_node[a];
                        if (member[u] && G.mark[u] != 2) { if (sp >= G.Scap) { G.err = 4; return -1; } G.stack[sp++] = u; G.check[u] = 0; ok = false; }
                    }
                }
            }
            if (mv != 2) {
                if (ok) {
                    G.mark[v] = 2;
                    if (G.check[v]) {
                        order[n_order++] = v;
                        for (int a = G.al_head[v]; a >= 0; a = G.al_next[a]) if (member[G.al_node[a]]) order[n_order++] = G.al_node[a];
                    }
                } else G.mark[v] = 1;
            }
            if (ok) --sp;
        }
    }
    for (int r = 0; r < n_order; ++r) rank[order[r]] = r;
    return n_order;
}

// Traceback over the DP matrix H ((V+1) rows of `ld` ints; row r+1 = node order[r], column j =
// j sequence bases consumed). Fills aln_node/aln_pos in reverse order; returns the number of pairs.
POA_HD int poa_traceback(PoaGraph &G, const int32_t *H, size_t ld, const uint8_t *s, int mode,
                         int m, int x, int g, int bi, int bj)
{
    int i = bi, j = bj, n = 0;
#define POA_AT(r, c) POA_LDH(&H[(size_t)(r) * ld + (size_t)(c)])
    while ((mode == 0) ? (POA_AT(i, j) != 0) : (i != 0 || j != 0)) {
        const int h = POA_AT(i, j);
        bool done = false;
        if (i != 0 && j != 0) {
            const int v = G.order[i - 1];
            const int sc = (G.letter[v] == s[j - 1]) ? m : x;
            if (G.in_head[v] < 0) {
                if (h == POA_AT(0, j - 1) + sc) { G.aln_node[n] = v; G.aln_pos[n++] = j - 1; i = 0; --j; done = true; }
            } else {
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                    const int pr = G.rank[G.e_from[e]] + 1;
                    if (h == POA_AT(pr, j - 1) + sc) { G.aln_node[n] = v; G.aln_pos[n++] = j - 1; i = pr; --j; done = true; break; }
                }
            }
        }
        if (!done && i != 0) {
            const int v = G.order[i - 1];
            if (G.in_head[v] < 0) {
                if (h == POA_AT(0, j) + g) { G.aln_node[n] = v; G.aln_pos[n++] = -1; i = 0; done = true; }
            } else {
                for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                    const int pr = G.rank[G.e_from[e]] + 1;
                    if (h == POA_AT(pr, j) + g) { G.aln_node[n] = v; G.aln_pos[n++] = -1; i = pr; done = true; break; }
                }
            }
        }
        if (!done && j != 0) {
            if (h == POA_AT(i, j - 1) + g) { G.aln_node[n] = -1; G.aln_pos[n++] = j - 1; --j; done = true; }
        }
        if (!done) break;
    }
#undef POA_AT
    return n;
}

// Adds a sequence along its alignment (n_aln pairs stored in reverse order in G.aln_*), then
// re-sorts (spoa re-sorts after every sequence).
POA_HD void poa_add_alignment(PoaGraph &G, int n_aln, const uint8_t *s, const uint8_t *q, int L)
{
    if (L == 0) return;
    int first_pos = -1, last_pos = -1;
    for (int t = n_aln - 1; t >= 0; --t)
        if (G.aln_pos[t] >= 0) { if (first_pos < 0) first_pos = G.aln_pos[t]; last_pos = G.aln_pos[t]; }
    if (first_pos < 0) {
        poa_add_chain(G, s, q, 0, L);
        G.n_seqs++;
        poa_topo_sort(G);
        return;
    }
    const int before = G.V;
    poa_add_chain(G, s, q, 0, first_pos);
    int head = (G.V == before) ? -1 : G.V - 1;
    const int tail = poa_add_chain(G, s, q, last_pos + 1, L);
    int prev_w = head == -1 ? 0 : poa_weight(q, first_pos - 1);
    for (int t = n_aln - 1; t >= 0; --t) {
        const int pos = G.aln_pos[t];
        if (pos < 0) continue;
        const uint8_t c = s[pos];
        const int an = G.aln_node[t];
        int node;
        if (an < 0) {
            node = poa_add_node(G, c);
        } else if (G.letter[an] == c) {
            node = an;
        } else {
            node = -1;
            for (int a = G.al_head[an]; a >= 0; a = G.al_next[a])
                if (G.letter[G.al_node[a]] == c) { node = G.al_node[a]; break; }
            if (node < 0) {
                node = poa_add_node(G, c);
                for (int a = G.al_head[an]; a >= 0; a = G.al_next[a]) {
                    const int o = G.al_node[a];
                    poa_add_aligned(G, node, o);
                    poa_add_aligned(G, o, node);
                }
                poa_add_aligned(G, node, an);
                poa_add_aligned(G, an, node);
            }
        }
        G.cover[node]++;
        if (head != -1) poa_add_edge(G, head, node, prev_w + poa_weight(q, pos));
        head = node;
        prev_w = poa_weight(q, pos);
    }
    if (tail != -1) poa_add_edge(G, head, tail, prev_w + poa_weight(q, last_pos + 1));
    G.n_seqs++;
    poa_topo_sort(G);
}

// Contents of `src` into the (larger) arrays of `dst`, which has been bound to its own memory.
inline void poa_graph_copy(PoaGraph &dst, const PoaGraph &src)
{
    const size_t V = (size_t)src.V, E = (size_t)src.E, A = (size_t)src.A;
    dst.V = src.V; dst.E = src.E; dst.A = src.A; dst.n_seqs = src.n_seqs; dst.err = src.err;
#define POA_CP(field, count) memcpy(dst.field, src.field, (count) * sizeof(*src.field));
    POA_CP(letter, V) POA_CP(cover, V) POA_CP(in_head, V) POA_CP(in_tail, V) POA_CP(out_head, V) POA_CP(out_tail, V)
    POA_CP(al_head, V) POA_CP(al_tail, V) POA_CP(order, V) POA_CP(rank, V)
    POA_CP(e_from, E) POA_CP(e_to, E) POA_CP(e_w, E) POA_CP(e_next_in, E) POA_CP(e_next_out, E)
    POA_CP(al_node, A) POA_CP(al_next, A)
#undef POA_CP
}

POA_HD void poa_relax(PoaGraph &G, int v, bool skip_dead)
{
    for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
        const int u = G.e_from[e];
        const int64_t w = G.e_w[e];
        if (skip_dead && G.score[u] == -1) continue;
        if (G.score[v] < w || (G.score[v] == w && G.pred[v] != -1 && G.score[G.pred[v]] <= G.score[u])) {
            G.score[v] = w; G.pred[v] = u;
        }
    }
    if (G.pred[v] != -1) G.score[v] += G.score[G.pred[v]];
}

// Heaviest bundle; writes the consensus (optionally coverage-trimmed) to out, returns its length
// (or -1 when it does not fit). Uses aln_node as scratch for the path.
POA_HD int poa_consensus(PoaGraph &G, int trim, uint8_t *out, int cap)
{
    const int V = G.V;
    if (V == 0) return 0;
    for (int v = 0; v < V; ++v) { G.score[v] = -1; G.pred[v] = -1; }
    int best = G.order[0];
    for (int r = 0; r < V; ++r) {
        const int v = G.order[r];
        poa_relax(G, v, false);
        if (G.score[best] < G.score[v]) best = v;
    }
    while (G.out_head[best] >= 0) {
        const int r0 = G.rank[best];
        for (int e = G.out_head[best]; e >= 0; e = G.e_next_out[e])
            for (int e2 = G.in_head[G.e_to[e]]; e2 >= 0; e2 = G.e_next_in[e2])
                if (G.e_from[e2] != best) G.score[G.e_from[e2]] = -1;
        int64_t ms = 0; int mid = -1;
        for (int r = r0 + 1; r < V; ++r) {
            const int v = G.order[r];
            G.score[v] = -1; G.pred[v] = -1;
            poa_relax(G, v, true);
            if (ms < G.score[v]) { ms = G.score[v]; mid = v; }
        }
        if (mid < 0) break;
        best = mid;
    }
    int n = 0;
    while (best != -1) { G.aln_node[n++] = best; best = G.pred[best]; }   // reverse order
    int b = 0, e = n;                      // indices into the forward path: forward i = aln_node[n-1-i]
    if (trim) {
        const int need = (G.n_seqs - 1) / 2;
        while (b < e && G.cover[G.aln_node[n - 1 - b]] < need) ++b;
        while (e > b && G.cover[G.aln_node[n - e]] < need) --e;
        if (b >= e) { b = 0; e = n; }
    }
    if (e - b > cap) return -1;
    for (int i = b; i < e; ++i) out[i - b] = G.letter[G.aln_node[n - 1 - i]];
    return e - b;
}
