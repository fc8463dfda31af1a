// Host build of the product's POA graph core (ngspeciesid_b200/csrc/poa_core.cuh) with a serial DP,
// so that tests can check the graph logic against oracle/poa_oracle.cpp without a GPU.
// This file is test scaffolding: it is never linked into libngsid.so.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../ngspeciesid_b200/csrc/poa_core.cuh"

static int host_consensus_or_view(const char **seqs, const char **quals, int n, int mode, int m, int x,
                                  int g, int trim, char *out, int cap, int Vcap, int view_begin, int view_end, int *view_order, int view_cap)
{
    int Lmax = 1;
    for (int i = 0; i < n; ++i) Lmax = std::max<int>(Lmax, (int)strlen(seqs[i]));
    const int Ecap = Vcap * 6, Acap = Vcap * 6, Scap = Vcap * 14;
    std::vector<uint8_t> mem(poa_graph_bytes(Vcap, Ecap, Acap, Scap, Lmax));
    PoaGraph G;
    poa_graph_bind(G, mem.data(), Vcap, Ecap, Acap, Scap, Lmax);
    std::vector<int32_t> H;
    for (int si = 0; si < n; ++si) {
        const uint8_t *s = (const uint8_t *)seqs[si];
        const int L = (int)strlen(seqs[si]);
        const uint8_t *q = (quals && quals[si] && quals[si][0]) ? (const uint8_t *)quals[si] : nullptr;
        int n_aln = 0;
        if (G.V > 0 && L > 0) {
            const size_t ld = (size_t)L + 1;
            H.assign((size_t)(G.V + 1) * ld, 0);
            if (mode == 1) for (int j = 1; j <= L; ++j) H[j] = j * g;
            int best = mode == 0 ? 0 : POA_NEG, bi = 0, bj = 0;
            for (int r = 0; r < G.V; ++r) {
                const int v = G.order[r];
                const int row = r + 1;
                if (mode == 1) {
                    int p = POA_NEG;
                    if (G.in_head[v] < 0) p = 0;
                    for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) p = std::max(p, H[(size_t)(G.rank[G.e_from[e]] + 1) * ld]);
                    H[(size_t)row * ld] = p + g;
                }
                for (int j = 1; j <= L; ++j) {
                    const int sc = (G.letter[v] == s[j - 1]) ? m : x;
                    int h = POA_NEG;
                    if (G.in_head[v] < 0) h = std::max(H[j - 1] + sc, H[j] + g);
                    for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
                        const size_t pr = (size_t)(G.rank[G.e_from[e]] + 1) * ld;
                        h = std::max(h, std::max(H[pr + j - 1] + sc, H[pr + j] + g));
                    }
                    h = std::max(h, H[(size_t)row * ld + j - 1] + g);
                    if (mode == 0) { if (h < 0) h = 0; if (h > best) { best = h; bi = row; bj = j; } }
                    H[(size_t)row * ld + j] = h;
                }
                if (mode == 1 && G.out_head[v] < 0 && H[(size_t)row * ld + L] > best) { best = H[(size_t)row * ld + L]; bi = row; bj = L; }
            }
            if (!(mode == 0 && best == 0)) n_aln = poa_traceback(G, H.data(), ld, s, mode, m, x, g, bi, bj);
        }
        poa_add_alignment(G, n_aln, s, q, L);
        if (G.err) return -100 - G.err;
    }
    if (view_order) {
        // the sub-graph view of the finished graph (poa_subgraph_view): node ids in view order
        if (view_begin < 0 || view_end < view_begin || view_end >= G.V) return -2;
        std::vector<uint8_t> member((size_t)G.V);
        std::vector<int32_t> order((size_t)G.V), rank((size_t)G.V);
        const int nv = poa_subgraph_view(G, view_begin, view_end, member.data(), order.data(), rank.data());
        if (nv < 0 || nv > view_cap) return -3;
        for (int r = 0; r < nv; ++r) { view_order[r] = order[r]; if (rank[order[r]] != r || !member[order[r]]) return -4; }
        return nv;
    }
    int len = poa_consensus(G, trim, (uint8_t *)out, cap - 1);
    if (len >= 0) out[len] = 0;
    return len;
}

extern "C" int poa_core_host_consensus(const char **seqs, const char **quals, int n, int mode, int m, int x,
                                       int g, int trim, char *out, int cap, int Vcap)
{
    return host_consensus_or_view(seqs, quals, n, mode, m, x, g, trim, out, cap, Vcap, -1, -1, nullptr, 0);
}

// graph of the n sequences (added like poa_core_host_consensus), then the view [begin, end] of it
extern "C" int poa_core_host_subview(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                                     int begin, int end, int *order_out, int order_cap, int Vcap)
{
    return host_consensus_or_view(seqs, quals, n, mode, m, x, g, 0, nullptr, 0, Vcap, begin, end, order_out, order_cap);
}
