/*
 * oracle/poa_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the partial-order-alignment consensus the reference obtains by shelling out
 * to the third-party binaries spoa (4.0.7, Dockerfile:14) and racon (1.4.20, Dockerfile:15), none
 * of which is vendored under /root/reference or installable here.
 *
 * Reference call sites:
 *   modules/consensus.py:83-92    run_spoa:  `spoa <fastq> -l 0 -r 0 -g -2`  -> line 2 of stdout
 *   modules/consensus.py:107-126  run_racon: racon_iter x (minimap2 -x map-ont ; racon) (see
 *                                 oracle/consensus_oracle.py for the windowing restatement)
 *
 * PARITY UNPINNED: no spoa/racon output exists anywhere in the reference (it has no tests), so
 * this file follows spoa's published algorithm (Lee 2002 POA; Vaser 2017 racon/spoa):
 *   - graph of nodes (one letter each), weighted directed edges, groups of mutually "aligned"
 *     nodes (alternative letters of one column);
 *   - a sequence is aligned to the graph by DP over the topologically ordered nodes
 *     (rows) x sequence positions (columns); spoa's engine picks LINEAR gaps when g >= e, so
 *     `-g -2` (e defaults to -6) means gap = -2 per base; `-l 0` = local (Smith-Waterman),
 *     m = +5, n = -4; racon's windows use global alignment with m = 3, n = -5, g = -4;
 *   - FASTQ qualities are per-base weights (q - 33); an edge accumulates the sum of the two
 *     base weights of every sequence that traverses it;
 *   - consensus = heaviest bundle: per node the heaviest in-edge (ties -> predecessor with the
 *     higher score), best-scoring node, completed forward to a sink, then backtracked.
 * Traceback preference: diagonal through the predecessors in edge-insertion order, then
 * "graph node against a gap", then "sequence base against a gap".
 */
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Edge { int from, to; long long w; };

struct Graph {
    std::vector<char> letter;
    std::vector<std::vector<int>> in, out;      // edge ids, insertion order
    std::vector<std::vector<int>> aligned;      // node ids
    std::vector<Edge> edges;
    std::vector<int> order;                     // topological order (rank -> node)
    std::vector<int> rank;                      // node -> rank
    std::vector<int> coverage;                  // sequences through the node
    int n_seqs = 0;
    int order_mode = 0;                         // 0: spoa's DFS sort after every sequence, 1: path insertion

    int add_node(char c) {
        letter.push_back(c); in.emplace_back(); out.emplace_back(); aligned.emplace_back();
        coverage.push_back(0);
        return (int)letter.size() - 1;
    }
    void add_edge(int a, int b, long long w) {
        for (int e : out[a]) if (edges[e].to == b) { edges[e].w += w; return; }
        edges.push_back({a, b, w});
        out[a].push_back((int)edges.size() - 1);
        in[b].push_back((int)edges.size() - 1);
    }
    // chain of new nodes for seq[b, e); returns first node or -1
    int add_chain(const char *s, const int *wt, int b, int e) {
        if (b >= e) return -1;
        int first = add_node(s[b]);
        coverage[first]++;
        int prev = first;
        for (int i = b + 1; i < e; ++i) {
            int v = add_node(s[i]);
            coverage[v]++;
            add_edge(prev, v, (long long)wt[i - 1] + wt[i]);
            prev = v;
        }
        return first;
    }
    // depth-first topological sort keeping aligned nodes adjacent
    void topo_sort() {
        const int n = (int)letter.size();
        order.clear();
        std::vector<uint8_t> mark(n, 0), check(n, 1);
        std::vector<int> st;
        for (int i = 0; i < n; ++i) {
            if (mark[i]) continue;
            st.push_back(i);
            while (!st.empty()) {
                int v = st.back();
                bool ok = true;
                if (mark[v] != 2) {
                    for (int e : in[v]) if (mark[edges[e].from] != 2) { st.push_back(edges[e].from); ok = false; }
                    if (check[v]) for (int a : aligned[v]) if (mark[a] != 2) { st.push_back(a); check[a] = 0; ok = false; }
                    if (ok) {
                        mark[v] = 2;
                        if (check[v]) { order.push_back(v); for (int a : aligned[v]) order.push_back(a); }
                    } else mark[v] = 1;
                }
                if (ok) st.pop_back();
            }
        }
        rank.assign(n, 0);
        for (int r = 0; r < n; ++r) rank[order[r]] = r;
    }
};

struct Pair { int node, pos; };

// The part of the graph a racon window aligns a layer to when the layer does not span the window
// (racon window.cpp generate_consensus -> spoa Graph::subgraph(begin, end) of the vendored spoa 3.x): the
// nodes reached from backbone node `end` by walking in-edges and aligned nodes while the node id stays
// >= `begin` (backbone nodes carry their window position as id: the backbone is the first sequence of
// the graph), with the edges and aligned links between them, re-sorted by the same depth-first rule.
struct View {
    std::vector<int> order, rank;            // rank of a non-member: -1
    std::vector<uint8_t> member;
};

static View subgraph_view(const Graph &G, int begin, int end)
{
    const int n = (int)G.letter.size();
    View W;
    W.member.assign(n, 0);
    std::vector<int> st;
    st.push_back(end);
    while (!st.empty()) {
        const int v = st.back(); st.pop_back();
        if (W.member[v] || v < begin) continue;
        for (int e : G.in[v]) st.push_back(G.edges[e].from);
        for (int a : G.aligned[v]) st.push_back(a);
        W.member[v] = 1;
    }
    // Graph::topo_sort restricted to the members (ids ascending = the sub-graph's own node order)
    std::vector<uint8_t> mark(n, 0), check(n, 1);
    for (int i = 0; i < n; ++i) {
        if (!W.member[i] || mark[i]) continue;
        st.push_back(i);
        while (!st.empty()) {
            const int v = st.back();
            bool ok = true;
            if (mark[v] != 2) {
                for (int e : G.in[v]) { const int u = G.edges[e].from; if (W.member[u] && mark[u] != 2) { st.push_back(u); ok = false; } }
                if (check[v]) for (int a : G.aligned[v]) if (W.member[a] && mark[a] != 2) { st.push_back(a); check[a] = 0; ok = false; }
                if (ok) {
                    mark[v] = 2;
                    if (check[v]) { W.order.push_back(v); for (int a : G.aligned[v]) if (W.member[a]) W.order.push_back(a); }
                } else mark[v] = 1;
            }
            if (ok) st.pop_back();
        }
    }
    W.rank.assign(n, -1);
    for (size_t r = 0; r < W.order.size(); ++r) W.rank[W.order[r]] = (int)r;
    return W;
}

static View whole_view(const Graph &G)
{
    View W;
    W.order = G.order; W.rank = G.rank;
    W.member.assign(G.letter.size(), 1);
    return W;
}

// mode 0 = local, 1 = global (both with linear gap g); W: the nodes taking part, in topological order
static std::vector<Pair> align(const Graph &G0, const View &W, const char *s, int L, int mode, int m, int x, int g)
{
    // the DP below reads the graph through `G`: in / out lists restricted to the members of the view
    struct Sub {
        const Graph &g; const View &w;
        std::vector<std::vector<int>> in;        // member in-edges (edge ids, insertion order) per node
        std::vector<uint8_t> has_out;
        const std::vector<char> &letter; const std::vector<Edge> &edges; const std::vector<int> &order, &rank;
        Sub(const Graph &g_, const View &w_) : g(g_), w(w_), letter(g_.letter), edges(g_.edges), order(w_.order), rank(w_.rank) {
            in.resize(g.letter.size()); has_out.assign(g.letter.size(), 0);
            for (int v : w.order) {
                for (int e : g.in[v]) if (w.member[g.edges[e].from]) in[v].push_back(e);
                for (int e : g.out[v]) if (w.member[g.edges[e].to]) has_out[v] = 1;
            }
        }
    } G(G0, W);
    const int V = (int)W.order.size();
    std::vector<Pair> aln;
    if (V == 0 || L == 0) return aln;
    const int NEG = INT_MIN / 4;
    std::vector<int> H((size_t)(V + 1) * (L + 1), 0);
    auto at = [&](int r, int j) -> int & { return H[(size_t)r * (L + 1) + j]; };
    if (mode == 1) {
        for (int j = 1; j <= L; ++j) at(0, j) = j * g;
    }
    int best = mode == 0 ? 0 : NEG, bi = 0, bj = 0;
    for (int r = 0; r < V; ++r) {
        const int v = G.order[r];
        const char c = G.letter[v];
        const int row = r + 1;
        if (mode == 1) {
            int p = NEG;
            if (G.in[v].empty()) p = 0;
            for (int e : G.in[v]) p = std::max(p, at(G.rank[G.edges[e].from] + 1, 0));
            at(row, 0) = p + g;
        }
        for (int j = 1; j <= L; ++j) {
            const int sc = (c == s[j - 1]) ? m : x;
            int h = NEG;
            if (G.in[v].empty()) {
                h = std::max(at(0, j - 1) + sc, at(0, j) + g);
            } else {
                for (int e : G.in[v]) {
                    const int pr = G.rank[G.edges[e].from] + 1;
                    h = std::max(h, std::max(at(pr, j - 1) + sc, at(pr, j) + g));
                }
            }
            h = std::max(h, at(row, j - 1) + g);
            if (mode == 0) {
                if (h < 0) h = 0;
                if (h > best) { best = h; bi = row; bj = j; }
            }
            at(row, j) = h;
        }
        if (mode == 1 && !G.has_out[v]) {
            if (at(row, L) > best) { best = at(row, L); bi = row; bj = L; }
        }
    }
    if (mode == 0 && best == 0) return aln;
    // traceback
    int i = bi, j = bj;
    std::vector<Pair> rev;
    while ((mode == 0) ? (at(i, j) != 0) : (i != 0 || j != 0)) {
        const int h = at(i, j);
        bool done = false;
        if (i != 0 && j != 0) {
            const int v = G.order[i - 1];
            const int sc = (G.letter[v] == s[j - 1]) ? m : x;
            if (G.in[v].empty()) {
                if (h == at(0, j - 1) + sc) { rev.push_back({v, j - 1}); i = 0; --j; done = true; }
            } else {
                for (int e : G.in[v]) {
                    const int pr = G.rank[G.edges[e].from] + 1;
                    if (h == at(pr, j - 1) + sc) { rev.push_back({v, j - 1}); i = pr; --j; done = true; break; }
                }
            }
        }
        if (!done && i != 0) {
            const int v = G.order[i - 1];
            if (G.in[v].empty()) {
                if (h == at(0, j) + g) { rev.push_back({v, -1}); i = 0; done = true; }
            } else {
                for (int e : G.in[v]) {
                    const int pr = G.rank[G.edges[e].from] + 1;
                    if (h == at(pr, j) + g) { rev.push_back({v, -1}); i = pr; done = true; break; }
                }
            }
        }
        if (!done && j != 0) {
            if (h == at(i, j - 1) + g) { rev.push_back({-1, j - 1}); --j; done = true; }
        }
        if (!done) break;   // cannot happen for a consistent matrix
    }
    aln.assign(rev.rbegin(), rev.rend());
    return aln;
}

// Order maintenance without a full sort ("path insertion"): a run of new nodes between the old
// path nodes p and s goes immediately after p (in path order); a run at the start of the path
// goes immediately before s; a path without old nodes is appended. The alignment is monotone in
// the current order, so the result is again a topological order.
static void insert_path_order(Graph &G, const std::vector<int> &path, int old_V)
{
    std::vector<int> run;
    int last_old = -1;
    auto place_after = [&](int p, const std::vector<int> &r) {
        auto it = std::find(G.order.begin(), G.order.end(), p);
        G.order.insert(it + 1, r.begin(), r.end());
    };
    for (size_t i = 0; i < path.size(); ++i) {
        const int v = path[i];
        if (v >= old_V) { run.push_back(v); continue; }
        if (!run.empty()) {
            if (last_old >= 0) place_after(last_old, run);
            else { auto it = std::find(G.order.begin(), G.order.end(), v); G.order.insert(it, run.begin(), run.end()); }
            run.clear();
        }
        last_old = v;
    }
    if (!run.empty()) {
        if (last_old >= 0) place_after(last_old, run);
        else G.order.insert(G.order.end(), run.begin(), run.end());
    }
    G.rank.assign(G.letter.size(), 0);
    for (size_t r = 0; r < G.order.size(); ++r) G.rank[G.order[r]] = (int)r;
}

static void add_alignment(Graph &G, const std::vector<Pair> &aln, const char *s, const int *wt, int L)
{
    if (L == 0) return;
    const int old_V = (int)G.letter.size();
    std::vector<int> path;
    std::vector<int> valid;
    for (const Pair &p : aln) if (p.pos >= 0) valid.push_back(p.pos);
    if (valid.empty()) {
        G.add_chain(s, wt, 0, L);
        G.n_seqs++;
        if (G.order_mode == 0) G.topo_sort();
        else { for (int v = old_V; v < (int)G.letter.size(); ++v) path.push_back(v); insert_path_order(G, path, old_V); }
        return;
    }
    const int before = (int)G.letter.size();
    G.add_chain(s, wt, 0, valid.front());
    int head = ((int)G.letter.size() == before) ? -1 : (int)G.letter.size() - 1;
    for (int v = before; v < (int)G.letter.size(); ++v) path.push_back(v);
    const int tail_first_id = (int)G.letter.size();
    int tail = G.add_chain(s, wt, valid.back() + 1, L);
    const int tail_end_id = (int)G.letter.size();
    long long prev_w = head == -1 ? 0 : wt[valid.front() - 1];
    for (const Pair &p : aln) {
        if (p.pos < 0) continue;
        const char c = s[p.pos];
        int node;
        if (p.node < 0) {
            node = G.add_node(c);
        } else if (G.letter[p.node] == c) {
            node = p.node;
        } else {
            node = -1;
            for (int a : G.aligned[p.node]) if (G.letter[a] == c) { node = a; break; }
            if (node < 0) {
                node = G.add_node(c);
                for (int a : G.aligned[p.node]) { G.aligned[node].push_back(a); G.aligned[a].push_back(node); }
                G.aligned[node].push_back(p.node);
                G.aligned[p.node].push_back(node);
            }
        }
        G.coverage[node]++;
        path.push_back(node);
        if (head != -1) G.add_edge(head, node, prev_w + wt[p.pos]);
        head = node;
        prev_w = wt[p.pos];
    }
    if (tail != -1) G.add_edge(head, tail, prev_w + wt[valid.back() + 1]);
    for (int v = tail_first_id; v < tail_end_id; ++v) path.push_back(v);
    G.n_seqs++;
    if (G.order_mode == 0) G.topo_sort();
    else insert_path_order(G, path, old_V);
}

static std::vector<int> heaviest_bundle(const Graph &G)
{
    const int V = (int)G.letter.size();
    std::vector<int> path;
    if (V == 0) return path;
    std::vector<long long> score(V, -1);
    std::vector<int> pred(V, -1);
    auto relax = [&](int v, bool skip_dead) {
        for (int e : G.in[v]) {
            const Edge &E = G.edges[e];
            if (skip_dead && score[E.from] == -1) continue;
            if (score[v] < E.w || (score[v] == E.w && pred[v] != -1 && score[pred[v]] <= score[E.from])) {
                score[v] = E.w; pred[v] = E.from;
            }
        }
        if (pred[v] != -1) score[v] += score[pred[v]];
    };
    int best = G.order[0];
    for (int r = 0; r < V; ++r) {
        int v = G.order[r];
        relax(v, false);
        if (score[best] < score[v]) best = v;
    }
    // complete the branch to a sink
    while (!G.out[best].empty()) {
        const int r0 = G.rank[best];
        for (int e : G.out[best])
            for (int e2 : G.in[G.edges[e].to])
                if (G.edges[e2].from != best) score[G.edges[e2].from] = -1;
        long long ms = 0; int mid = -1;
        for (int r = r0 + 1; r < V; ++r) {
            int v = G.order[r];
            score[v] = -1; pred[v] = -1;
            relax(v, true);
            if (ms < score[v]) { ms = score[v]; mid = v; }
        }
        if (mid < 0) break;
        best = mid;
    }
    while (best != -1) { path.push_back(best); best = pred[best]; }
    std::reverse(path.begin(), path.end());
    return path;
}

}  // namespace

extern "C" {

/*
 * POA consensus of n sequences added in the given order. quals may be NULL (all weights 1) or
 * hold PHRED+33 strings (weight = q - 33; an empty string means weight 0 for that sequence, which
 * is how a racon window treats its backbone). mode 0 local / 1 global; linear gap g (< 0).
 * trim != 0 applies racon's coverage trimming to the consensus ends (coverage >= (n-1)/2).
 * Returns the consensus length (written to out, capacity cap) or -1.
 */
int oracle_poa_consensus_ex(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                            int trim, int order_mode, char *out, int cap, int *n_nodes_out);

int oracle_poa_consensus(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                         int trim, char *out, int cap, int *n_nodes_out)
{
    return oracle_poa_consensus_ex(seqs, quals, n, mode, m, x, g, trim, 0, out, cap, n_nodes_out);
}

/* order_mode 0: spoa's topological re-sort after every sequence (reference-faithful);
 * order_mode 1: "path insertion" order maintenance (what the CUDA kernel does; any topological
 * order gives a valid POA, the two differ only in tie-breaks between equal scores). */
int oracle_poa_consensus_sub(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                             int trim, int order_mode, const int *sub_b, const int *sub_e, char *out, int cap, int *n_nodes_out);

int oracle_poa_consensus_ex(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                            int trim, int order_mode, char *out, int cap, int *n_nodes_out)
{
    return oracle_poa_consensus_sub(seqs, quals, n, mode, m, x, g, trim, order_mode, nullptr, nullptr, out, cap, n_nodes_out);
}

/* sub_b / sub_e (or NULL): per sequence the backbone positions [sub_b, sub_e] of the first sequence the
 * sequence covers; >= 0 aligns it to that sub-graph only (racon's treatment of layers that do not span
 * their window), -1 to the whole graph. */
int oracle_poa_consensus_sub(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                             int trim, int order_mode, const int *sub_b, const int *sub_e, char *out, int cap, int *n_nodes_out)
{
    Graph G;
    G.order_mode = order_mode;
    std::vector<int> wt;
    for (int i = 0; i < n; ++i) {
        const int L = (int)strlen(seqs[i]);
        wt.assign(L, 1);
        if (quals) {
            const int ql = (int)strlen(quals[i]);
            for (int t = 0; t < L; ++t) wt[t] = (t < ql) ? (int)quals[i][t] - 33 : 0;
        }
        std::vector<Pair> aln;
        if (!G.letter.empty()) {
            const bool sub = sub_b && sub_b[i] >= 0 && sub_e[i] >= sub_b[i] && sub_e[i] < (int)G.letter.size();
            aln = align(G, sub ? subgraph_view(G, sub_b[i], sub_e[i]) : whole_view(G), seqs[i], L, mode, m, x, g);
        }
        add_alignment(G, aln, seqs[i], wt.data(), L);
    }
    std::vector<int> path = heaviest_bundle(G);
    int b = 0, e = (int)path.size();
    if (trim) {
        const int need = (G.n_seqs - 1) / 2;
        while (b < e && G.coverage[path[b]] < need) ++b;
        while (e > b && G.coverage[path[e - 1]] < need) --e;
        if (b >= e) { b = 0; e = (int)path.size(); }
    }
    if (e - b + 1 > cap) return -1;
    for (int i = b; i < e; ++i) out[i - b] = G.letter[path[i]];
    out[e - b] = 0;
    if (n_nodes_out) *n_nodes_out = (int)G.letter.size();
    return e - b;
}

/* Test hook: graph of the n sequences (added as in oracle_poa_consensus_ex, order_mode 0), then the node ids of
 * subgraph_view(begin, end) in view order. Returns their number, -2 for a range outside the graph, -3 if cap is small. */
int oracle_poa_subview(const char **seqs, const char **quals, int n, int mode, int m, int x, int g,
                       int begin, int end, int *order_out, int cap)
{
    Graph G;
    std::vector<int> wt;
    for (int i = 0; i < n; ++i) {
        const int L = (int)strlen(seqs[i]);
        wt.assign(L, 1);
        if (quals) {
            const int ql = (int)strlen(quals[i]);
            for (int t = 0; t < L; ++t) wt[t] = (t < ql) ? (int)quals[i][t] - 33 : 0;
        }
        std::vector<Pair> aln;
        if (!G.letter.empty()) aln = align(G, whole_view(G), seqs[i], L, mode, m, x, g);
        add_alignment(G, aln, seqs[i], wt.data(), L);
    }
    if (begin < 0 || end < begin || end >= (int)G.letter.size()) return -2;
    const View W = subgraph_view(G, begin, end);
    if ((int)W.order.size() > cap) return -3;
    for (size_t r = 0; r < W.order.size(); ++r) order_out[r] = W.order[r];
    return (int)W.order.size();
}

}  // extern "C"
