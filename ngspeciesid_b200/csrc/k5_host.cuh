// Host side of K5 (see k5_rows.cuh): graphs of all jobs on the host, one kernel launch per layer step
// over every job that still has a layer, graph update + spoa's re-sort on host threads in between.
#pragma once
#include <omp.h>
#include <chrono>
#include <thread>
#include "k5_rows.cuh"

namespace k5host {

// ---- pinned host staging that only grows
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostGraph {
    std::vector<uint8_t> mem;
    PoaGraph G;
    int Lmax = 1;
    bool init = false;
    // the rows of the current layer step: the whole graph in its topological order, or the sub-graph a layer
    // that does not span its window is aligned to (poa_subgraph_view)
    bool sub = false;
    int Vs = 0;
    std::vector<int32_t> v_order, v_rank;
    std::vector<uint8_t> v_member;
    const int32_t *row_order() const { return sub ? v_order.data() : G.order; }
    const int32_t *row_rank() const { return sub ? v_rank.data() : G.rank; }
    int rows() const { return sub ? Vs : G.V; }
    void alloc(int Vcap, int lmax) {
        const int Ecap = Vcap * 4, Acap = Vcap * 8, Scap = Ecap + Acap + Vcap + 64;
        Lmax = lmax;
        mem.assign(poa_graph_bytes(Vcap, Ecap, Acap, Scap, lmax) + 64, 0);
        poa_graph_bind(G, mem.data(), Vcap, Ecap, Acap, Scap, lmax);
        init = true;
    }
    // room for one more layer of L bases (every base may add a node, two edges, six aligned entries)
    void reserve(int L) {
        if (!init) { alloc(std::max(2048, 4 * L + 64), L); return; }
        if (G.V + L + 2 <= G.Vcap && G.E + 2 * L + 4 <= G.Ecap && G.A + 8 * L + 8 <= G.Acap && L <= Lmax) return;
        int Vcap = G.Vcap;
        while (G.V + L + 2 > Vcap || G.E + 2 * L + 4 > Vcap * 4 || G.A + 8 * L + 8 > Vcap * 8) Vcap *= 2;
        HostGraph n;
        n.alloc(Vcap, std::max(L, Lmax));
        poa_graph_copy(n.G, G);
        mem.swap(n.mem);
        G = n.G;
        Lmax = n.Lmax;
    }
};

struct Layer { const uint8_t *s; const uint8_t *q; int L; int64_t arena_off; };

// rows of the graph (or of the current sub-graph view) in topological order for the kernel; returns the number
// of overflow entries. In a view, edges from / to nodes outside it do not exist.
static int64_t build_rows(const HostGraph &hg, int ring, uint4 *rows, uint2 *plans, int32_t *ovf, bool count_only, int *max_np)
{
    const PoaGraph &G = hg.G;
    const int Vr = hg.rows();
    const int32_t *order = hg.row_order(), *rank = hg.row_rank();
    const uint8_t *member = hg.sub ? hg.v_member.data() : nullptr;
    int64_t n_ovf = 0;
    const bool pack16 = Vr + 1 < 65535;
    const int inl = pack16 ? K5R_INLINE16 : K5R_INLINE32;
    for (int r = 0; r < Vr; ++r) {
        const int v = order[r];
        int np = 0;
        for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) if (!member || member[G.e_from[e]]) ++np;
        if (np > *max_np) *max_np = np;
        if (count_only) { if (np > inl) n_ovf += np; continue; }
        bool sink = true;
        for (int e = G.out_head[v]; e >= 0 && sink; e = G.e_next_out[e]) if (!member || member[G.e_to[e]]) sink = false;
        uint32_t f[4] = {(uint32_t)G.letter[v] | ((uint32_t)std::min(np, 255) << 8) | (sink ? K5R_FLAG_SINK : 0u) |
                         (pack16 ? K5R_FLAG_PACK16 : 0u), 0u, 0u, 0u};
        int u = 0;
        if (np > inl) f[1] = (uint32_t)n_ovf;
        // plan of the row for the pipelined loop: previous row, virtual row, two more near rows
        uint32_t chain_u = 255, virt_u = np == 0 ? 0u : 255u, npre = 0, generic = np > K5R_MAXE ? 1u : 0u, py = 0;
        for (int e = G.in_head[v]; e >= 0; e = G.e_next_in[e]) {
            if (member && !member[G.e_from[e]]) continue;
            const int p = rank[G.e_from[e]] + 1;
            if (np > inl) ovf[n_ovf + u] = p;
            else if (pack16) f[1 + (u >> 1)] |= (uint32_t)p << ((u & 1) * 16);
            else f[1 + u] = (uint32_t)p;
            const int dist = (r + 1) - p;
            if (dist >= ring) rows[p - 1].x |= K5R_FLAG_STORE;               // a far successor reads it from global memory
            if (dist == 1) chain_u = (uint32_t)u;
            else if (dist < ring && npre < 2 && u < 255) { py |= ((uint32_t)dist | ((uint32_t)u << 8)) << (16 * npre); ++npre; }
            else generic = 1;
            ++u;
        }
        if (np > inl) n_ovf += np;
        rows[r] = make_uint4(f[0], f[1], f[2], f[3]);
        plans[r] = make_uint2(chain_u | (virt_u << 8) | (npre << 16) | (generic << 24), py);
    }
    return n_ovf;
}

__global__ void k_gather_layers(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual, const int64_t *__restrict__ off,
                                const uint8_t *__restrict__ aux, const int64_t *__restrict__ aoff,
                                const int32_t *__restrict__ src, const int32_t *__restrict__ beg, const int32_t *__restrict__ len,
                                const int64_t *__restrict__ lay_off, int64_t n_layers, uint8_t *__restrict__ o_seq, uint8_t *__restrict__ o_qual)
{
    const int64_t wi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wi >= n_layers) return;
    const uint32_t lane = lane_id();
    const int sr = src[wi], b = beg[wi], L = len[wi];
    const uint8_t *s = (sr >= 0 ? seq + off[sr] : aux + aoff[-sr - 1]) + b;
    const uint8_t *q = sr >= 0 ? qual + off[sr] + b : nullptr;
    uint8_t *os = o_seq + lay_off[wi], *oq = o_qual + lay_off[wi];
    for (int i = lane; i < L; i += 32) { os[i] = s[i]; oq[i] = q ? q[i] : (uint8_t)33; }
}

}  // namespace k5host

static int host_threads(const ngsid_ctx *ctx)
{
    if (const char *e = getenv("NGSID_HOST_THREADS")) return std::max(1, atoi(e));
    const int hw = (int)std::thread::hardware_concurrency();
    const int ranks = ctx->nccl_comm ? std::max(1, ctx->nccl_nranks) : 1;
    return std::max(1, std::min(16, (hw > 0 ? hw : 4) / ranks));
}

static int poa_consensus_impl(ngsid_ctx *ctx, const ngsid_poa_params *params, int64_t n_jobs,
                              const int64_t *job_off, const int32_t *layer_src, const int32_t *layer_begin,
                              const int32_t *layer_len, const int32_t *layer_sub_begin, const int32_t *layer_sub_end,
                              const uint8_t *aux_seq, const int64_t *aux_off,
                              int64_t n_aux, uint8_t *out_seq, int64_t out_stride, int32_t *out_len,
                              int32_t *out_nodes)
{
    using namespace k5host;
    if (!ctx || !params || n_jobs < 0) return NGSID_EINVAL;
    if (n_jobs == 0) return NGSID_OK;
    if (!job_off || !layer_src || !layer_begin || !layer_len || !out_seq || !out_len || out_stride < 1) return NGSID_EINVAL;
    if (params->gap >= 0) return fail(ctx, NGSID_EINVAL, "gap must be negative");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n_layers = job_off[n_jobs];
    int Lmax = 1;
    std::vector<int64_t> lay_off((size_t)n_layers + 1, 0);
    int64_t max_job_layers = 0;
    for (int64_t j = 0; j < n_jobs; ++j) max_job_layers = std::max(max_job_layers, job_off[j + 1] - job_off[j]);
    for (int64_t l = 0; l < n_layers; ++l) {
        const int src = layer_src[l];
        int64_t len;
        if (src >= 0) { if (src >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "layer read out of range"); len = ctx->h_off[src + 1] - ctx->h_off[src]; }
        else { int64_t x = -(int64_t)src - 1; if (x >= n_aux) return fail(ctx, NGSID_EINVAL, "layer aux out of range"); len = aux_off[x + 1] - aux_off[x]; }
        if (layer_begin[l] < 0 || layer_len[l] < 0 || (int64_t)layer_begin[l] + layer_len[l] > len) return fail(ctx, NGSID_EINVAL, "layer range out of bounds");
        Lmax = std::max(Lmax, (int)layer_len[l]);
        lay_off[l + 1] = lay_off[l] + ((layer_len[l] + 15) & ~15);
    }
    if (Lmax + 1 > K5R_MAXW * K5R_TILE) return fail(ctx, NGSID_EUNSUPPORTED, "POA layer longer than 4095 bases");
    int rc = upload_aux(ctx, aux_seq, aux_off, n_aux);
    if (rc) return rc;

    // ---- bases + qualities of every layer: one arena on the device (read by the DP) and on the host (graph update)
    const size_t arena = (size_t)lay_off[n_layers] + 64;
    static thread_local PinBuf pin_lay, pin_meta, pin_path;
    CUDA_TRY(ctx, ctx->d_poa_arena.ensure(2 * arena));
    CUDA_TRY(ctx, pin_lay.ensure(2 * arena));
    CUDA_TRY(ctx, ctx->d_lsrc.ensure((size_t)n_layers * 4 + 16));
    CUDA_TRY(ctx, ctx->d_lbeg.ensure((size_t)n_layers * 4 + 16));
    CUDA_TRY(ctx, ctx->d_llen.ensure((size_t)n_layers * 4 + 16));
    CUDA_TRY(ctx, ctx->d_job_off.ensure((size_t)(n_layers + 1) * 8));
    uint8_t *d_lseq = ctx->d_poa_arena.as<uint8_t>(), *d_lqual = d_lseq + arena;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lsrc.p, layer_src, (size_t)n_layers * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lbeg.p, layer_begin, (size_t)n_layers * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_llen.p, layer_len, (size_t)n_layers * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_job_off.p, lay_off.data(), (size_t)(n_layers + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n_layers) {
        k_gather_layers<<<(unsigned)((n_layers + 7) / 8), 256, 0, ctx->stream>>>(
            ctx->d_seq.as<uint8_t>(), ctx->d_qual.as<uint8_t>(), ctx->d_off.as<int64_t>(),
            n_aux > 0 ? ctx->d_auxseq.as<uint8_t>() : nullptr, n_aux > 0 ? ctx->d_aoff.as<int64_t>() : nullptr,
            ctx->d_lsrc.as<int32_t>(), ctx->d_lbeg.as<int32_t>(), ctx->d_llen.as<int32_t>(), ctx->d_job_off.as<int64_t>(),
            n_layers, d_lseq, d_lqual);
        KERNEL_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(pin_lay.p, d_lseq, 2 * arena, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const uint8_t *h_lseq = pin_lay.as<uint8_t>(), *h_lqual = h_lseq + arena;

    const auto t_call = std::chrono::steady_clock::now();
    double ms_dev = 0.0, ms_host = 0.0;
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    const int T = host_threads(ctx);
    const int ring = (Lmax + 1 <= 8 * K5R_TILE) ? 16 : ((Lmax + 1 <= 16 * K5R_TILE) ? 8 : 4);
    std::vector<HostGraph> graphs((size_t)n_jobs);
    std::vector<int> jerr((size_t)n_jobs, 0);
    std::vector<int32_t> dp_jobs;                     // jobs that run the DP in this step
    std::vector<K5RJob> desc;
    std::vector<int64_t> n_ovf_of((size_t)n_jobs, 0);
    CUDA_TRY(ctx, ctx->d_poa_err.ensure(64));
    const int max_nodes = params->max_nodes;
    int64_t total_cells = 0;
    long long dbg[5] = {0, 0, 0, 0, 0}, dbg_dp = 0, dbg_tb = 0, dbg_wait = 0, ph[4] = {0, 0, 0, 0};

    for (int64_t step = 0; step < max_job_layers; ++step) {
        // ---- layers that need no DP (first layer of a job, empty layers) go straight into the graph
        auto t_h = std::chrono::steady_clock::now();
        dp_jobs.clear();
        for (int64_t j = 0; j < n_jobs; ++j) {
            if (jerr[j] || job_off[j] + step >= job_off[j + 1]) continue;
            const int64_t l = job_off[j] + step;
            const int L = layer_len[l];
            HostGraph &hg = graphs[j];
            hg.reserve(std::max(L, 1));
            if (hg.G.V == 0 || L == 0) {
                const uint8_t *q = layer_src[l] >= 0 ? h_lqual + lay_off[l] : nullptr;
                poa_add_alignment(hg.G, 0, h_lseq + lay_off[l], q, L);
                if (hg.G.err) jerr[j] = hg.G.err;
            } else dp_jobs.push_back((int32_t)j);
        }
        const int nd = (int)dp_jobs.size();
        if (nd == 0) { ms_host += since(t_h); continue; }
        // ---- sizes (a layer with a sub-graph range is aligned to that part of its graph only)
        int Lstep = 1, max_np = 0;
        std::vector<int> np_max((size_t)nd, 0);
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
        for (int x = 0; x < nd; ++x) {
            const int j = dp_jobs[x];
            const int64_t l = job_off[j] + step;
            HostGraph &hg = graphs[j];
            hg.sub = false;
            if (layer_sub_begin && layer_sub_begin[l] >= 0 && layer_sub_end[l] >= layer_sub_begin[l] && layer_sub_end[l] < hg.G.V) {
                hg.v_order.resize((size_t)hg.G.V); hg.v_rank.resize((size_t)hg.G.V); hg.v_member.resize((size_t)hg.G.V);
                hg.Vs = poa_subgraph_view(hg.G, layer_sub_begin[l], layer_sub_end[l], hg.v_member.data(), hg.v_order.data(), hg.v_rank.data());
                if (hg.Vs < 0) { jerr[j] = hg.G.err ? hg.G.err : 4; hg.Vs = 0; }
                hg.sub = true;
            }
            n_ovf_of[j] = build_rows(hg, ring, nullptr, nullptr, nullptr, true, &np_max[x]);
        }
        desc.assign((size_t)nd, K5RJob());
        int64_t meta_n = 0, ovf_n = 0, mat_n = 0, path_n = 0;
        for (int x = 0; x < nd; ++x) {
            const int j = dp_jobs[x];
            const int64_t l = job_off[j] + step;
            Lstep = std::max(Lstep, (int)layer_len[l]);
            max_np = std::max(max_np, np_max[x]);
        }
        if (max_np > K5R_MAXE) return fail(ctx, NGSID_EUNSUPPORTED, "a graph node has more than 119 in-edges");
        const int NW = (Lstep + 1 + K5R_TILE - 1) / K5R_TILE;
        const int ld = NW * K5R_TILE;
        for (int x = 0; x < nd; ++x) {
            const int j = dp_jobs[x];
            const int64_t l = job_off[j] + step;
            const PoaGraph &G = graphs[j].G;
            const int Vr = graphs[j].rows();
            K5RJob &D = desc[x];
            D.V = Vr; D.L = layer_len[l]; D.mode = params->mode;
            D.match = params->match; D.mismatch = params->mismatch; D.gap = params->gap;
            D.seq_off = lay_off[l];
            D.meta_off = meta_n; meta_n += Vr;
            D.ovf_off = ovf_n; ovf_n += n_ovf_of[j];
            D.mat_off = mat_n; mat_n += (int64_t)(Vr + 1) * ld;
            D.path_off = path_n; path_n += Vr + D.L + 2;
            total_cells += (int64_t)Vr * D.L;
            if (max_nodes > 0 && G.V > max_nodes) return fail(ctx, NGSID_EUNSUPPORTED, "POA graph larger than max_nodes");
        }
        // ---- rows of every graph into pinned memory, then to the device
        const size_t b_desc = ((size_t)nd * sizeof(K5RJob) + 255) & ~(size_t)255;
        const size_t b_meta = ((size_t)meta_n * 16 + 255) & ~(size_t)255;
        const size_t b_ovf = ((size_t)ovf_n * 4 + 255) & ~(size_t)255;
        const size_t b_plan = ((size_t)meta_n * 8 + 255) & ~(size_t)255;
        CUDA_TRY(ctx, pin_meta.ensure(b_desc + b_meta + b_ovf + b_plan));
        uint8_t *hm = pin_meta.as<uint8_t>();
        memcpy(hm, desc.data(), (size_t)nd * sizeof(K5RJob));
        uint4 *h_rows = reinterpret_cast<uint4 *>(hm + b_desc);
        int32_t *h_ovf = reinterpret_cast<int32_t *>(hm + b_desc + b_meta);
        uint2 *h_plan = reinterpret_cast<uint2 *>(hm + b_desc + b_meta + b_ovf);
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
        for (int x = 0; x < nd; ++x) {
            int dummy = 0;
            build_rows(graphs[dp_jobs[x]], ring, h_rows + desc[x].meta_off, h_plan + desc[x].meta_off, h_ovf + desc[x].ovf_off, false, &dummy);
        }
        ms_host += since(t_h);
        auto t_d = std::chrono::steady_clock::now();
        CUDA_TRY(ctx, ctx->d_poa_meta.ensure(b_desc + b_meta + b_ovf + b_plan));
        CUDA_TRY(ctx, ctx->d_poa_h.ensure((size_t)mat_n * 4 + 256));
        CUDA_TRY(ctx, ctx->d_poa_dir.ensure((size_t)mat_n + 256));
        CUDA_TRY(ctx, ctx->d_poa_out.ensure((size_t)path_n * 8 + (size_t)nd * 64 + 256));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_poa_meta.p, hm, b_desc + b_meta + b_ovf + b_plan, cudaMemcpyHostToDevice, ctx->stream));
        int32_t *d_out = reinterpret_cast<int32_t *>(ctx->d_poa_out.as<uint8_t>() + (size_t)path_n * 8);
        CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, (size_t)nd * 64, ctx->stream));
        K5RArgs A;
        A.jobs = ctx->d_poa_meta.as<K5RJob>();
        A.layers = d_lseq;
        A.meta = reinterpret_cast<const uint4 *>(ctx->d_poa_meta.as<uint8_t>() + b_desc);
        A.ovf = reinterpret_cast<const int32_t *>(ctx->d_poa_meta.as<uint8_t>() + b_desc + b_meta);
        A.plan = reinterpret_cast<const uint2 *>(ctx->d_poa_meta.as<uint8_t>() + b_desc + b_meta + b_ovf);
        A.H = ctx->d_poa_h.as<int32_t>(); A.DIR = ctx->d_poa_dir.as<uint8_t>();
        A.path = ctx->d_poa_out.as<int2>(); A.out = d_out; A.ld = ld; A.ring = ring;
        const size_t smem = k5r_smem_bytes(NW, ring);
#define K5R_LAUNCH(MODE, MAXT)                                                                                     \
        do {                                                                                                       \
            CUDA_TRY(ctx, cudaFuncSetAttribute(k5r_layer_kernel<MODE, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k5r_layer_kernel<MODE, MAXT><<<nd, NW * 32, smem, ctx->stream>>>(A);                                   \
        } while (0)
        if (params->mode) { if (NW <= 8) K5R_LAUNCH(1, 256); else K5R_LAUNCH(1, 1024); }
        else { if (NW <= 8) K5R_LAUNCH(0, 256); else K5R_LAUNCH(0, 1024); }
#undef K5R_LAUNCH
        KERNEL_CHECK(ctx);
        CUDA_TRY(ctx, pin_path.ensure((size_t)path_n * 8 + (size_t)nd * 64));
        CUDA_TRY(ctx, cudaMemcpyAsync(pin_path.p, ctx->d_poa_out.p, (size_t)path_n * 8 + (size_t)nd * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ms_dev += since(t_d);
        t_h = std::chrono::steady_clock::now();
        const int2 *h_path = pin_path.as<int2>();
        const int32_t *h_out = reinterpret_cast<const int32_t *>(pin_path.as<uint8_t>() + (size_t)path_n * 8);
        for (int x = 0; x < nd; ++x) {
            dbg[0] = std::max<long long>(dbg[0], h_out[x * 16 + 3]); dbg[1] = std::max<long long>(dbg[1], h_out[x * 16 + 4]);
            dbg[2] += h_out[x * 16 + 5]; dbg[3] += desc[x].V; dbg[4] = std::max<long long>(dbg[4], h_out[x * 16 + 6]);
            for (int q = 0; q < 4; ++q) ph[q] += h_out[x * 16 + 8 + q];
        }
        dbg_dp += dbg[0]; dbg_tb += dbg[1]; dbg_wait += dbg[4]; dbg[0] = dbg[1] = dbg[4] = 0;
        // ---- graph update + re-sort on host threads
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
        for (int x = 0; x < nd; ++x) {
            const int j = dp_jobs[x];
            const int64_t l = job_off[j] + step;
            PoaGraph &G = graphs[j].G;
            if (h_out[x * 16 + 2]) { jerr[j] = h_out[x * 16 + 2]; continue; }
            const int n = h_out[x * 16];
            const int2 *pp = h_path + desc[x].path_off;
            const int32_t *row_node = graphs[j].row_order();
            for (int t = 0; t < n; ++t) {
                G.aln_node[t] = pp[t].x > 0 ? row_node[pp[t].x - 1] : -1;
                G.aln_pos[t] = pp[t].y;
            }
            const uint8_t *q = layer_src[l] >= 0 ? h_lqual + lay_off[l] : nullptr;
            poa_add_alignment(G, n, h_lseq + lay_off[l], q, layer_len[l]);
            if (G.err) jerr[j] = G.err;
        }
        ms_host += since(t_h);
    }
    // ---- heaviest bundle per job
    int worst = 0;
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int64_t j = 0; j < n_jobs; ++j) {
        int len = -1;
        if (!jerr[j]) {
            if (graphs[j].init) len = poa_consensus(graphs[j].G, params->trim, out_seq + (size_t)j * out_stride, (int)out_stride);
            else len = 0;
            if (len < 0) jerr[j] = 5;
        }
        out_len[j] = len;
        if (out_nodes) out_nodes[j] = graphs[j].init ? graphs[j].G.V : 0;
    }
    for (int64_t j = 0; j < n_jobs; ++j) worst = std::max(worst, jerr[j]);
    ctx->poa_cells = total_cells;
    ctx->poa_ms[0] = (float)ms_dev; ctx->poa_ms[1] = (float)ms_host; ctx->poa_ms[2] = (float)since(t_call);
    if (getenv("NGSID_POA_TIMING"))
        fprintf(stderr, "[k5] jobs %lld layers %lld steps %lld threads %d: device+copies %.1f ms, host graphs %.1f ms, call %.1f ms, %.3g cells\n",
                (long long)n_jobs, (long long)n_layers, (long long)max_job_layers, T, ms_dev, ms_host, ctx->poa_ms[2], (double)total_cells);
    if (getenv("NGSID_POA_TIMING"))
        fprintf(stderr, "[k5]   slowest job per step, summed: DP %.1f Mcycles, traceback %.1f Mcycles, warp 1 waiting %.1f Mcycles; rows %lld, far-predecessor loads %lld\n",
                dbg_dp * 1024e-6, dbg_tb * 1024e-6, dbg_wait * 1024e-6, dbg[3], dbg[2]);
    if (getenv("NGSID_POA_TIMING"))
        fprintf(stderr, "[k5]   warp 1, all jobs: cycles per row in back-pressure %.0f, predecessors %.0f, scan %.0f, stores %.0f\n",
                ph[0] * 1024.0 / std::max<long long>(1, dbg[3]), ph[1] * 1024.0 / std::max<long long>(1, dbg[3]),
                ph[2] * 1024.0 / std::max<long long>(1, dbg[3]), ph[3] * 1024.0 / std::max<long long>(1, dbg[3]));
    if (worst) {
        char msg[160];
        snprintf(msg, sizeof msg, "POA failed (code %d: 1-3 graph capacity, 5 output buffer, 7 in-edges, 8 no end cell)", worst);
        return fail(ctx, NGSID_EUNSUPPORTED, msg);
    }
    return NGSID_OK;
}

extern "C" int ngsid_poa_consensus(ngsid_ctx *ctx, const ngsid_poa_params *params, int64_t n_jobs,
                                   const int64_t *job_off, const int32_t *layer_src, const int32_t *layer_begin,
                                   const int32_t *layer_len, const uint8_t *aux_seq, const int64_t *aux_off,
                                   int64_t n_aux, uint8_t *out_seq, int64_t out_stride, int32_t *out_len,
                                   int32_t *out_nodes)
{
    return poa_consensus_impl(ctx, params, n_jobs, job_off, layer_src, layer_begin, layer_len, nullptr, nullptr,
                              aux_seq, aux_off, n_aux, out_seq, out_stride, out_len, out_nodes);
}

extern "C" int ngsid_poa_consensus_sub(ngsid_ctx *ctx, const ngsid_poa_params *params, int64_t n_jobs,
                                       const int64_t *job_off, const int32_t *layer_src, const int32_t *layer_begin,
                                       const int32_t *layer_len, const int32_t *layer_sub_begin, const int32_t *layer_sub_end,
                                       const uint8_t *aux_seq, const int64_t *aux_off,
                                       int64_t n_aux, uint8_t *out_seq, int64_t out_stride, int32_t *out_len,
                                       int32_t *out_nodes)
{
    if ((layer_sub_begin == nullptr) != (layer_sub_end == nullptr)) return NGSID_EINVAL;
    return poa_consensus_impl(ctx, params, n_jobs, job_off, layer_src, layer_begin, layer_len, layer_sub_begin, layer_sub_end,
                              aux_seq, aux_off, n_aux, out_seq, out_stride, out_len, out_nodes);
}
