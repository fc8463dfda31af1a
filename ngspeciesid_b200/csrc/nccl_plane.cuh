// Multi-GPU data plane of libngsid.so: one process per GPU, NCCL over NVLink / NVSwitch.
// Replaces the exchange step of the reference's --t N mode (modules/parallelize.py:153-187: the
// process pool hands (clusters, representatives, minimizer_database) of every batch back to the
// parent through pickles) and the per-cluster read files of the consensus step
// (modules/consensus.py:249-278, 186-246: reads of a cluster are written to a FASTQ that spoa /
// racon read back). Here the representatives travel device to device with their minimizer records
// and quality statistics (nothing is recomputed on the receiving side), and the reads of a cluster
// travel to the GPU that owns its consensus in one all-to-all.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy a host process has already loaded,
// e.g. torch's, or the system one), so the library has no link-time dependency on it and the
// single-GPU entry points work without it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace ncclplane {

struct Api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

static Api &api_state() { static Api A; return A; }

static Api *api()
{
    Api &A = api_state();
    static bool tried = false;
    if (tried) return A.handle ? &A : nullptr;
    tried = true;
    const char *names[] = {getenv("NGSID_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        A.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (A.handle) break;
    }
    if (!A.handle) { A.why = "libnccl.so.2 not found (set NGSID_NCCL_LIB)"; return nullptr; }
#define NCCL_SYM(field, name)                                                                   \
    *(void **)(&A.field) = dlsym(A.handle, name);                                               \
    if (!A.field) { A.why = std::string("missing symbol ") + name; dlclose(A.handle); A.handle = nullptr; return nullptr; }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank") NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllGather, "ncclAllGather") NCCL_SYM(AllReduce, "ncclAllReduce") NCCL_SYM(Send, "ncclSend") NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart") NCCL_SYM(GroupEnd, "ncclGroupEnd") NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return &A;
}

#define NCCL_TRY(ctx, call)                                                                     \
    do {                                                                                        \
        ncclResult_t _r = (call);                                                               \
        if (_r != ncclSuccess) {                                                                \
            (ctx)->err = std::string(#call) + ": " + ncclplane::api()->GetErrorString(_r);      \
            return NGSID_ECUDA;                                                                 \
        }                                                                                       \
    } while (0)

// ---- gather kernels: one warp per selected read -------------------------------------------------
struct GatherArgs {
    const int32_t *sel; int64_t n;
    const uint8_t *seq, *qual; const int64_t *off;
    const Minimizer *mins; const int64_t *moff; const uint32_t *nmin, *lenc;
    const double *errc, *erru; const uint8_t *bucket;
    const int64_t *so, *mo;                 // per selected read: offset of its bases / minimizers in the send buffers
    uint8_t *o_seq, *o_qual; Minimizer *o_mins;
    int32_t *o_len; uint32_t *o_lenc, *o_nmin; double *o_errc, *o_erru; uint8_t *o_bucket;
};

__global__ void k_gather_reads(GatherArgs G)
{
    const int64_t wi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wi >= G.n) return;
    const uint32_t lane = lane_id();
    const int32_t r = G.sel[wi];
    const int64_t a = G.off[r];
    const int L = (int)(G.off[r + 1] - a);
    uint8_t *os = G.o_seq + G.so[wi], *oq = G.o_qual + G.so[wi];
    for (int i = lane; i < L; i += 32) { os[i] = G.seq[a + i]; oq[i] = G.qual[a + i]; }
    if (G.mins) {
        const Minimizer *m = G.mins + G.moff[r];
        const int nm = (int)G.nmin[r];
        Minimizer *om = G.o_mins + G.mo[wi];
        for (int i = lane; i < nm; i += 32) om[i] = m[i];
        if (lane == 0) {
            G.o_lenc[wi] = G.lenc[r]; G.o_nmin[wi] = G.nmin[r];
            G.o_errc[wi] = G.errc[r]; G.o_erru[wi] = G.erru[r]; G.o_bucket[wi] = G.bucket[r];
        }
    }
    if (lane == 0) G.o_len[wi] = L;
}

// dense minimizer records (per read: src[mo[i] .. mo[i] + nmin[i])) -> the slack layout of a context
__global__ void k_scatter_mins(const Minimizer *__restrict__ src, const int64_t *__restrict__ mo,
                               const uint32_t *__restrict__ nmin, const int64_t *__restrict__ moff,
                               Minimizer *__restrict__ dst, int64_t n)
{
    const int64_t wi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wi >= n) return;
    const uint32_t lane = lane_id();
    const Minimizer *s = src + mo[wi];
    Minimizer *d = dst + moff[wi];
    const int nm = (int)nmin[wi];
    for (int i = lane; i < nm; i += 32) d[i] = s[i];
}

// reads [0, n) of a context -> reads [n, 2n): reverse complement, qualities reversed
__global__ void k_revcomp(uint8_t *__restrict__ seq, uint8_t *__restrict__ qual, const int64_t *__restrict__ off, int64_t n)
{
    const int64_t wi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wi >= n) return;
    const uint32_t lane = lane_id();
    const int64_t a = off[wi], b = off[n + wi];
    const int L = (int)(off[wi + 1] - a);
    for (int i = lane; i < L; i += 32) {
        const uint8_t c = seq[a + L - 1 - i];
        seq[b + i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
        qual[b + i] = qual[a + L - 1 - i];
    }
}

}  // namespace ncclplane

// ================================================================================ communicator
extern "C" int ngsid_nccl_unique_id(uint8_t *out, int64_t cap)
{
    if (!out || cap < (int64_t)sizeof(ncclUniqueId)) return NGSID_EINVAL;
    ncclplane::Api *N = ncclplane::api();
    if (!N) return NGSID_EUNSUPPORTED;
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return NGSID_ECUDA;
    memcpy(out, &id, sizeof id);
    return (int)sizeof id;
}

extern "C" int ngsid_nccl_init(ngsid_ctx *ctx, const uint8_t *unique_id, int rank, int nranks)
{
    if (!ctx || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return NGSID_EINVAL;
    ncclplane::Api *N = ncclplane::api();
    if (!N) return fail(ctx, NGSID_EUNSUPPORTED, "NCCL is not available: " + ncclplane::api_state().why);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm && ctx->nccl_owned) N->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm;
    NCCL_TRY(ctx, N->CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm; ctx->nccl_owned = true; ctx->nccl_rank = rank; ctx->nccl_nranks = nranks;
    return NGSID_OK;
}

// A second context on the same GPU (merge / consensus contexts) uses the communicator of the first.
extern "C" int ngsid_nccl_share(ngsid_ctx *ctx, ngsid_ctx *owner)
{
    if (!ctx || !owner || ctx == owner) return NGSID_EINVAL;
    if (ctx->device != owner->device) return fail(ctx, NGSID_EINVAL, "contexts live on different GPUs");
    if (ctx->nccl_comm && ctx->nccl_owned) ncclplane::api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = owner->nccl_comm; ctx->nccl_owned = false;
    ctx->nccl_rank = owner->nccl_rank; ctx->nccl_nranks = owner->nccl_nranks;
    return NGSID_OK;
}

extern "C" int ngsid_nccl_finalize(ngsid_ctx *ctx)
{
    if (!ctx) return NGSID_EINVAL;
    if (ctx->nccl_comm && !ctx->nccl_owned) ctx->nccl_comm = nullptr;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        ncclplane::api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nccl_nranks = 1; ctx->nccl_rank = 0;
    return NGSID_OK;
}

// A context without a communicator behaves as a world of one rank, so the same driver code runs
// on one GPU without NCCL.
static inline int nccl_world(const ngsid_ctx *ctx) { return ctx->nccl_comm ? ctx->nccl_nranks : 1; }

// ================================================================================ small collectives
// Variable-size all-gather of host bytes (accession strings, consensus strings, plans):
// recv = rank 0's bytes | rank 1's bytes | ...; counts[r] = bytes of rank r.
extern "C" int ngsid_allgather_bytes(ngsid_ctx *ctx, const uint8_t *send, int64_t n_send, uint8_t *recv,
                                     int64_t recv_cap, int64_t *counts)
{
    if (!ctx || n_send < 0 || (n_send > 0 && !send) || !counts) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int W = nccl_world(ctx);
    if (W == 1) {
        counts[0] = n_send;
        if (!recv) return NGSID_OK;
        if (n_send > recv_cap) return fail(ctx, NGSID_EINVAL, "allgather_bytes: receive buffer too small");
        if (n_send) memcpy(recv, send, (size_t)n_send);
        return NGSID_OK;
    }
    ncclplane::Api *N = ncclplane::api();
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    CUDA_TRY(ctx, ctx->d_cc_a.ensure((size_t)W * 8 + 64));
    int64_t *d_cnt = ctx->d_cc_a.as<int64_t>();
    CUDA_TRY(ctx, cudaMemcpyAsync(d_cnt + ctx->nccl_rank, &n_send, 8, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, N->AllGather(d_cnt + ctx->nccl_rank, d_cnt, 1, ncclInt64, comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(counts, d_cnt, (size_t)W * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t mx = 0, total = 0;
    for (int r = 0; r < W; ++r) { mx = std::max(mx, counts[r]); total += counts[r]; }
    if (!recv) return NGSID_OK;                      // sizes only: the caller allocates and calls again
    if (total > recv_cap) return fail(ctx, NGSID_EINVAL, "allgather_bytes: receive buffer too small");
    if (mx == 0) return NGSID_OK;
    const size_t slot = ((size_t)mx + 15) & ~(size_t)15;
    CUDA_TRY(ctx, ctx->d_cc_b.ensure(slot * W + 64));
    uint8_t *buf = ctx->d_cc_b.as<uint8_t>();
    if (n_send) CUDA_TRY(ctx, cudaMemcpyAsync(buf + slot * ctx->nccl_rank, send, (size_t)n_send, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, N->AllGather(buf + slot * ctx->nccl_rank, buf, slot, ncclChar, comm, ctx->stream));
    int64_t o = 0;
    for (int r = 0; r < W; ++r) {
        if (counts[r]) CUDA_TRY(ctx, cudaMemcpyAsync(recv + o, buf + slot * r, (size_t)counts[r], cudaMemcpyDeviceToHost, ctx->stream));
        o += counts[r];
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

// In-place all-reduce of a host int32 / int64 array (op 0 sum, 1 max): decisions of a merge round,
// cluster sizes.
extern "C" int ngsid_allreduce(ngsid_ctx *ctx, void *buf, int64_t n, int elem_bytes, int op)
{
    if (!ctx || n < 0 || (n > 0 && !buf) || (elem_bytes != 4 && elem_bytes != 8) || op < 0 || op > 1) return NGSID_EINVAL;
    if (n == 0 || nccl_world(ctx) == 1) return NGSID_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclplane::Api *N = ncclplane::api();
    CUDA_TRY(ctx, ctx->d_cc_b.ensure((size_t)n * elem_bytes + 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_cc_b.p, buf, (size_t)n * elem_bytes, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, N->AllReduce(ctx->d_cc_b.p, ctx->d_cc_b.p, (size_t)n, elem_bytes == 4 ? ncclInt32 : ncclInt64,
                               op == 0 ? ncclSum : ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(buf, ctx->d_cc_b.p, (size_t)n * elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

// ================================================================================ representatives
// All ranks call this with their surviving representatives (read indices of `ctx`, in processing
// order). Afterwards `dst` (a second context on the same GPU) holds the representatives of every
// rank -- rank 0's first -- as an uploaded read set WITH its K1 results (minimizer records, counts,
// compressed lengths) and K0 results (error rates, buckets), moved device to device:
// ngsid_cluster() can run on `dst` right away. counts[r] = representatives that came from rank r.
extern "C" int ngsid_gather_representatives(ngsid_ctx *ctx, const int32_t *reps, int64_t n_reps, ngsid_ctx *dst,
                                            int64_t *counts)
{
    if (!ctx || !dst || dst == ctx || n_reps < 0 || (n_reps > 0 && !reps) || !counts) return NGSID_EINVAL;
    if (!ctx->have_min || !ctx->have_q) return fail(ctx, NGSID_ESTATE, "run ngsid_minimizers and ngsid_quality_stats first");
    if (dst->device != ctx->device) return fail(ctx, NGSID_EINVAL, "destination context must live on the same GPU");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = fetch_nmin(ctx);
    if (rc) return rc;
    const int W = nccl_world(ctx), me = ctx->nccl_comm ? ctx->nccl_rank : 0;
    ncclplane::Api *N = W > 1 ? ncclplane::api() : nullptr;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // ---- local totals and send offsets
    std::vector<int64_t> so(n_reps + 1, 0), mo(n_reps + 1, 0);
    for (int64_t i = 0; i < n_reps; ++i) {
        const int32_t r = reps[i];
        if (r < 0 || r >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "representative index out of range");
        so[i + 1] = so[i] + (ctx->h_off[r + 1] - ctx->h_off[r]);
        mo[i + 1] = mo[i] + ctx->h_nmin[r];
    }
    int64_t mine[4] = {n_reps, so[n_reps], mo[n_reps], 0};
    std::vector<int64_t> all((size_t)W * 4);
    if (W > 1) {
        CUDA_TRY(ctx, ctx->d_cc_a.ensure((size_t)W * 32 + 64));
        int64_t *d = ctx->d_cc_a.as<int64_t>();
        CUDA_TRY(ctx, cudaMemcpyAsync(d + 4 * me, mine, 32, cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ctx, N->AllGather(d + 4 * me, d, 4, ncclInt64, comm, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), d, (size_t)W * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(all.data(), mine, 32);
    }
    int64_t mx_n = 0, mx_b = 0, mx_m = 0, tot_n = 0;
    for (int r = 0; r < W; ++r) {
        counts[r] = all[4 * r];
        mx_n = std::max(mx_n, all[4 * r]); mx_b = std::max(mx_b, all[4 * r + 1]); mx_m = std::max(mx_m, all[4 * r + 2]);
        tot_n += all[4 * r];
    }
    if (tot_n == 0) { int64_t z = 0; return reads_layout(dst, &z, 0); }
    // ---- one slot per rank in every gathered array (slot = the largest rank's size, 16-byte multiples)
    auto up = [](int64_t x) { return (size_t)((x + 15) & ~(int64_t)15); };
    const size_t s_b = up(mx_b), s_m = up(mx_m * 8), s_n4 = up(mx_n * 4), s_n8 = up(mx_n * 8), s_n1 = up(mx_n);
    const size_t o_seq = 0, o_qual = o_seq + s_b * W, o_mins = o_qual + s_b * W, o_len = o_mins + s_m * W,
                 o_lenc = o_len + s_n4 * W, o_nmin = o_lenc + s_n4 * W, o_errc = o_nmin + s_n4 * W,
                 o_erru = o_errc + s_n8 * W, o_bucket = o_erru + s_n8 * W, total = o_bucket + s_n1 * W;
    CUDA_TRY(ctx, ctx->d_cc_b.ensure(total + 64));
    uint8_t *B = ctx->d_cc_b.as<uint8_t>();
    CUDA_TRY(ctx, ctx->d_cc_c.ensure((size_t)(n_reps + 1) * 20 + 64));
    int64_t *d_so = ctx->d_cc_c.as<int64_t>(), *d_mo = d_so + (n_reps + 1);
    int32_t *d_sel = reinterpret_cast<int32_t *>(d_mo + (n_reps + 1));
    if (n_reps) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_so, so.data(), (size_t)(n_reps + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_mo, mo.data(), (size_t)(n_reps + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_sel, reps, (size_t)n_reps * 4, cudaMemcpyHostToDevice, ctx->stream));
        ncclplane::GatherArgs G;
        G.sel = d_sel; G.n = n_reps;
        G.seq = ctx->d_seq.as<uint8_t>(); G.qual = ctx->d_qual.as<uint8_t>(); G.off = ctx->d_off.as<int64_t>();
        G.mins = ctx->d_mins.as<Minimizer>(); G.moff = ctx->d_moff.as<int64_t>();
        G.nmin = ctx->d_nmin.as<uint32_t>(); G.lenc = ctx->d_lenc.as<uint32_t>();
        G.errc = ctx->d_errc.as<double>(); G.erru = ctx->d_erru.as<double>(); G.bucket = ctx->d_bucket.as<uint8_t>();
        G.so = d_so; G.mo = d_mo;
        G.o_seq = B + o_seq + s_b * me; G.o_qual = B + o_qual + s_b * me;
        G.o_mins = reinterpret_cast<Minimizer *>(B + o_mins + s_m * me);
        G.o_len = reinterpret_cast<int32_t *>(B + o_len + s_n4 * me);
        G.o_lenc = reinterpret_cast<uint32_t *>(B + o_lenc + s_n4 * me);
        G.o_nmin = reinterpret_cast<uint32_t *>(B + o_nmin + s_n4 * me);
        G.o_errc = reinterpret_cast<double *>(B + o_errc + s_n8 * me);
        G.o_erru = reinterpret_cast<double *>(B + o_erru + s_n8 * me);
        G.o_bucket = B + o_bucket + s_n1 * me;
        ncclplane::k_gather_reads<<<(unsigned)((n_reps + 7) / 8), 256, 0, ctx->stream>>>(G);
        KERNEL_CHECK(ctx);
    }
    if (W > 1) {
        NCCL_TRY(ctx, N->GroupStart());
        const size_t offs[9] = {o_seq, o_qual, o_mins, o_len, o_lenc, o_nmin, o_errc, o_erru, o_bucket};
        const size_t slots[9] = {s_b, s_b, s_m, s_n4, s_n4, s_n4, s_n8, s_n8, s_n1};
        for (int a = 0; a < 9; ++a)
            if (slots[a]) NCCL_TRY(ctx, N->AllGather(B + offs[a] + slots[a] * me, B + offs[a], slots[a], ncclChar, comm, ctx->stream));
        NCCL_TRY(ctx, N->GroupEnd());
    }
    // ---- lengths of every representative -> layout of dst
    std::vector<int32_t> lens((size_t)tot_n);
    std::vector<uint32_t> nmins((size_t)tot_n);
    {
        int64_t o = 0;
        for (int r = 0; r < W; ++r) {
            if (counts[r]) {
                CUDA_TRY(ctx, cudaMemcpyAsync(lens.data() + o, B + o_len + s_n4 * r, (size_t)counts[r] * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(nmins.data() + o, B + o_nmin + s_n4 * r, (size_t)counts[r] * 4, cudaMemcpyDeviceToHost, ctx->stream));
            }
            o += counts[r];
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int64_t> offs((size_t)tot_n + 1, 0), dmo((size_t)tot_n + 1, 0);
    for (int64_t i = 0; i < tot_n; ++i) { offs[i + 1] = offs[i] + lens[i]; dmo[i + 1] = dmo[i] + nmins[i]; }
    // dst works on ctx's stream order: everything below is enqueued on dst's stream after a sync of ctx's
    CUDA_TRY(ctx, cudaStreamSynchronize(dst->stream));
    struct StreamSwap {                               // same device: run the install on the producing stream
        ngsid_ctx *d; cudaStream_t keep;
        StreamSwap(ngsid_ctx *d_, cudaStream_t s) : d(d_), keep(d_->stream) { d->stream = s; }
        ~StreamSwap() { d->stream = keep; }
    } swap(dst, ctx->stream);
    auto restore = [&](int code) { if (code && !dst->err.empty()) ctx->err = dst->err; return code; };
    rc = reads_layout(dst, offs.data(), tot_n);
    if (rc) return restore(rc);
    {
        int64_t ob = 0;
        for (int r = 0; r < W; ++r) {
            const size_t nb = (size_t)all[4 * r + 1];
            if (nb) {
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_seq.as<uint8_t>() + ob, B + o_seq + s_b * r, nb, cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_qual.as<uint8_t>() + ob, B + o_qual + s_b * r, nb, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            ob += (int64_t)nb;
        }
    }
    rc = reads_finish(dst);
    if (rc) return restore(rc);
    rc = k1_prepare(dst, ctx->k, ctx->w);             // slack layout of the minimizer records
    if (rc) return restore(rc);
    CUDA_TRY(ctx, dst->d_errc.ensure((size_t)(tot_n + 1) * 8));
    CUDA_TRY(ctx, dst->d_erru.ensure((size_t)(tot_n + 1) * 8));
    CUDA_TRY(ctx, dst->d_bucket.ensure((size_t)tot_n + 64));
    CUDA_TRY(ctx, ctx->d_cc_a.ensure((size_t)(tot_n + 1) * 8 + 64));
    {
        // dense minimizer offsets are per rank slot: rebase them
        std::vector<int64_t> src_mo((size_t)tot_n);
        int64_t i0 = 0;
        for (int r = 0; r < W; ++r) {
            const int64_t base = (int64_t)((o_mins + s_m * r) / 8);
            for (int64_t i = 0; i < counts[r]; ++i) src_mo[i0 + i] = base + (dmo[i0 + i] - dmo[i0]);
            if (counts[r]) {
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_nmin.as<uint32_t>() + i0, B + o_nmin + s_n4 * r, (size_t)counts[r] * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_lenc.as<uint32_t>() + i0, B + o_lenc + s_n4 * r, (size_t)counts[r] * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_errc.as<double>() + i0, B + o_errc + s_n8 * r, (size_t)counts[r] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_erru.as<double>() + i0, B + o_erru + s_n8 * r, (size_t)counts[r] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_bucket.as<uint8_t>() + i0, B + o_bucket + s_n1 * r, (size_t)counts[r], cudaMemcpyDeviceToDevice, ctx->stream));
            }
            i0 += counts[r];
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_cc_a.p, src_mo.data(), (size_t)tot_n * 8, cudaMemcpyHostToDevice, ctx->stream));
        ncclplane::k_scatter_mins<<<(unsigned)((tot_n + 7) / 8), 256, 0, ctx->stream>>>(
            reinterpret_cast<const Minimizer *>(B), ctx->d_cc_a.as<int64_t>(), dst->d_nmin.as<uint32_t>(),
            dst->d_moff.as<int64_t>(), dst->d_mins.as<Minimizer>(), tot_n);
        KERNEL_CHECK(ctx);
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // src_mo goes out of scope
    }
    dst->have_min = true; dst->have_q = true; dst->h_nmin_valid = false;
    return restore(NGSID_OK);
}

// ================================================================================ reads to consensus owners
// All-to-all of reads: this rank sends read read_idx[i] of `ctx` (bases + qualities) to rank
// dest[i] with a caller-defined 64-bit tag (e.g. cluster ordinal << 32 | position in the cluster's
// read order); entries must be grouped by destination in ascending rank order. Afterwards `dst`
// holds what this rank received as an uploaded read set (packed, ready for K4 / K5): first the
// reads sent by rank 0 in the order it listed them, then rank 1's, ... out_tag (capacity tag_cap)
// receives their tags, recv_counts[r] the number of reads that came from rank r.
extern "C" int ngsid_exchange_reads(ngsid_ctx *ctx, const int32_t *read_idx, const int32_t *dest, const int64_t *tag,
                                    int64_t n_send, ngsid_ctx *dst, int64_t *out_tag, int64_t tag_cap, int64_t *recv_counts)
{
    if (!ctx || !dst || dst == ctx || n_send < 0 || (n_send > 0 && (!read_idx || !dest || !tag)) || !recv_counts) return NGSID_EINVAL;
    if (dst->device != ctx->device) return fail(ctx, NGSID_EINVAL, "destination context must live on the same GPU");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int W = nccl_world(ctx), me = ctx->nccl_comm ? ctx->nccl_rank : 0;
    ncclplane::Api *N = W > 1 ? ncclplane::api() : nullptr;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // ---- per destination: reads and bases
    std::vector<int64_t> cnt((size_t)W * 2, 0), so(n_send + 1, 0);
    for (int64_t i = 0; i < n_send; ++i) {
        const int32_t r = read_idx[i], d = dest[i];
        if (r < 0 || r >= ctx->n_reads || d < 0 || d >= W) return fail(ctx, NGSID_EINVAL, "exchange_reads: index out of range");
        if (i && d < dest[i - 1]) return fail(ctx, NGSID_EINVAL, "exchange_reads: entries must be grouped by ascending destination");
        const int64_t L = ctx->h_off[r + 1] - ctx->h_off[r];
        so[i + 1] = so[i] + L;
        cnt[2 * d] += 1; cnt[2 * d + 1] += L;
    }
    std::vector<int64_t> mat((size_t)W * W * 2);          // mat[(src * W + dst) * 2 + {reads, bases}]
    if (W > 1) {
        CUDA_TRY(ctx, ctx->d_cc_a.ensure((size_t)W * W * 16 + 64));
        int64_t *d = ctx->d_cc_a.as<int64_t>();
        CUDA_TRY(ctx, cudaMemcpyAsync(d + (size_t)me * W * 2, cnt.data(), (size_t)W * 16, cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ctx, N->AllGather(d + (size_t)me * W * 2, d, (size_t)W * 2, ncclInt64, comm, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(mat.data(), d, (size_t)W * W * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        mat = cnt;
    }
    int64_t rn = 0, rb = 0;
    std::vector<int64_t> r_n0(W + 1, 0), r_b0(W + 1, 0), s_n0(W + 1, 0), s_b0(W + 1, 0);
    for (int r = 0; r < W; ++r) {
        recv_counts[r] = mat[((size_t)r * W + me) * 2];
        r_n0[r + 1] = r_n0[r] + mat[((size_t)r * W + me) * 2];
        r_b0[r + 1] = r_b0[r] + mat[((size_t)r * W + me) * 2 + 1];
        s_n0[r + 1] = s_n0[r] + cnt[2 * r];
        s_b0[r + 1] = s_b0[r] + cnt[2 * r + 1];
    }
    rn = r_n0[W]; rb = r_b0[W];
    if (rn > tag_cap || (rn > 0 && !out_tag)) return fail(ctx, NGSID_EINVAL, "exchange_reads: tag buffer too small");
    // ---- send side: gather the selected reads into contiguous buffers (destination order)
    const size_t sb = (size_t)so[n_send];
    CUDA_TRY(ctx, ctx->d_cc_b.ensure(2 * sb + (size_t)n_send * 12 + 256));
    uint8_t *S = ctx->d_cc_b.as<uint8_t>();
    uint8_t *s_seq = S, *s_qual = S + ((sb + 15) & ~(size_t)15);
    int64_t *s_tag = reinterpret_cast<int64_t *>(s_qual + ((sb + 15) & ~(size_t)15));
    int32_t *s_len = reinterpret_cast<int32_t *>(s_tag + n_send);
    CUDA_TRY(ctx, ctx->d_cc_c.ensure((size_t)(n_send + 1) * 12 + 64));
    int64_t *d_so = ctx->d_cc_c.as<int64_t>();
    int32_t *d_sel = reinterpret_cast<int32_t *>(d_so + (n_send + 1));
    if (n_send) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_so, so.data(), (size_t)(n_send + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_sel, read_idx, (size_t)n_send * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(s_tag, tag, (size_t)n_send * 8, cudaMemcpyHostToDevice, ctx->stream));
        ncclplane::GatherArgs G;
        memset(&G, 0, sizeof G);
        G.sel = d_sel; G.n = n_send;
        G.seq = ctx->d_seq.as<uint8_t>(); G.qual = ctx->d_qual.as<uint8_t>(); G.off = ctx->d_off.as<int64_t>();
        G.so = d_so; G.o_seq = s_seq; G.o_qual = s_qual; G.o_len = s_len;
        ncclplane::k_gather_reads<<<(unsigned)((n_send + 7) / 8), 256, 0, ctx->stream>>>(G);
        KERNEL_CHECK(ctx);
    }
    // ---- receive side
    CUDA_TRY(ctx, dst->d_cc_b.ensure(2 * (((size_t)rb + 15) & ~(size_t)15) + (size_t)rn * 12 + 256));
    uint8_t *R = dst->d_cc_b.as<uint8_t>();
    uint8_t *r_seq = R, *r_qual = R + (((size_t)rb + 15) & ~(size_t)15);
    int64_t *r_tag = reinterpret_cast<int64_t *>(r_qual + (((size_t)rb + 15) & ~(size_t)15));
    int32_t *r_len = reinterpret_cast<int32_t *>(r_tag + rn);
    if (W > 1) {
        NCCL_TRY(ctx, N->GroupStart());
        for (int p = 0; p < W; ++p) {
            const int64_t sn = cnt[2 * p], sbp = cnt[2 * p + 1];
            const int64_t qn = mat[((size_t)p * W + me) * 2], qb = mat[((size_t)p * W + me) * 2 + 1];
            if (sn) {
                NCCL_TRY(ctx, N->Send(s_seq + s_b0[p], (size_t)sbp, ncclChar, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Send(s_qual + s_b0[p], (size_t)sbp, ncclChar, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Send(s_tag + s_n0[p], (size_t)sn, ncclInt64, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Send(s_len + s_n0[p], (size_t)sn, ncclInt32, p, comm, ctx->stream));
            }
            if (qn) {
                NCCL_TRY(ctx, N->Recv(r_seq + r_b0[p], (size_t)qb, ncclChar, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Recv(r_qual + r_b0[p], (size_t)qb, ncclChar, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Recv(r_tag + r_n0[p], (size_t)qn, ncclInt64, p, comm, ctx->stream));
                NCCL_TRY(ctx, N->Recv(r_len + r_n0[p], (size_t)qn, ncclInt32, p, comm, ctx->stream));
            }
        }
        NCCL_TRY(ctx, N->GroupEnd());
    } else if (n_send) {
        CUDA_TRY(ctx, cudaMemcpyAsync(r_seq, s_seq, sb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(r_qual, s_qual, sb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(r_tag, s_tag, (size_t)n_send * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(r_len, s_len, (size_t)n_send * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    std::vector<int32_t> lens((size_t)rn);
    if (rn) {
        CUDA_TRY(ctx, cudaMemcpyAsync(lens.data(), r_len, (size_t)rn * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(out_tag, r_tag, (size_t)rn * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> offs((size_t)rn + 1, 0);
    for (int64_t i = 0; i < rn; ++i) offs[i + 1] = offs[i] + lens[i];
    if (offs[rn] != rb) return fail(ctx, NGSID_ECUDA, "exchange_reads: received lengths do not add up");
    CUDA_TRY(ctx, cudaStreamSynchronize(dst->stream));
    int rc = reads_layout(dst, offs.data(), rn);
    if (rc) { ctx->err = dst->err; return rc; }
    if (rb) {
        CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_seq.p, r_seq, (size_t)rb, cudaMemcpyDeviceToDevice, dst->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(dst->d_qual.p, r_qual, (size_t)rb, cudaMemcpyDeviceToDevice, dst->stream));
    }
    rc = reads_finish(dst);
    if (rc) { ctx->err = dst->err; return rc; }
    return NGSID_OK;
}

// Doubles the read set of a context: read n + i becomes the reverse complement of read i (qualities
// reversed). The polishing step aligns the reads of a merged centre in both orientations
// (modules/consensus.py:148-183 merges reverse-complement clusters; minimap2 picks the strand).
extern "C" int ngsid_append_revcomp(ngsid_ctx *ctx)
{
    if (!ctx) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->n_reads;
    if (n == 0) return NGSID_OK;
    if (2 * n >= (int64_t)1 << 31) return fail(ctx, NGSID_EINVAL, "too many reads");
    const size_t nb = (size_t)ctx->total_bases;
    std::vector<int64_t> offs((size_t)2 * n + 1);
    for (int64_t i = 0; i <= n; ++i) offs[i] = ctx->h_off[i];
    for (int64_t i = 1; i <= n; ++i) offs[n + i] = offs[n] + ctx->h_off[i];
    // keep the bases: reads_layout may reallocate d_seq / d_qual
    DevBuf keep_s, keep_q;
    CUDA_TRY(ctx, keep_s.ensure(nb + 64));
    CUDA_TRY(ctx, keep_q.ensure(nb + 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(keep_s.p, ctx->d_seq.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(keep_q.p, ctx->d_qual.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = reads_layout(ctx, offs.data(), 2 * n);
    if (rc == NGSID_OK) {
        cudaMemcpyAsync(ctx->d_seq.p, keep_s.p, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(ctx->d_qual.p, keep_q.p, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        ncclplane::k_revcomp<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(ctx->d_seq.as<uint8_t>(), ctx->d_qual.as<uint8_t>(),
                                                                           ctx->d_off.as<int64_t>(), n);
        ctx->launches++;
        rc = reads_finish(ctx);
    }
    cudaStreamSynchronize(ctx->stream);
    keep_s.release(); keep_q.release();
    return rc;
}

// Host copy of the bases and qualities of reads [begin, end) of a context (reads that arrived through
// ngsid_exchange_reads have no host copy on this rank).
extern "C" int ngsid_download_reads(ngsid_ctx *ctx, int64_t begin, int64_t end, uint8_t *seq, uint8_t *qual, int64_t *offsets)
{
    if (!ctx || begin < 0 || end < begin || end > ctx->n_reads || !offsets) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t a = ctx->h_off[begin], b = ctx->h_off[end];
    for (int64_t i = begin; i <= end; ++i) offsets[i - begin] = ctx->h_off[i] - a;
    if (b > a) {
        if (seq) CUDA_TRY(ctx, cudaMemcpyAsync(seq, ctx->d_seq.as<uint8_t>() + a, (size_t)(b - a), cudaMemcpyDeviceToHost, ctx->stream));
        if (qual) CUDA_TRY(ctx, cudaMemcpyAsync(qual, ctx->d_qual.as<uint8_t>() + a, (size_t)(b - a), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NGSID_OK;
}
