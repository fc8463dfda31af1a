// libngsid.so: C ABI (include/ngsid.h) + host-side orchestration of the sm_100a kernels.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "ngsid_internal.cuh"
#include "k1_minimizers.cuh"
#include "k0_sortscore.cuh"
#include "k1_stream.cuh"
#include "fastq_ingest.cuh"
#include "k2_map.cuh"
#include "k4_align.cuh"
#include "k4_trace.cuh"

static const size_t SMEM_BUDGET = 200 * 1024;

// ================================================================================ context
extern "C" int ngsid_version(void) { return 1; }

extern "C" int ngsid_ctx_create(int device_id, ngsid_ctx **out)
{
    if (!out) return NGSID_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return NGSID_ECUDA;
    if (device_id < 0 || device_id >= count) return NGSID_EINVAL;
    if (cudaSetDevice(device_id) != cudaSuccess) return NGSID_ECUDA;
    ngsid_ctx *ctx = new ngsid_ctx();
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->up_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->up_ev[1], cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return NGSID_ECUDA;
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 2; ++j)
            if (cudaEventCreate(&ctx->pev[i][j]) != cudaSuccess) { delete ctx; return NGSID_ECUDA; }
    {
        // K1 stream kernel: homopolymer-compression table (k1_stream.cuh), once per device
        std::vector<uint32_t> lut(1024);
        for (uint32_t i = 0; i < 1024; ++i) lut[i] = k1s_lut_entry(i);
        if (cudaMemcpyToSymbol(k1s_lut_dev, lut.data(), 4096) != cudaSuccess) { delete ctx; return NGSID_ECUDA; }
    }
    *out = ctx;
    return NGSID_OK;
}

extern "C" float ngsid_phase_ms(ngsid_ctx *ctx, int which)
{
    if (!ctx || which < 0 || which > 8) return -1.f;
    if (which >= 6) return ctx->poa_ms[which - 6];
    if (which >= 4) return ctx->pev_valid[which] ? ctx->phase_acc[which] : -1.f;
    if (!ctx->pev_valid[which]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(ctx->pev[which][1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, ctx->pev[which][0], ctx->pev[which][1]) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int ngsid_set_option(ngsid_ctx *ctx, int option, int value)
{
    if (!ctx) return NGSID_EINVAL;
    if (option == 1) {
        if (value < 0 || value > 1) return fail(ctx, NGSID_EINVAL, "option 1 takes 0 or 1");
        ctx->k1_variant = value == 0 ? 2 : 0;
        ctx->have_min = false;
        return NGSID_OK;
    }
    if (option == 2) { ctx->use_payload_k4 = value != 0; return NGSID_OK; }
    if (option == 3) {
        if (value < 0 || value > 2) return fail(ctx, NGSID_EINVAL, "option 3 takes 0, 1 or 2");
        ctx->k4_shape = value;
        return NGSID_OK;
    }
    if (option == 4) {
        if (value < 0 || value > 2) return fail(ctx, NGSID_EINVAL, "option 4 takes 0, 1 or 2");
        ctx->k4_tb = value;
        return NGSID_OK;
    }
    return fail(ctx, NGSID_EINVAL, "unknown option");
}

extern "C" int ngsid_nccl_finalize(ngsid_ctx *ctx);

extern "C" void ngsid_ctx_destroy(ngsid_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ngsid_nccl_finalize(ctx);
    DevBuf *bufs[] = {&ctx->d_seq, &ctx->d_qual, &ctx->d_off, &ctx->d_packed, &ctx->d_woff, &ctx->d_flag, &ctx->d_rflag,
                      &ctx->d_moff, &ctx->d_mins, &ctx->d_nmin, &ctx->d_lenc, &ctx->d_errc, &ctx->d_erru,
                      &ctx->d_bucket, &ctx->d_phred, &ctx->d_thr, &ctx->d_keys, &ctx->d_heads, &ctx->d_nodes,
                      &ctx->d_cursor, &ctx->d_slot_read, &ctx->d_slot_pos, &ctx->d_slot_state, &ctx->d_order,
                      &ctx->d_accrank, &ctx->d_dec, &ctx->d_aux, &ctx->d_via, &ctx->d_list, &ctx->d_scratch,
                      &ctx->d_params, &ctx->d_req, &ctx->d_reqn, &ctx->d_acache, &ctx->d_k4cnt, &ctx->d_k4score,
                      &ctx->d_newslots, &ctx->d_cc_a, &ctx->d_cc_b, &ctx->d_cc_c, &ctx->d_aovf, &ctx->d_aovf_head, &ctx->d_ss_tab, &ctx->d_ss_score, &ctx->d_ss_err, &ctx->d_poa_dir, &ctx->d_poa_arena, &ctx->d_poa_meta, &ctx->d_poa_h, &ctx->d_poa_out, &ctx->d_poa_len, &ctx->d_poa_nodes, &ctx->d_poa_err, &ctx->d_job_off, &ctx->d_lsrc, &ctx->d_lbeg, &ctx->d_llen, &ctx->d_trace, &ctx->d_ends, &ctx->d_auxseq, &ctx->d_aoff, &ctx->d_win, &ctx->d_match, &ctx->d_cols, &ctx->d_pa, &ctx->d_pb, &ctx->d_po, &ctx->d_pm,
                      &ctx->d_cl[0], &ctx->d_cl[1], &ctx->d_cl[2], &ctx->d_cl[3], &ctx->d_cl[4], &ctx->d_cl[5]};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 2; ++j) if (ctx->pev[i][j]) cudaEventDestroy(ctx->pev[i][j]);
    cudaEventDestroy(ctx->up_ev[0]); cudaEventDestroy(ctx->up_ev[1]);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *ngsid_last_error(const ngsid_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t ngsid_launch_count(const ngsid_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t ngsid_poa_cells(const ngsid_ctx *ctx) { return ctx ? ctx->poa_cells : 0; }
extern "C" void ngsid_reset_launch_count(ngsid_ctx *ctx) { if (ctx) ctx->launches = 0; }
extern "C" int ngsid_sync(ngsid_ctx *ctx)
{
    if (!ctx) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

extern "C" int ngsid_pinned_alloc(void **out, int64_t bytes)
{
    if (!out || bytes < 0) return NGSID_EINVAL;
    *out = nullptr;
    return cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? NGSID_OK : NGSID_ECUDA;
}

extern "C" int ngsid_pinned_free(void *p)
{
    return (!p || cudaFreeHost(p) == cudaSuccess) ? NGSID_OK : NGSID_ECUDA;
}

// ================================================================================ upload + pack
// Layout of a new read set: host offsets, word offsets of the packed reads, device buffers sized,
// d_off / d_woff uploaded. The caller then fills d_seq / d_qual (host or peer data) and calls
// reads_finish().
static int reads_layout(ngsid_ctx *ctx, const int64_t *offsets, int64_t n_reads)
{
    if (n_reads >= (int64_t)1 << 31) return fail(ctx, NGSID_EINVAL, "more than 2^31-1 reads");
    ctx->have_min = ctx->have_q = false;
    ctx->h_nmin_valid = false;
    ctx->n_reads = n_reads;
    ctx->h_off.assign(offsets, offsets + n_reads + 1);
    if (ctx->h_off[0] != 0) return fail(ctx, NGSID_EINVAL, "offsets[0] must be 0");
    ctx->h_woff.resize(n_reads + 1);
    int64_t wsum = 0;
    int maxlen = 0;
    for (int64_t i = 0; i < n_reads; ++i) {
        int64_t L = offsets[i + 1] - offsets[i];
        if (L < 0 || L > (1 << 24)) return fail(ctx, NGSID_EINVAL, "bad read length");
        ctx->h_woff[i] = wsum;
        wsum += (((L + 15) / 16 + 1) + 3) / 4 * 4;   // >= one spare word; reads start 16-B aligned
        maxlen = std::max<int>(maxlen, (int)L);
    }
    ctx->h_woff[n_reads] = wsum;
    ctx->total_bases = offsets[n_reads];
    ctx->total_words = wsum;
    ctx->max_len = maxlen;
    if (n_reads == 0) return NGSID_OK;
    size_t nb = (size_t)ctx->total_bases;
    CUDA_TRY(ctx, ctx->d_seq.ensure(nb + 64));
    CUDA_TRY(ctx, ctx->d_qual.ensure(nb + 64));
    CUDA_TRY(ctx, ctx->d_off.ensure((n_reads + 1) * sizeof(int64_t)));
    CUDA_TRY(ctx, ctx->d_woff.ensure((n_reads + 1) * sizeof(int64_t)));
    CUDA_TRY(ctx, ctx->d_packed.ensure((size_t)wsum * 4 + 64));
    CUDA_TRY(ctx, ctx->d_flag.ensure(64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_off.p, ctx->h_off.data(), (n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_woff.p, ctx->h_woff.data(), (n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    return NGSID_OK;
}

// 2-bit packing of the bases now in d_seq. Reads with a base outside ACGT are remembered: their
// minimizers come from the exception path (k1_exceptions.cuh), everything else treats their bases as
// the raw characters the reference compares.
static int reads_finish(ngsid_ctx *ctx)
{
    const int64_t n_reads = ctx->n_reads;
    if (n_reads == 0) return NGSID_OK;
    CUDA_TRY(ctx, ctx->d_rflag.ensure((size_t)n_reads + 64));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flag.p, 0, 64, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_rflag.p, 0, (size_t)n_reads, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_packed.p, 0, (size_t)ctx->total_words * 4 + 64, ctx->stream));
    int blocks = (int)std::min<int64_t>((n_reads + 7) / 8, (int64_t)ctx->sm_count * 16);
    cudaEventRecord(ctx->pev[0][0], ctx->stream);
    k_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_seq.as<uint8_t>(), ctx->d_off.as<int64_t>(),
                                                   ctx->d_woff.as<int64_t>(), ctx->d_packed.as<uint32_t>(),
                                                   n_reads, ctx->d_flag.as<int>(), ctx->d_rflag.as<uint8_t>());
    cudaEventRecord(ctx->pev[0][1], ctx->stream);
    ctx->pev_valid[0] = true;
    KERNEL_CHECK(ctx);
    int flag = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->x_reads.clear();
    ctx->xkmers.clear();
    ctx->xkmer_id.clear();
    if (flag) {
        std::vector<uint8_t> rf((size_t)n_reads);
        CUDA_TRY(ctx, cudaMemcpyAsync(rf.data(), ctx->d_rflag.p, (size_t)n_reads, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t r = 0; r < n_reads; ++r) if (rf[r]) ctx->x_reads.push_back((int32_t)r);
    }
    return NGSID_OK;
}

extern "C" int ngsid_upload_reads(ngsid_ctx *ctx, const uint8_t *seq, const uint8_t *qual,
                                  const int64_t *offsets, int64_t n_reads)
{
    if (!ctx || !offsets || n_reads < 0 || (n_reads > 0 && (!seq || !qual))) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = reads_layout(ctx, offsets, n_reads);
    if (rc || n_reads == 0) return rc;
    const size_t nb = (size_t)ctx->total_bases;
    // In pieces of 4 MB with at most two of them queued: a copy engine serves its queue in submission order, and
    // another context on this GPU (a clustering pass running under this upload, multi_gpu.Pipeline.prefetch) has
    // small latency-critical copies to put into the same queue -- behind 150 MB they would wait 3 ms each.
    const size_t piece = (size_t)4 << 20;
    const uint8_t *src[2] = {seq, qual};
    uint8_t *dst[2] = {ctx->d_seq.as<uint8_t>(), ctx->d_qual.as<uint8_t>()};
    int q = 0;
    for (int a = 0; a < 2; ++a)
        for (size_t o = 0; o < nb; o += piece, ++q) {
            const size_t len = std::min(piece, nb - o);
            if (q >= 2) CUDA_TRY(ctx, cudaEventSynchronize(ctx->up_ev[q & 1]));
            CUDA_TRY(ctx, cudaMemcpyAsync(dst[a] + o, src[a] + o, len, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaEventRecord(ctx->up_ev[q & 1], ctx->stream));
        }
    return reads_finish(ctx);
}

// ================================================================================ ingest (host)
extern "C" int ngsid_fastq_parse(const uint8_t *buf, int64_t len, int64_t cap_records,
                                 uint8_t *seq_out, uint8_t *qual_out, int64_t *name_off, int32_t *name_len,
                                 int64_t *seq_off, int64_t *qual_off, uint8_t *has_qual, int64_t *n_records)
{
    if (len < 0 || (len > 0 && !buf) || cap_records < 0 || !n_records) return NGSID_EINVAL;
    const bool wr = seq_out != nullptr;
    if (wr && (!qual_out || !name_off || !name_len || !seq_off || !qual_off || !has_qual)) return NGSID_EINVAL;
    fastq_ingest::Out O = {cap_records, 0, seq_out, qual_out, name_off, name_len, seq_off, qual_off, has_qual, 0, 0};
    *n_records = fastq_ingest::parse(buf, len, O);
    return (wr && *n_records > cap_records) ? NGSID_EINVAL : NGSID_OK;
}

// ================================================================================ K1
static int k1_prepare(ngsid_ctx *ctx, int k, int w)
{
    if (k < 2 || k > 15) return fail(ctx, NGSID_EUNSUPPORTED, "k must be in 2..15 in this build");
    if (w < k || w > 100) return fail(ctx, NGSID_EINVAL, "need k <= w <= 100");
    if (ctx->have_min && ctx->k == k && ctx->w == w) return NGSID_OK;
    int64_t n = ctx->n_reads;
    ctx->h_moff.resize(n + 1);
    int64_t s = 0;
    for (int64_t i = 0; i < n; ++i) {
        ctx->h_moff[i] = s;
        int64_t L = ctx->h_off[i + 1] - ctx->h_off[i];
        s += (std::max<int64_t>(1, L - w + 1) + 7) / 8 * 8;        // rows of 8 records stay 64-byte aligned
    }
    ctx->h_moff[n] = s;
    CUDA_TRY(ctx, ctx->d_moff.ensure((n + 1) * sizeof(int64_t)));
    CUDA_TRY(ctx, ctx->d_mins.ensure((size_t)s * sizeof(Minimizer) + 64));
    CUDA_TRY(ctx, ctx->d_nmin.ensure((n + 1) * sizeof(uint32_t)));
    CUDA_TRY(ctx, ctx->d_lenc.ensure((n + 1) * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_moff.p, ctx->h_moff.data(), (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->k = k; ctx->w = w;
    ctx->have_min = false;
    ctx->h_nmin_valid = false;
    return NGSID_OK;
}

static int k1_launch_generic(ngsid_ctx *ctx, const int32_t *list, const int32_t *list_n, int64_t n_work)
{
    int lcap = ((ctx->max_len + 31) / 32) * 32 + 32;
    while (k1_smem_per_warp(lcap) % 16 != 0) lcap += 4;
    size_t per_warp = k1_smem_per_warp(lcap);
    int wpb = (int)std::min<size_t>(8, SMEM_BUDGET / per_warp);
    if (wpb < 1) return fail(ctx, NGSID_EUNSUPPORTED, "read too long for the K1 shared-memory tile");
    size_t smem = per_warp * wpb;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k1_minimizers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k1_minimizers_kernel, wpb * 32, smem));
    bps = std::max(1, bps);
    int64_t need = (n_work + wpb - 1) / wpb;
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)ctx->sm_count * bps));
    k1_minimizers_kernel<<<blocks, wpb * 32, smem, ctx->stream>>>(
        ctx->d_packed.as<uint32_t>(), ctx->d_woff.as<int64_t>(), ctx->d_off.as<int64_t>(),
        ctx->d_moff.as<int64_t>(), ctx->d_mins.as<Minimizer>(), ctx->d_nmin.as<uint32_t>(),
        ctx->d_lenc.as<uint32_t>(), ctx->n_reads, ctx->k, ctx->w, lcap, list, list_n);
    KERNEL_CHECK(ctx);
    return NGSID_OK;
}

static int k1_launch(ngsid_ctx *ctx)
{
    // stream kernel (thread per read for compression + window minima, warp per 32 reads for the output) for
    // windows of 8 k-mers, k <= 13, as long as its per-thread regions fit; the generic warp-per-read kernel
    // for every other (k, w) and for the reads the stream kernel hands over (short, or poorly compressing)
    const K1SGeom g = k1s_geometry(ctx->max_len, ctx->k);
    const size_t smem = k1s_smem_bytes(g);
    const bool fast = (ctx->w - ctx->k + 1 == 8) && ctx->k <= 13 && ctx->k >= 2 && ctx->k1_variant != 0 &&
                      smem <= K1S_SMEM_LIMIT;
    if (!fast) return k1_launch_generic(ctx, nullptr, nullptr, ctx->n_reads);
    CUDA_TRY(ctx, ctx->d_newslots.ensure((size_t)(ctx->n_reads + 16) * 4));
    int32_t *slow_n = ctx->d_newslots.as<int32_t>();
    int32_t *slow_list = slow_n + 4;
    CUDA_TRY(ctx, cudaMemsetAsync(slow_n, 0, 16, ctx->stream));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k1_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (int)((ctx->n_reads + K1S_THREADS - 1) / K1S_THREADS);
    k1_stream_kernel<<<blocks, K1S_THREADS, smem, ctx->stream>>>(
        ctx->d_packed.as<uint32_t>(), ctx->d_woff.as<int64_t>(), ctx->d_off.as<int64_t>(),
        ctx->d_moff.as<int64_t>(), ctx->d_mins.as<Minimizer>(), ctx->d_nmin.as<uint32_t>(),
        ctx->d_lenc.as<uint32_t>(), ctx->n_reads, ctx->k, ctx->w, g.sw, g.bw, g.rs, g.scap, slow_list, slow_n);
    KERNEL_CHECK(ctx);
    return k1_launch_generic(ctx, slow_list, slow_n, std::min<int64_t>(ctx->n_reads, 4096));
}

#include "k1_exceptions.cuh"

extern "C" int ngsid_minimizers(ngsid_ctx *ctx, int k, int w)
{
    if (!ctx) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = k1_prepare(ctx, k, w);
    if (rc) return rc;
    if (ctx->n_reads == 0) { ctx->have_min = true; return NGSID_OK; }
    cudaEventRecord(ctx->pev[1][0], ctx->stream);
    rc = k1_launch(ctx);
    cudaEventRecord(ctx->pev[1][1], ctx->stream);
    ctx->pev_valid[1] = true;
    if (rc) return rc;
    rc = k1_exceptions(ctx);
    if (rc) return rc;
    ctx->have_min = true;
    ctx->h_nmin_valid = false;
    return NGSID_OK;
}

extern "C" int ngsid_minimizers_timed(ngsid_ctx *ctx, int k, int w, int iters, float *avg_ms)
{
    if (!ctx || iters < 1 || !avg_ms) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = k1_prepare(ctx, k, w);
    if (rc) return rc;
    if (ctx->n_reads == 0) { *avg_ms = 0.f; return NGSID_OK; }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; ++i) {
        rc = k1_launch(ctx);
        if (rc) return rc;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    rc = k1_exceptions(ctx);
    if (rc) return rc;
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *avg_ms = ms / iters;
    ctx->have_min = true;
    ctx->h_nmin_valid = false;
    return NGSID_OK;
}

static int fetch_nmin(ngsid_ctx *ctx)
{
    if (ctx->h_nmin_valid) return NGSID_OK;
    ctx->h_nmin.resize(ctx->n_reads);
    if (ctx->n_reads) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_nmin.data(), ctx->d_nmin.p, ctx->n_reads * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->h_nmin_valid = true;
    return NGSID_OK;
}

extern "C" int ngsid_get_minimizers(ngsid_ctx *ctx, int64_t begin, int64_t end, uint32_t *len_c,
                                    uint32_t *counts, uint32_t *kmer, uint32_t *pos, int64_t cap,
                                    int64_t *n_total)
{
    if (!ctx || begin < 0 || end < begin || end > ctx->n_reads) return NGSID_EINVAL;
    if (!ctx->have_min) return fail(ctx, NGSID_ESTATE, "ngsid_minimizers has not run");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = fetch_nmin(ctx);
    if (rc) return rc;
    int64_t n = end - begin, total = 0;
    for (int64_t i = begin; i < end; ++i) total += ctx->h_nmin[i];
    if (n_total) *n_total = total;
    if (counts) for (int64_t i = 0; i < n; ++i) counts[i] = ctx->h_nmin[begin + i];
    if (len_c && n) {
        CUDA_TRY(ctx, cudaMemcpyAsync(len_c, ctx->d_lenc.as<uint32_t>() + begin, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (!kmer && !pos) return NGSID_OK;
    if (total > cap) return fail(ctx, NGSID_EINVAL, "minimizer output buffer too small");
    if (n == 0) return NGSID_OK;
    // the slack-CSR region [moff[begin], moff[end]) comes back in one copy, compaction on the host
    int64_t lo = ctx->h_moff[begin], hi = ctx->h_moff[end];
    std::vector<Minimizer> tmp((size_t)(hi - lo));
    CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), ctx->d_mins.as<Minimizer>() + lo, (size_t)(hi - lo) * sizeof(Minimizer), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t o = 0;
    for (int64_t i = begin; i < end; ++i) {
        const Minimizer *m = tmp.data() + (ctx->h_moff[i] - lo);
        for (uint32_t j = 0; j < ctx->h_nmin[i]; ++j, ++o) {
            if (kmer) kmer[o] = m[j].x;
            if (pos) pos[o] = m[j].y;
        }
    }
    return NGSID_OK;
}

// ================================================================================ K0
extern "C" int ngsid_quality_stats(ngsid_ctx *ctx, const double *phred_p, const double *bucket_thresholds)
{
    if (!ctx || !phred_p || !bucket_thresholds) return NGSID_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t n = ctx->n_reads;
    CUDA_TRY(ctx, ctx->d_phred.ensure(128 * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_thr.ensure(16 * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_errc.ensure((n + 1) * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_erru.ensure((n + 1) * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_bucket.ensure(n + 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_phred.p, phred_p, 128 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_thr.p, bucket_thresholds, 14 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 8);
        cudaEventRecord(ctx->pev[2][0], ctx->stream);
        k0_quality_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_seq.as<uint8_t>(), ctx->d_qual.as<uint8_t>(),
                                                           ctx->d_off.as<int64_t>(), ctx->d_phred.as<double>(),
                                                           ctx->d_thr.as<double>(), ctx->d_errc.as<double>(),
                                                           ctx->d_erru.as<double>(), ctx->d_bucket.as<uint8_t>(), n);
        cudaEventRecord(ctx->pev[2][1], ctx->stream);
        ctx->pev_valid[2] = true;
        KERNEL_CHECK(ctx);
    }
    ctx->have_q = true;
    return NGSID_OK;
}

// ================================================================================ sort stage (row f.1)
extern "C" int ngsid_sort_scores(ngsid_ctx *ctx, int k, const double *phred_p_capped, const double *phred_p_uncapped,
                                 double *out_score, double *out_err_rate)
{
    if (!ctx || !phred_p_capped || !phred_p_uncapped || !out_score || !out_err_rate) return NGSID_EINVAL;
    if (k < 1 || k > 64) return fail(ctx, NGSID_EINVAL, "k out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->n_reads;
    if (n == 0) return NGSID_OK;
    CUDA_TRY(ctx, ctx->d_ss_tab.ensure(2 * 128 * sizeof(double) + 16 * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_ss_score.ensure((n + 1) * sizeof(double)));
    CUDA_TRY(ctx, ctx->d_ss_err.ensure(2 * (n + 1) * sizeof(double) + n + 64));
    double *tab = ctx->d_ss_tab.as<double>();
    double thr[14];
    for (int t = 0; t < 14; ++t) thr[t] = 1e300;          // buckets are not used here
    CUDA_TRY(ctx, cudaMemcpyAsync(tab, phred_p_capped, 128 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(tab + 128, phred_p_uncapped, 128 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(tab + 256, thr, sizeof thr, cudaMemcpyHostToDevice, ctx->stream));
    k0s_sortscore_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->d_qual.as<uint8_t>(), ctx->d_off.as<int64_t>(), tab, k, ctx->d_ss_score.as<double>(), n);
    KERNEL_CHECK(ctx);
    // mean error probability over the raw qualities with the uncapped table
    // (get_sorted_fastq_for_cluster.py:145-146): the histogram + compensated sum of K0
    double *erru = ctx->d_ss_err.as<double>(), *errc = erru + (n + 1);
    uint8_t *bucket = reinterpret_cast<uint8_t *>(errc + (n + 1));
    int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 8);
    k0_quality_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_seq.as<uint8_t>(), ctx->d_qual.as<uint8_t>(),
                                                       ctx->d_off.as<int64_t>(), tab + 128, tab + 256, errc, erru, bucket, n);
    KERNEL_CHECK(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(out_score, ctx->d_ss_score.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_err_rate, erru, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

extern "C" int ngsid_get_quality_stats(ngsid_ctx *ctx, int64_t begin, int64_t end, double *err_compressed,
                                       double *err_raw, uint8_t *bucket)
{
    if (!ctx || begin < 0 || end < begin || end > ctx->n_reads) return NGSID_EINVAL;
    if (!ctx->have_q) return fail(ctx, NGSID_ESTATE, "ngsid_quality_stats has not run");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t n = end - begin;
    if (n == 0) return NGSID_OK;
    if (err_compressed) CUDA_TRY(ctx, cudaMemcpyAsync(err_compressed, ctx->d_errc.as<double>() + begin, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (err_raw) CUDA_TRY(ctx, cudaMemcpyAsync(err_raw, ctx->d_erru.as<double>() + begin, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (bucket) CUDA_TRY(ctx, cudaMemcpyAsync(bucket, ctx->d_bucket.as<uint8_t>() + begin, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

// ================================================================================ K4 launcher
static int k4_launch(ngsid_ctx *ctx, const int32_t *pa, const int32_t *pb, const int32_t *po,
                     const int32_t *pm, int stride, int64_t n_pairs, int k, int32_t *out_count,
                     int32_t *out_score)
{
    if (n_pairs == 0) return NGSID_OK;
    int n2cap = ((ctx->max_len + 15) / 16) * 16 + 16;
    size_t per_warp = k4_smem_per_warp(n2cap);
    int wpb = (int)std::min<size_t>(4, SMEM_BUDGET / per_warp);
    if (wpb < 1) return fail(ctx, NGSID_EUNSUPPORTED, "read too long for the K4 shared-memory row buffer");
    size_t smem = per_warp * wpb;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k4_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k4_align_kernel, wpb * 32, smem));
    bps = std::max(1, bps);
    int64_t need = (n_pairs + wpb - 1) / wpb;
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)ctx->sm_count * bps));
    k4_align_kernel<<<blocks, wpb * 32, smem, ctx->stream>>>(ctx->d_seq.as<uint8_t>(), ctx->d_off.as<int64_t>(),
                                                             pa, pb, po, pm, stride, n_pairs, k, n2cap,
                                                             out_count, out_score);
    KERNEL_CHECK(ctx);
    return NGSID_OK;
}


// ---- K4 trace variant: DP with packed trace + traceback, processed in chunks of trace slots ----
static int k4t_run(ngsid_ctx *ctx, const int32_t *pa, const int32_t *pb, const int32_t *po, const int32_t *pm,
                   int stride, int64_t n_pairs, int k, int max_n1, int max_n2, bool have_aux,
                   int32_t *out_count, int32_t *out_score, int32_t *out_match, int32_t *out_cols,
                   K4TWindow *out_win, int window)
{
    if (n_pairs == 0) return NGSID_OK;
    int n2cap = ((max_n2 + 15) / 16) * 16 + 16;
    // throughput shape: one warp per pair
    size_t per_warp = k4t_smem_per_warp(n2cap);
    int wpb = (int)std::min<size_t>(4, SMEM_BUDGET / per_warp);
    if (wpb < 1) return fail(ctx, NGSID_EUNSUPPORTED, "sequence too long for the K4 shared-memory row buffer");
    size_t smem = per_warp * wpb;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k4t_dp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k4t_dp_kernel<false>, wpb * 32, smem));
    bps = std::max(1, bps);
    // latency shape: one block of W warps per pair (small rounds of the greedy pass)
    const int npass_max = (max_n1 + 32 * K4T_RPL - 1) / (32 * K4T_RPL);
    const int W = std::max(1, std::min(K4T_MAXW, npass_max));
    const size_t smem_multi = k4t_smem_multi(n2cap, W);
    const bool multi_ok = W > 1 && smem_multi <= SMEM_BUDGET;
    int bps_multi = 1;
    if (multi_ok) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k4t_dp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_multi));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_multi, k4t_dp_kernel<true>, W * 32, smem_multi));
        bps_multi = std::max(1, bps_multi);
    }
    // traceback: one warp per pair, shared-memory row ring + both sequences
    const int ncap = ((std::max(max_n1, max_n2) + 15) / 16) * 16 + 16;
    const size_t tb_per_warp = k4t_tb_smem_per_warp(ncap);
    int tb_wpb = (int)std::min<size_t>(4, SMEM_BUDGET / tb_per_warp);
    if (tb_wpb < 1) return fail(ctx, NGSID_EUNSUPPORTED, "sequence too long for the K4 traceback buffers");
    const size_t tb_smem = tb_per_warp * tb_wpb;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k4t_traceback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem));

    const size_t slot_words = k4t_trace_words(max_n1, max_n2);
    // trace arena: 4 GB, 16 GB for bulk launches (the thread-per-pair traceback is latency bound:
    // one launch over 50 k pairs costs what one over 13 k does)
    static const int64_t tb_thread_min = getenv("NGSID_K4_TB_THREAD_MIN") ? atoll(getenv("NGSID_K4_TB_THREAD_MIN")) : 12288;
    const bool thread_tb_ok = out_win == nullptr && ctx->k4_tb != 1;
    const size_t budget = (thread_tb_ok && n_pairs > 2 * tb_thread_min) ? (size_t)16 << 30 : (size_t)4 << 30;
    int64_t slots = (int64_t)std::max<size_t>(1, budget / (slot_words * 4));
    slots = std::min<int64_t>(slots, n_pairs);
    CUDA_TRY(ctx, ctx->d_trace.ensure((size_t)slots * slot_words * 4));
    CUDA_TRY(ctx, ctx->d_ends.ensure((size_t)slots * sizeof(K4TEnd)));
    K4TSeqs Q = {ctx->d_seq.as<uint8_t>(), ctx->d_off.as<int64_t>(),
                 have_aux ? ctx->d_auxseq.as<uint8_t>() : nullptr, have_aux ? ctx->d_aoff.as<int64_t>() : nullptr};
    for (int64_t p0 = 0; p0 < n_pairs; p0 += slots) {
        const int64_t c = std::min<int64_t>(slots, n_pairs - p0);
        // fewer pairs than the one-warp-per-pair shape has resident warps: take the latency shape (a lone
        // warp needs 0.83 ms for a 750 x 750 pair, a block of strips 0.29 ms; A/B on the bench: +5 %)
        static const double multi_factor = getenv("NGSID_K4_MULTI_FACTOR") ? atof(getenv("NGSID_K4_MULTI_FACTOR")) : 1.0;
        const bool multi = multi_ok && ctx->k4_shape != 1 &&
                           (ctx->k4_shape == 2 || (double)c * multi_factor <= (double)((int64_t)ctx->sm_count * bps * wpb / W));
        if (multi) {
            int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(c, (int64_t)ctx->sm_count * bps_multi));
            k4t_dp_kernel<true><<<blocks, W * 32, smem_multi, ctx->stream>>>(Q, pa, pb, po, stride, p0, c, n2cap,
                                                                             ctx->d_trace.as<uint32_t>(), slot_words,
                                                                             ctx->d_ends.as<K4TEnd>());
        } else {
            int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((c + wpb - 1) / wpb, (int64_t)ctx->sm_count * bps));
            k4t_dp_kernel<false><<<blocks, wpb * 32, smem, ctx->stream>>>(Q, pa, pb, po, stride, p0, c, n2cap,
                                                                          ctx->d_trace.as<uint32_t>(), slot_words,
                                                                          ctx->d_ends.as<K4TEnd>());
        }
        KERNEL_CHECK(ctx);
        if (thread_tb_ok && (ctx->k4_tb == 2 || c >= tb_thread_min))
            k4t_traceback_thread_kernel<<<(unsigned)((c + 127) / 128), 128, 0, ctx->stream>>>(
                Q, pa, pb, pm, stride, p0, c, k, ctx->d_trace.as<uint32_t>(), slot_words, ctx->d_ends.as<K4TEnd>(),
                out_count, out_score, out_match, out_cols);
        else
            k4t_traceback_kernel<<<(unsigned)((c + tb_wpb - 1) / tb_wpb), tb_wpb * 32, tb_smem, ctx->stream>>>(
                Q, pa, pb, pm, stride, p0, c, k, ctx->d_trace.as<uint32_t>(), slot_words, ctx->d_ends.as<K4TEnd>(),
                out_count, out_score, out_match, out_cols, out_win, window, ncap);
        KERNEL_CHECK(ctx);
    }
    return NGSID_OK;
}

static int k4_dispatch(ngsid_ctx *ctx, const int32_t *pa, const int32_t *pb, const int32_t *po, const int32_t *pm,
                       int stride, int64_t n_pairs, int k, int32_t *out_count, int32_t *out_score)
{
    if (ctx->use_payload_k4 && ctx->x_reads.empty()) return k4_launch(ctx, pa, pb, po, pm, stride, n_pairs, k, out_count, out_score);
    return k4t_run(ctx, pa, pb, po, pm, stride, n_pairs, k, ctx->max_len, ctx->max_len, false,
                   out_count, out_score, nullptr, nullptr, nullptr, 500);
}

extern "C" int ngsid_sg_block_align(ngsid_ctx *ctx, const int32_t *read_a, const int32_t *read_b,
                                    const int32_t *open, const int32_t *match_id, int64_t n_pairs, int k,
                                    int32_t *out_count, int32_t *out_score)
{
    if (!ctx || n_pairs < 0 || (n_pairs > 0 && (!read_a || !read_b || !open || !match_id || !out_count)))
        return NGSID_EINVAL;
    if (k < 1 || k > 15) return fail(ctx, NGSID_EUNSUPPORTED, "k must be in 1..15 in this build");
    if (n_pairs == 0) return NGSID_OK;
    for (int64_t i = 0; i < n_pairs; ++i) {
        if (read_a[i] < 0 || read_a[i] >= ctx->n_reads || read_b[i] < 0 || read_b[i] >= ctx->n_reads)
            return fail(ctx, NGSID_EINVAL, "pair index out of range");
        int64_t n1 = ctx->h_off[read_a[i] + 1] - ctx->h_off[read_a[i]];
        int64_t n2 = ctx->h_off[read_b[i] + 1] - ctx->h_off[read_b[i]];
        if (n1 < 1 || n2 < 1) return fail(ctx, NGSID_EINVAL, "empty sequence in alignment pair");
        if (n1 + n2 > 65000) return fail(ctx, NGSID_EUNSUPPORTED, "alignment pair too long for the 16-bit window counter");
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t nb = (size_t)n_pairs * sizeof(int32_t);
    CUDA_TRY(ctx, ctx->d_pa.ensure(nb)); CUDA_TRY(ctx, ctx->d_pb.ensure(nb));
    CUDA_TRY(ctx, ctx->d_po.ensure(nb)); CUDA_TRY(ctx, ctx->d_pm.ensure(nb));
    CUDA_TRY(ctx, ctx->d_k4cnt.ensure(nb)); CUDA_TRY(ctx, ctx->d_k4score.ensure(nb));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pa.p, read_a, nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pb.p, read_b, nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_po.p, open, nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pm.p, match_id, nb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = k4_dispatch(ctx, ctx->d_pa.as<int32_t>(), ctx->d_pb.as<int32_t>(), ctx->d_po.as<int32_t>(),
                         ctx->d_pm.as<int32_t>(), 1, n_pairs, k, ctx->d_k4cnt.as<int32_t>(), ctx->d_k4score.as<int32_t>());
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(out_count, ctx->d_k4cnt.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_score) CUDA_TRY(ctx, cudaMemcpyAsync(out_score, ctx->d_k4score.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}


static int upload_aux(ngsid_ctx *ctx, const uint8_t *aux_seq, const int64_t *aux_off, int64_t n_aux)
{
    if (n_aux <= 0) return NGSID_OK;
    if (!aux_seq || !aux_off || aux_off[0] != 0) return fail(ctx, NGSID_EINVAL, "bad auxiliary sequence arena");
    CUDA_TRY(ctx, ctx->d_auxseq.ensure((size_t)aux_off[n_aux] + 64));
    CUDA_TRY(ctx, ctx->d_aoff.ensure((size_t)(n_aux + 1) * sizeof(int64_t)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_auxseq.p, aux_seq, (size_t)aux_off[n_aux], cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_aoff.p, aux_off, (size_t)(n_aux + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    return NGSID_OK;
}

extern "C" int ngsid_sg_align_paths(ngsid_ctx *ctx, const int32_t *a, const int32_t *b, const int32_t *open,
                                    int64_t n_pairs, const uint8_t *aux_seq, const int64_t *aux_off, int64_t n_aux,
                                    int window, int32_t *out_score, int32_t *out_match, int32_t *out_cols,
                                    int32_t *out_win)
{
    if (!ctx || n_pairs < 0 || (n_pairs > 0 && (!a || !b || !open)) || window < 1) return NGSID_EINVAL;
    if (n_pairs == 0) return NGSID_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int max1 = 1, max2 = 1;
    auto seqlen = [&](int r, int64_t &L) -> bool {
        if (r >= 0) { if (r >= ctx->n_reads) return false; L = ctx->h_off[r + 1] - ctx->h_off[r]; }
        else { int64_t x = -(int64_t)r - 1; if (x >= n_aux) return false; L = aux_off[x + 1] - aux_off[x]; }
        return true;
    };
    for (int64_t i = 0; i < n_pairs; ++i) {
        int64_t l1 = 0, l2 = 0;
        if (!seqlen(a[i], l1) || !seqlen(b[i], l2)) return fail(ctx, NGSID_EINVAL, "pair index out of range");
        if (l1 < 1 || l2 < 1) return fail(ctx, NGSID_EINVAL, "empty sequence in alignment pair");
        if (out_win && (l2 + window - 1) / window > K4T_MAXWIN) return fail(ctx, NGSID_EUNSUPPORTED, "target longer than 16 windows");
        max1 = std::max<int>(max1, (int)l1); max2 = std::max<int>(max2, (int)l2);
    }
    int rc = upload_aux(ctx, aux_seq, aux_off, n_aux);
    if (rc) return rc;
    size_t nb = (size_t)n_pairs * sizeof(int32_t);
    CUDA_TRY(ctx, ctx->d_pa.ensure(nb)); CUDA_TRY(ctx, ctx->d_pb.ensure(nb)); CUDA_TRY(ctx, ctx->d_po.ensure(nb));
    CUDA_TRY(ctx, ctx->d_k4score.ensure(nb)); CUDA_TRY(ctx, ctx->d_match.ensure(nb)); CUDA_TRY(ctx, ctx->d_cols.ensure(nb));
    if (out_win) CUDA_TRY(ctx, ctx->d_win.ensure((size_t)n_pairs * K4T_MAXWIN * sizeof(K4TWindow)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pa.p, a, nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pb.p, b, nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_po.p, open, nb, cudaMemcpyHostToDevice, ctx->stream));
    rc = k4t_run(ctx, ctx->d_pa.as<int32_t>(), ctx->d_pb.as<int32_t>(), ctx->d_po.as<int32_t>(), nullptr, 1, n_pairs,
                 13, max1, max2, n_aux > 0, nullptr, ctx->d_k4score.as<int32_t>(), ctx->d_match.as<int32_t>(),
                 ctx->d_cols.as<int32_t>(), out_win ? ctx->d_win.as<K4TWindow>() : nullptr, window);
    if (rc) return rc;
    if (out_score) CUDA_TRY(ctx, cudaMemcpyAsync(out_score, ctx->d_k4score.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_match) CUDA_TRY(ctx, cudaMemcpyAsync(out_match, ctx->d_match.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_cols) CUDA_TRY(ctx, cudaMemcpyAsync(out_cols, ctx->d_cols.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_win) CUDA_TRY(ctx, cudaMemcpyAsync(out_win, ctx->d_win.p, (size_t)n_pairs * K4T_MAXWIN * sizeof(K4TWindow), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}


// ================================================================================ K5
#include "k5_host.cuh"

// ================================================================================ clustering driver
namespace {

// pair of events around a launch; resolved (summed per kind) at the end of the pass
static void ev_begin(ngsid_ctx *ctx, int kind)
{
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        ctx->ev_pool.push_back(a); ctx->ev_pool.push_back(b);
    }
    if (ctx->ev_kind.size() < ctx->ev_pool.size() / 2) ctx->ev_kind.resize(ctx->ev_pool.size() / 2);
    ctx->ev_kind[ctx->ev_used / 2] = kind;
    cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream);
}
static void ev_end(ngsid_ctx *ctx)
{
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) return;
    cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], ctx->stream);
    ctx->ev_used += 2;
}
static void ev_resolve(ngsid_ctx *ctx)
{
    ctx->phase_acc[4] = ctx->phase_acc[5] = 0.f;
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess)
            ctx->phase_acc[ctx->ev_kind[i / 2]] += ms;
    }
    ctx->pev_valid[4] = ctx->pev_valid[5] = true;
    ctx->ev_used = 0;
}

struct ClusterRun {
    ngsid_ctx *ctx;
    int64_t n;                         // entries of `order`
    std::vector<int32_t> slot_read, slot_pos;
    std::vector<uint8_t> slot_state;
    int n_slots = 0, slots_on_device = 0;
    uint32_t table_cap = 0;
    DevBuf keys_alt, heads_alt;        // spare pair for growth
    int64_t pairs_inserted = 0;
    int scap = 0, map_warps = 0, map_blocks = 0;
    int aovf_cap = 0;                  // nodes of the alignment-result overflow pool
    int slot_cap = 0;                  // entries of d_slot_read / d_slot_pos / d_slot_state
    int64_t node_cap = 0;              // posting nodes of d_nodes
    DevBuf d_list2, d_err, d_spec_u, d_spec_mat;
    std::vector<int32_t> h_dec, h_spec_u;
    ngsid_cluster_stats st;
    MapArgs A;

    int ensure_table(int64_t pairs_after);
    int ensure_scratch(int slots_after);
    int ensure_slot_node_capacity(int64_t pairs_after);
    int push_slots();
    int set_state(int slot, uint8_t v);
    int insert_slots(int slot0, int count);
    int run_map(const int32_t *d_list, int n_list);
    int fetch_dec(int lo, int hi);
    int prefetch_alignments(const std::vector<int32_t> &U, int lo, int hi);
};

int ClusterRun::ensure_table(int64_t pairs_after)
{
    if (table_cap && pairs_after * 2 <= (int64_t)table_cap) return NGSID_OK;
    uint32_t ncap = 1u << 16;
    while ((int64_t)ncap < pairs_after * 4) ncap <<= 1;
    DevBuf &nk = table_cap ? keys_alt : ctx->d_keys;
    DevBuf &nh = table_cap ? heads_alt : ctx->d_heads;
    CUDA_TRY(ctx, nk.ensure((size_t)ncap * 4));
    CUDA_TRY(ctx, nh.ensure((size_t)ncap * 4));
    CUDA_TRY(ctx, cudaMemsetAsync(nk.p, 0xff, (size_t)ncap * 4, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(nh.p, 0xff, (size_t)ncap * 4, ctx->stream));
    if (table_cap) {
        MapTable nt = {nk.as<uint32_t>(), nh.as<int32_t>(), A.table.nodes, ncap - 1};
        k2_rehash_kernel<<<(table_cap + 255) / 256, 256, 0, ctx->stream>>>(A.table.keys, A.table.heads, table_cap, nt);
        KERNEL_CHECK(ctx);
        std::swap(ctx->d_keys, keys_alt);
        std::swap(ctx->d_heads, heads_alt);
    }
    table_cap = ncap;
    A.table.keys = ctx->d_keys.as<uint32_t>();
    A.table.heads = ctx->d_heads.as<int32_t>();
    A.table.cap_mask = ncap - 1;
    return NGSID_OK;
}

int ClusterRun::ensure_scratch(int slots_after)
{
    if (slots_after <= scap) return NGSID_OK;
    int ncap = 1024;
    while (ncap < slots_after * 2) ncap <<= 1;
    // warps in flight bounded by a 6 GB scratch budget
    int64_t budget = (int64_t)6 << 30;
    int64_t max_warps = budget / ((int64_t)ncap * 12);
    int bps = 5;                                                     // blocks of 8 warps per SM (48 registers: 5 fit)
    if (const char *e = getenv("NGSID_MAP_BPS")) bps = std::max(1, std::min(8, atoi(e)));
    int blocks = (int)std::min<int64_t>((int64_t)ctx->sm_count * bps, std::max<int64_t>(1, max_warps / 8));
    map_blocks = blocks;
    map_warps = blocks * 8;
    size_t bytes = (size_t)map_warps * 12 * ncap;
    CUDA_TRY(ctx, ctx->d_scratch.ensure(bytes));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scratch.p, 0, bytes, ctx->stream));
    scap = ncap;
    A.scratch = ctx->d_scratch.as<uint32_t>();
    A.scap = scap;
    return NGSID_OK;
}

// Slots of tentative representatives that a surprise invalidates are never reused, and the reads
// behind them are inserted again under new slots, so neither the slot arrays nor the posting nodes
// are bounded by the number of reads: both grow here, contents kept.
int ClusterRun::ensure_slot_node_capacity(int64_t pairs_after)
{
    if (n_slots > slot_cap) {
        const int ncap = std::max(n_slots + 16, slot_cap * 2);
        CUDA_TRY(ctx, ctx->d_slot_read.grow_keep((size_t)ncap * 4, (size_t)slots_on_device * 4, ctx->stream));
        CUDA_TRY(ctx, ctx->d_slot_pos.grow_keep((size_t)ncap * 4, (size_t)slots_on_device * 4, ctx->stream));
        CUDA_TRY(ctx, ctx->d_slot_state.grow_keep((size_t)ncap, (size_t)slots_on_device, ctx->stream));
        slot_cap = ncap;
        A.slot_read = ctx->d_slot_read.as<int32_t>();
        A.slot_pos = ctx->d_slot_pos.as<int32_t>();
        A.slot_state = ctx->d_slot_state.as<uint8_t>();
    }
    if (pairs_after + 1 > node_cap) {
        const int64_t ncap = std::max<int64_t>(pairs_after + 1024, node_cap * 2);
        CUDA_TRY(ctx, ctx->d_nodes.grow_keep((size_t)ncap * sizeof(PostingNode), (size_t)pairs_inserted * sizeof(PostingNode), ctx->stream));
        node_cap = ncap;
        A.table.nodes = ctx->d_nodes.as<PostingNode>();
    }
    return NGSID_OK;
}

int ClusterRun::push_slots()
{
    if (slots_on_device == n_slots) return NGSID_OK;
    int a = slots_on_device, c = n_slots - a;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_slot_read.as<int32_t>() + a, slot_read.data() + a, c * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_slot_pos.as<int32_t>() + a, slot_pos.data() + a, c * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_slot_state.as<uint8_t>() + a, slot_state.data() + a, c, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors may reallocate later
    slots_on_device = n_slots;
    A.n_slots = n_slots;
    return NGSID_OK;
}

int ClusterRun::set_state(int slot, uint8_t v)
{
    slot_state[slot] = v;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_slot_state.as<uint8_t>() + slot, &slot_state[slot], 1, cudaMemcpyHostToDevice, ctx->stream));
    return NGSID_OK;
}

int ClusterRun::insert_slots(int slot0, int count)
{
    if (count <= 0) return NGSID_OK;
    int64_t add = 0;
    for (int s = slot0; s < slot0 + count; ++s) add += ctx->h_nmin[slot_read[s]];
    int rc = ensure_table(pairs_inserted + add);
    if (rc) return rc;
    rc = ensure_scratch(n_slots);
    if (rc) return rc;
    rc = ensure_slot_node_capacity(pairs_inserted + add);
    if (rc) return rc;
    rc = push_slots();
    if (rc) return rc;
    k2_insert_kernel<<<(count + 7) / 8, 256, 0, ctx->stream>>>(A.table, ctx->d_cursor.as<int32_t>(), (int32_t)std::min<int64_t>(node_cap, 0x7fffffff),
                                                               d_err.as<int32_t>(), ctx->d_slot_read.as<int32_t>(),
                                                               slot0, count, ctx->d_mins.as<Minimizer>(),
                                                               ctx->d_moff.as<int64_t>(), ctx->d_nmin.as<uint32_t>());
    KERNEL_CHECK(ctx);
    pairs_inserted += add;
    return NGSID_OK;
}

int ClusterRun::run_map(const int32_t *d_list, int n_list)
{
    if (n_list <= 0) return NGSID_OK;
    const int32_t *list = d_list;
    int nl = n_list;
    while (true) {
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_reqn.p, 0, 4, ctx->stream));
        A.list = list; A.n_list = nl;
        int blocks = std::min(map_blocks, (nl + 7) / 8);
        ev_begin(ctx, 5);
        k2_map_kernel<<<blocks, 256, 0, ctx->stream>>>(A);
        ev_end(ctx);
        KERNEL_CHECK(ctx);
        st.n_map_launch_reads += nl;
        int32_t hdr[2] = {0, 0};
        CUDA_TRY(ctx, cudaMemcpyAsync(&hdr[0], ctx->d_reqn.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(&hdr[1], d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (hdr[1] == 2) return fail(ctx, NGSID_ENOMEM, "alignment result pool exhausted (more failed tied candidates than reads in the pass)");
        if (hdr[1]) return fail(ctx, NGSID_ECUDA, "posting node pool overrun (internal error)");
        int nreq = hdr[0];
        if (nreq == 0) return NGSID_OK;
        const AlignRequest *rq = ctx->d_req.as<AlignRequest>();
        ev_begin(ctx, 4);
        int rc = k4_dispatch(ctx, &rq->read_a, &rq->read_b, &rq->open, &rq->match_id, 6, nreq, ctx->k,
                             ctx->d_k4cnt.as<int32_t>(), nullptr);
        ev_end(ctx);
        if (rc) return rc;
        st.n_alignments += nreq;
        k2_apply_align_kernel<<<(nreq + 255) / 256, 256, 0, ctx->stream>>>(rq, nreq, ctx->d_k4cnt.as<int32_t>(),
                                                                           ctx->d_off.as<int64_t>(), A.params, A.acache,
                                                                           ctx->d_aovf_head.as<int32_t>(), ctx->d_aovf.as<AlignCacheOvf>(),
                                                                           d_err.as<int32_t>() + 4, aovf_cap, A.acache_inline,
                                                                           d_list2.as<int32_t>(), d_err.as<int32_t>());
        KERNEL_CHECK(ctx);
        list = d_list2.as<int32_t>();
        nl = nreq;
    }
}

// Alignment prefetch for the tentative representatives U[1..] of the tile [lo, hi). Their
// resolution is a chain (one batch per new representative) and every batch would otherwise pay
// the latency of its own small K4 launches. The alignment statistic depends on the pair only, so
// one launch aligns every (tentative read, earlier slot) pair that some batch could ask for
// (k2_map_kernel, spec_mode) and the batches find the results in the prefetch table.
int ClusterRun::prefetch_alignments(const std::vector<int32_t> &U, int lo, int hi)
{
    const int rows = (int)U.size() - 1, cols = n_slots;
    h_spec_u.assign((size_t)(hi - lo), -1);
    for (int u = 1; u <= rows; ++u) h_spec_u[U[u] - lo] = u - 1;
    CUDA_TRY(ctx, d_spec_u.ensure((size_t)(hi - lo) * 4 + 64));
    CUDA_TRY(ctx, d_spec_mat.ensure((size_t)rows * cols + 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_spec_u.p, h_spec_u.data(), (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_spec_mat.p, 0xff, (size_t)rows * cols, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_list.p, U.data() + 1, (size_t)rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_reqn.p, 0, 4, ctx->stream));
    A.spec_u = d_spec_u.as<int32_t>();
    A.spec_mat = d_spec_mat.as<int8_t>();
    A.spec_lo = lo; A.spec_hi = hi; A.spec_cols = cols;
    A.list = ctx->d_list.as<int32_t>(); A.n_list = rows;
    A.spec_mode = 1;
    ev_begin(ctx, 5);
    k2_map_kernel<<<std::min(map_blocks, (rows + 7) / 8), 256, 0, ctx->stream>>>(A);
    ev_end(ctx);
    A.spec_mode = 0;
    KERNEL_CHECK(ctx);
    int32_t nreq = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&nreq, ctx->d_reqn.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nreq == 0) return NGSID_OK;
    const AlignRequest *rq = ctx->d_req.as<AlignRequest>();
    ev_begin(ctx, 4);
    int rc = k4_dispatch(ctx, &rq->read_a, &rq->read_b, &rq->open, &rq->match_id, 6, nreq, ctx->k,
                         ctx->d_k4cnt.as<int32_t>(), nullptr);
    ev_end(ctx);
    if (rc) return rc;
    st.n_alignments += nreq;
    k2_apply_spec_kernel<<<(nreq + 255) / 256, 256, 0, ctx->stream>>>(rq, nreq, ctx->d_k4cnt.as<int32_t>(),
                                                                      ctx->d_off.as<int64_t>(), A.params, A.spec_u,
                                                                      A.spec_lo, A.spec_cols, A.spec_mat, d_err.as<int32_t>());
    KERNEL_CHECK(ctx);
    return NGSID_OK;
}

int ClusterRun::fetch_dec(int lo, int hi)
{
    if (hi <= lo) return NGSID_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dec.data() + lo, ctx->d_dec.as<int32_t>() + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NGSID_OK;
}

}  // namespace

extern "C" int ngsid_cluster(ngsid_ctx *ctx, const ngsid_cluster_params *params, const int32_t *order,
                             int64_t n_order, const int32_t *init_reps, int64_t n_init,
                             const uint32_t *acc_rank, int32_t *out_assign, uint8_t *out_via,
                             ngsid_cluster_stats *stats)
{
    if (!ctx || !params || n_order < 0 || (n_order > 0 && (!order || !out_assign)) || !acc_rank ||
        n_init < 0 || (n_init > 0 && !init_reps))
        return NGSID_EINVAL;
    if (!ctx->have_min || !ctx->have_q) return fail(ctx, NGSID_ESTATE, "run ngsid_minimizers and ngsid_quality_stats first");
    if (params->k != ctx->k || params->w != ctx->w) return fail(ctx, NGSID_ESTATE, "k/w differ from the extracted minimizers");
    if (n_order + n_init >= ((int64_t)1 << 30)) return fail(ctx, NGSID_EINVAL, "too many reads for one pass");
    for (int64_t i = 0; i < n_order; ++i)
        if (order[i] < 0 || order[i] >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "order index out of range");
    for (int64_t i = 0; i < n_init; ++i)
        if (init_reps[i] < 0 || init_reps[i] >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "init_reps index out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = fetch_nmin(ctx);
    if (rc) return rc;

    cudaEventRecord(ctx->pev[3][0], ctx->stream);
    ctx->pev_valid[3] = false;
    ctx->ev_used = 0;
    ClusterRun R;
    R.ctx = ctx;
    R.n = n_order;
    memset(&R.st, 0, sizeof(R.st));
    // the pass borrows its private device buffers from the context and hands them back on every way out
    struct ClusterLend {
        ClusterRun &r; ngsid_ctx *c;
        DevBuf *mine[6];
        ClusterLend(ClusterRun &r_, ngsid_ctx *c_) : r(r_), c(c_) {
            DevBuf *m[6] = {&r.keys_alt, &r.heads_alt, &r.d_list2, &r.d_err, &r.d_spec_u, &r.d_spec_mat};
            for (int i = 0; i < 6; ++i) { mine[i] = m[i]; std::swap(*mine[i], c->d_cl[i]); }
        }
        ~ClusterLend() { for (int i = 0; i < 6; ++i) std::swap(*mine[i], c->d_cl[i]); }
    } lend(R, ctx);
    const int n = (int)n_order;
    const int max_slots = (int)(n_init + n_order) + 1;
    int64_t max_pairs = 0;
    for (int64_t i = 0; i < n_init; ++i) max_pairs += ctx->h_nmin[init_reps[i]];
    for (int64_t i = 0; i < n_order; ++i) max_pairs += ctx->h_nmin[order[i]];
    if (max_pairs >= ((int64_t)1 << 31)) return fail(ctx, NGSID_EINVAL, "too many minimizers for one pass");

    DeviceClusterParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.k = params->k; hp.min_shared = params->min_shared; hp.symmetric = params->symmetric;
    hp.min_fraction = params->min_fraction; hp.mapped_threshold = params->mapped_threshold;
    hp.aligned_threshold = params->aligned_threshold;
    memcpy(hp.max_gap, params->max_gap, sizeof(hp.max_gap));

    CUDA_TRY(ctx, ctx->d_params.ensure(sizeof(hp)));
    CUDA_TRY(ctx, ctx->d_order.ensure((size_t)n * 4 + 64));
    CUDA_TRY(ctx, ctx->d_accrank.ensure((size_t)ctx->n_reads * 4 + 64));
    CUDA_TRY(ctx, ctx->d_dec.ensure((size_t)n * 4 + 64));
    CUDA_TRY(ctx, ctx->d_via.ensure((size_t)n + 64));
    CUDA_TRY(ctx, ctx->d_list.ensure((size_t)n * 4 + 64));
    CUDA_TRY(ctx, R.d_list2.ensure((size_t)n * 4 + 64));
    CUDA_TRY(ctx, R.d_err.ensure(64));
    CUDA_TRY(ctx, ctx->d_req.ensure((size_t)n * sizeof(AlignRequest) + 64));
    CUDA_TRY(ctx, ctx->d_reqn.ensure(64));
    CUDA_TRY(ctx, ctx->d_k4cnt.ensure((size_t)n * 4 + 64));
    CUDA_TRY(ctx, ctx->d_acache.ensure((size_t)n * ACACHE_N * sizeof(AlignCacheEntry) + 64));
    R.aovf_cap = std::max(4096, n);
    CUDA_TRY(ctx, ctx->d_aovf.ensure((size_t)R.aovf_cap * sizeof(AlignCacheOvf)));
    CUDA_TRY(ctx, ctx->d_aovf_head.ensure((size_t)n * 4 + 64));
    // Representatives are a small fraction of the reads on amplicon data; the slot arrays and the
    // posting nodes start at an eighth of the upper bound without surprises and grow on demand
    // (NGSID_TEST_SMALL_CAPS: tests start them tiny so that every run exercises the growth path).
    const bool small_caps = getenv("NGSID_TEST_SMALL_CAPS") != nullptr;
    R.slot_cap = small_caps ? (int)n_init + 4 : (int)n_init + std::max(1024, max_slots / 8);
    R.node_cap = small_caps ? 64 : std::max<int64_t>(1 << 16, max_pairs / 8);
    for (int64_t i = 0; i < n_init; ++i) R.node_cap += ctx->h_nmin[init_reps[i]];
    CUDA_TRY(ctx, ctx->d_nodes.ensure((size_t)R.node_cap * sizeof(PostingNode)));
    CUDA_TRY(ctx, ctx->d_cursor.ensure(64));
    CUDA_TRY(ctx, ctx->d_slot_read.ensure((size_t)R.slot_cap * 4));
    CUDA_TRY(ctx, ctx->d_slot_pos.ensure((size_t)R.slot_cap * 4));
    CUDA_TRY(ctx, ctx->d_slot_state.ensure((size_t)R.slot_cap));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_params.p, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
    if (n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_order.p, order, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_accrank.p, acc_rank, (size_t)ctx->n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_cursor.p, 0, 64, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(R.d_err.p, 0, 64, ctx->stream));
    if (n) {
        int64_t ne = (int64_t)n * ACACHE_N;
        k2_fill_acache_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_acache.as<AlignCacheEntry>(), ctx->d_aovf_head.as<int32_t>(), ne);
        KERNEL_CHECK(ctx);
    }

    MapArgs &A = R.A;
    memset(&A, 0, sizeof(A));
    A.table.nodes = ctx->d_nodes.as<PostingNode>();
    A.params = ctx->d_params.as<DeviceClusterParams>();
    A.order = ctx->d_order.as<int32_t>();
    A.slot_read = ctx->d_slot_read.as<int32_t>();
    A.slot_pos = ctx->d_slot_pos.as<int32_t>();
    A.slot_state = ctx->d_slot_state.as<uint8_t>();
    A.mins = ctx->d_mins.as<Minimizer>();
    A.moff = ctx->d_moff.as<int64_t>();
    A.nmin = ctx->d_nmin.as<uint32_t>();
    A.lenc = ctx->d_lenc.as<uint32_t>();
    A.bucket = ctx->d_bucket.as<uint8_t>();
    A.erru = ctx->d_erru.as<double>();
    A.acc_rank = ctx->d_accrank.as<uint32_t>();
    A.acache = ctx->d_acache.as<AlignCacheEntry>();
    A.aovf_head = ctx->d_aovf_head.as<int32_t>();
    A.aovf = ctx->d_aovf.as<AlignCacheOvf>();
    A.acache_inline = ACACHE_N;
    if (const char *e = getenv("NGSID_TEST_ACACHE_INLINE")) A.acache_inline = std::max(1, std::min(ACACHE_N, atoi(e)));
    A.dec = ctx->d_dec.as<int32_t>();
    A.via = ctx->d_via.as<uint8_t>();
    A.req = ctx->d_req.as<AlignRequest>();
    A.req_n = ctx->d_reqn.as<int32_t>();
    A.err_flag = R.d_err.as<int32_t>();

    R.slot_read.reserve(R.slot_cap); R.slot_pos.reserve(R.slot_cap); R.slot_state.reserve(R.slot_cap);
    R.h_dec.assign(n, DEC_NEW);

    auto cleanup = [&](int code) { return code; };     // the borrowed buffers go back to the context (ClusterLend)
    const bool use_prefetch = getenv("NGSID_NO_PREFETCH") == nullptr;

    // ---- initial representatives (merge rounds of modules/parallelize.py:196-215)
    for (int64_t i = 0; i < n_init; ++i) {
        R.slot_read.push_back(init_reps[i]); R.slot_pos.push_back(-1); R.slot_state.push_back(SLOT_VALID);
    }
    R.n_slots = (int)n_init;
    rc = R.ensure_table(std::max<int64_t>(1, 0));
    if (rc) return cleanup(rc);
    rc = R.ensure_scratch(std::max(1, R.n_slots));
    if (rc) return cleanup(rc);
    rc = R.insert_slots(0, R.n_slots);
    if (rc) return cleanup(rc);

    // tiles grow 32 -> x8 -> ... -> 65536: every tile costs the latency of its own map / K4 / traceback
    // rounds (>= 0.6 ms), new representatives come early in a score-sorted amplicon set (A/B on the
    // bench: growth 2 / 4 / 8 / 16 = 2.41 / 2.60 / 2.68 / 2.52 M reads/s)
    const int tmax = params->tile_reads > 0 ? params->tile_reads : 65536;
    const int tgrow = getenv("NGSID_TILE_GROWTH") ? std::max(2, atoi(getenv("NGSID_TILE_GROWTH"))) : 8;
    const int tfirst = getenv("NGSID_TILE_FIRST") ? std::max(1, atoi(getenv("NGSID_TILE_FIRST"))) : 32;
    // a pass that starts from a table (a merge round of the --t N path, modules/parallelize.py:153-217) mostly
    // maps its reads onto that table: no reason to start with a small tile
    int T = std::min(n_init > 0 ? std::max(tfirst, 256) : tfirst, tmax);
    int pos = 0;
    std::vector<int32_t> U, h_list;
    while (pos < n) {
        const int hi = std::min(n, pos + T);
        R.st.n_tiles++;
        A.spec_u = nullptr;
        // ---- phase 1: speculative pass against the representatives known before the tile
        k2_iota_kernel<<<(hi - pos + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_list.as<int32_t>(), pos, hi - pos);
        KERNEL_CHECK(ctx);
        rc = R.run_map(ctx->d_list.as<int32_t>(), hi - pos);
        if (rc) return cleanup(rc);
        rc = R.fetch_dec(pos, hi);
        if (rc) return cleanup(rc);
        U.clear();
        for (int i = pos; i < hi; ++i) if (R.h_dec[i] == DEC_NEW) U.push_back(i);
        if (U.empty()) { pos = hi; T = (int)std::min<int64_t>((int64_t)T * tgrow, tmax); continue; }

        // ---- tentative representatives; the first one is certain
        const int slot0 = R.n_slots;
        for (size_t u = 0; u < U.size(); ++u) {
            R.slot_read.push_back(order[U[u]]); R.slot_pos.push_back(U[u]);
            R.slot_state.push_back(u == 0 ? SLOT_VALID : SLOT_TENTATIVE);
        }
        R.n_slots += (int)U.size();
        rc = R.insert_slots(slot0, (int)U.size());
        if (rc) return cleanup(rc);

        // ---- alignment prefetch for the chain below (request buffer: n entries)
        if (use_prefetch && U.size() >= 3 && (int64_t)(U.size() - 1) * R.n_slots <= (int64_t)n) {
            rc = R.prefetch_alignments(U, pos, hi);
            if (rc) return cleanup(rc);
        }

        // ---- phase 2: resolve the tentative ones in order against the certain ones before them.
        // All unresolved ones are mapped in one batch against the representatives that are certain
        // so far. The decision of the first of them is final; as long as a resolved one turns out
        // not to be a representative the certain set is unchanged, so the next decision of the batch
        // is final too. The batch ends at the first one that IS a representative (the certain set
        // grew: the rest is mapped again). Batches = new representatives in the tile, not tentatives.
        for (size_t u = 1; u < U.size();) {
            h_list.assign(U.begin() + u, U.end());
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_list.p, h_list.data(), h_list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            rc = R.run_map(ctx->d_list.as<int32_t>(), (int)h_list.size());
            if (rc) return cleanup(rc);
            rc = R.fetch_dec(U[u], U.back() + 1);
            if (rc) return cleanup(rc);
            R.st.n_chain_steps++;
            while (u < U.size()) {
                const bool is_new = R.h_dec[U[u]] == DEC_NEW;
                rc = R.set_state(slot0 + (int)u, is_new ? SLOT_VALID : SLOT_DEAD);
                if (rc) return cleanup(rc);
                ++u;
                if (is_new) break;
            }
        }

        // ---- phase 3: the other reads after the first new representative see the new ones
        const int first = U[0];
        h_list.clear();
        {
            size_t ui = 0;
            for (int i = first + 1; i < hi; ++i) {
                while (ui < U.size() && U[ui] < i) ++ui;
                if (ui < U.size() && U[ui] == i) continue;
                if (R.h_dec[i] == DEC_SKIP) continue;
                h_list.push_back(i);
            }
        }
        int surprise = -1;
        if (!h_list.empty()) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_list.p, h_list.data(), h_list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            rc = R.run_map(ctx->d_list.as<int32_t>(), (int)h_list.size());
            if (rc) return cleanup(rc);
            rc = R.fetch_dec(first + 1, hi);
            if (rc) return cleanup(rc);
            for (int i : h_list) if (R.h_dec[i] == DEC_NEW) { surprise = i; break; }
        }
        if (surprise < 0) { pos = hi; T = (int)std::min<int64_t>((int64_t)T * tgrow, tmax); continue; }

        // ---- a read that had been assigned became a representative: everything after it in the
        // tile is re-done with it in the table
        R.st.n_surprises++;
        for (size_t u = 0; u < U.size(); ++u)
            if (U[u] > surprise && R.slot_state[slot0 + u] != SLOT_DEAD) {
                rc = R.set_state(slot0 + (int)u, SLOT_DEAD);
                if (rc) return cleanup(rc);
            }
        const int s_slot = R.n_slots;
        R.slot_read.push_back(order[surprise]); R.slot_pos.push_back(surprise); R.slot_state.push_back(SLOT_VALID);
        R.n_slots++;
        rc = R.insert_slots(s_slot, 1);
        if (rc) return cleanup(rc);
        if (hi > surprise + 1) {
            int64_t ne = (int64_t)(hi - surprise - 1) * ACACHE_N;
            k2_fill_acache_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(
                ctx->d_acache.as<AlignCacheEntry>() + (size_t)(surprise + 1) * ACACHE_N, ctx->d_aovf_head.as<int32_t>() + surprise + 1, ne);
            KERNEL_CHECK(ctx);
        }
        pos = surprise + 1;
        T = std::max(32, T / 2);
    }

    cudaEventRecord(ctx->pev[3][1], ctx->stream);
    ctx->pev_valid[3] = true;
    // ---- results
    std::vector<uint8_t> h_via(n);
    unsigned long long cells = 0;
    if (n) {
        CUDA_TRY(ctx, cudaMemcpyAsync(h_via.data(), ctx->d_via.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(&cells, R.d_err.as<uint8_t>() + 8, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ev_resolve(ctx);
    R.st.align_cells = (int64_t)cells;
    for (int i = 0; i < n; ++i) {
        int d = R.h_dec[i];
        out_assign[i] = d;
        if (out_via) out_via[i] = (d >= 0) ? h_via[i] : 0;
        if (d == DEC_SKIP) continue;
        R.st.n_processed++;
        if (d == DEC_NEW) R.st.n_new_reps++;
        else if (h_via[i] == 1) R.st.n_mapped++;
        else R.st.n_aln_passed++;
    }
    if (stats) *stats = R.st;
    return cleanup(NGSID_OK);
}

// ---- stand-alone hit table (test / inspection entry): see include/ngsid.h
extern "C" int ngsid_hit_counts(ngsid_ctx *ctx, const int32_t *reps, int64_t n_reps, const int32_t *reads, int64_t n,
                                uint32_t *out_count, uint32_t *out_possum)
{
    if (!ctx || n_reps < 0 || n < 0 || (n_reps > 0 && !reps) || (n > 0 && (!reads || !out_count || !out_possum))) return NGSID_EINVAL;
    if (!ctx->have_min) return fail(ctx, NGSID_ESTATE, "ngsid_minimizers has not run");
    for (int64_t i = 0; i < n_reps; ++i) if (reps[i] < 0 || reps[i] >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "representative out of range");
    for (int64_t i = 0; i < n; ++i) if (reads[i] < 0 || reads[i] >= ctx->n_reads) return fail(ctx, NGSID_EINVAL, "read out of range");
    if (n == 0 || n_reps == 0) return NGSID_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = fetch_nmin(ctx);
    if (rc) return rc;
    int64_t pairs = 0;
    for (int64_t i = 0; i < n_reps; ++i) pairs += ctx->h_nmin[reps[i]];
    uint32_t cap = 1u << 12;
    while ((int64_t)cap < pairs * 4) cap <<= 1;
    DevBuf keys, heads, nodes, cursor, err, d_reps, d_reads, d_cnt, d_sum;
    auto done = [&](int code) { DevBuf *b[] = {&keys, &heads, &nodes, &cursor, &err, &d_reps, &d_reads, &d_cnt, &d_sum}; for (DevBuf *x : b) x->release(); return code; };
#define HC_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->err = cudaGetErrorString(_e); return done(NGSID_ECUDA); } } while (0)
    HC_TRY(keys.ensure((size_t)cap * 4)); HC_TRY(heads.ensure((size_t)cap * 4)); HC_TRY(nodes.ensure((size_t)(pairs + 1) * sizeof(PostingNode)));
    HC_TRY(cursor.ensure(64)); HC_TRY(err.ensure(64)); HC_TRY(d_reps.ensure((size_t)n_reps * 4)); HC_TRY(d_reads.ensure((size_t)n * 4));
    HC_TRY(d_cnt.ensure((size_t)n * n_reps * 4)); HC_TRY(d_sum.ensure((size_t)n * n_reps * 4));
    HC_TRY(cudaMemsetAsync(keys.p, 0xff, (size_t)cap * 4, ctx->stream)); HC_TRY(cudaMemsetAsync(heads.p, 0xff, (size_t)cap * 4, ctx->stream));
    HC_TRY(cudaMemsetAsync(cursor.p, 0, 64, ctx->stream)); HC_TRY(cudaMemsetAsync(err.p, 0, 64, ctx->stream));
    HC_TRY(cudaMemsetAsync(d_cnt.p, 0, (size_t)n * n_reps * 4, ctx->stream)); HC_TRY(cudaMemsetAsync(d_sum.p, 0, (size_t)n * n_reps * 4, ctx->stream));
    HC_TRY(cudaMemcpyAsync(d_reps.p, reps, (size_t)n_reps * 4, cudaMemcpyHostToDevice, ctx->stream));
    HC_TRY(cudaMemcpyAsync(d_reads.p, reads, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    MapTable t = {keys.as<uint32_t>(), heads.as<int32_t>(), nodes.as<PostingNode>(), cap - 1};
    k2_insert_kernel<<<(unsigned)((n_reps + 7) / 8), 256, 0, ctx->stream>>>(t, cursor.as<int32_t>(), (int32_t)(pairs + 1), err.as<int32_t>(), d_reps.as<int32_t>(),
                                                                        0, (int)n_reps, ctx->d_mins.as<Minimizer>(), ctx->d_moff.as<int64_t>(), ctx->d_nmin.as<uint32_t>());
    k2_hits_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(t, d_reads.as<int32_t>(), (int)n, (int)n_reps, ctx->d_mins.as<Minimizer>(), ctx->d_moff.as<int64_t>(),
                                                                     ctx->d_nmin.as<uint32_t>(), d_reps.as<int32_t>(), d_cnt.as<uint32_t>(), d_sum.as<uint32_t>());
    ctx->launches += 2;
    HC_TRY(cudaGetLastError());
    HC_TRY(cudaMemcpyAsync(out_count, d_cnt.p, (size_t)n * n_reps * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HC_TRY(cudaMemcpyAsync(out_possum, d_sum.p, (size_t)n * n_reps * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HC_TRY(cudaStreamSynchronize(ctx->stream));
#undef HC_TRY
    return done(NGSID_OK);
}

#include "nccl_plane.cuh"
