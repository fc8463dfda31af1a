// K1 exception path: reads with a base outside ACGT (N, IUPAC codes, lower case).
// The reference compares raw characters everywhere (homopolymer compression modules/cluster.py:265,
// k-mer strings in get_kmer_minimizers :16-39, table keys :43-62), so such a read is legal input and
// its k-mers order by character code ('A' < 'C' < 'G' < 'N' < 'T' < 'a' ...). The kernels work on
// 2 bit / base; the handful of reads that do not fit are restated here on the host, exactly, and
// their minimizer records replace what the kernel wrote for them: a k-mer of ACGT only keeps its
// 2-bit code (it has to meet the same k-mer of other reads in the table), any other string gets the
// code 1 << 30 | index into the context's dictionary (equal strings, equal codes).
#pragma once

static uint32_t k1x_code(ngsid_ctx *ctx, const std::string &km, int k)
{
    bool plain = true;
    for (char c : km) plain = plain && (c == 'A' || c == 'C' || c == 'G' || c == 'T');
    if (plain && (int)km.size() == k) {
        uint32_t code = 0;
        for (char c : km) code = (code << 2) | (c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u);
        return code;
    }
    if (plain) {                                      // truncated suffix (compressed read shorter than w)
        uint32_t code = 1u;
        for (char c : km) code = (code << 2) | (c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u);
        return code | (1u << 31);
    }
    auto it = ctx->xkmer_id.find(km);
    if (it != ctx->xkmer_id.end()) return (1u << 30) | it->second;
    const uint32_t id = (uint32_t)ctx->xkmers.size();
    ctx->xkmers.push_back(km);
    ctx->xkmer_id.emplace(km, id);
    return (1u << 30) | id;
}

static int k1_exceptions(ngsid_ctx *ctx)
{
    if (ctx->x_reads.empty()) return NGSID_OK;
    const int k = ctx->k, W = ctx->w - ctx->k + 1;
    std::vector<uint8_t> raw;
    std::vector<Minimizer> recs;
    for (int32_t r : ctx->x_reads) {
        const int64_t a = ctx->h_off[r], L = ctx->h_off[r + 1] - a;
        raw.resize((size_t)L + 1);
        if (L) CUDA_TRY(ctx, cudaMemcpyAsync(raw.data(), ctx->d_seq.as<uint8_t>() + a, (size_t)L, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        std::string sc;                               // homopolymer-compressed
        for (int64_t i = 0; i < L; ++i) if (i == 0 || raw[i] != raw[i - 1]) sc.push_back((char)raw[i]);
        const int lc = (int)sc.size();
        recs.clear();
        if (lc >= k) {
            // get_kmer_minimizers (modules/cluster.py:16-39): leftmost lexicographic minimum of every window of
            // W k-mers, reported when its position changes; a read shorter than w compares truncated suffixes
            auto kmer = [&](int p) { return p < lc ? sc.substr((size_t)p, (size_t)k) : std::string(); };
            int best = 0;
            for (int p = 1; p < W; ++p) if (kmer(p) < kmer(best)) best = p;
            recs.push_back(make_uint2(k1x_code(ctx, kmer(best), k), (uint32_t)best));
            const int n_kmers = lc - k + 1;
            for (int right = W; right < n_kmers; ++right) {
                const int left = right - W + 1;
                if (best < left) {
                    best = left;
                    for (int p = left + 1; p <= right; ++p) if (kmer(p) < kmer(best)) best = p;
                    recs.push_back(make_uint2(k1x_code(ctx, kmer(best), k), (uint32_t)best));
                } else if (kmer(right) < kmer(best)) {
                    best = right;
                    recs.push_back(make_uint2(k1x_code(ctx, kmer(best), k), (uint32_t)best));
                }
            }
        }
        const int64_t cap = ctx->h_moff[r + 1] - ctx->h_moff[r];
        if ((int64_t)recs.size() > cap) return fail(ctx, NGSID_EUNSUPPORTED, "exception read has more minimizers than its slots");
        const uint32_t nm = (uint32_t)recs.size(), lcu = (uint32_t)lc;
        if (nm) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mins.as<Minimizer>() + ctx->h_moff[r], recs.data(), (size_t)nm * sizeof(Minimizer), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_nmin.as<uint32_t>() + r, &nm, 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lenc.as<uint32_t>() + r, &lcu, 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NGSID_OK;
}

// the string behind a minimizer code of the exception path (1 << 30 | index)
extern "C" int ngsid_kmer_string(ngsid_ctx *ctx, uint32_t code, char *out, int cap)
{
    if (!ctx || !out || cap < 1) return NGSID_EINVAL;
    if (!(code & (1u << 30)) || (code & (1u << 31))) return fail(ctx, NGSID_EINVAL, "not a dictionary code");
    const uint32_t id = code & ((1u << 30) - 1u);
    if (id >= ctx->xkmers.size()) return fail(ctx, NGSID_EINVAL, "dictionary code out of range");
    const std::string &s = ctx->xkmers[id];
    if ((int)s.size() + 1 > cap) return fail(ctx, NGSID_EINVAL, "buffer too small");
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
