// K_pack + K1: 2-bit packing, homopolymer compression and (k,w) minimizer extraction.
// Reference semantics: modules/cluster.py:265 (compression), modules/cluster.py:16-39
// (get_kmer_minimizers). One warp per read.
#pragma once
#include "ngsid_internal.cuh"

// ---------------------------------------------------------------------------------------------
// K_pack: ASCII -> 2 bit/base, first base in the most significant bits of each u32 word.
// A,C,G,T -> 0,1,2,3 (ASCII order, so unsigned compare of codes == string compare). Any other
// byte raises the error flag (this build handles the ACGT alphabet only). The unused tail of a
// read's last word repeats the read's last base, so the tail never starts a new homopolymer run
// (k1_stream_kernel compresses whole words).
__global__ void k_pack_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ off,
                              const int64_t *__restrict__ woff, uint32_t *__restrict__ packed,
                              int64_t n_reads, int *__restrict__ bad_flag, uint8_t *__restrict__ read_flag)
{
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t lane = lane_id();
    for (int64_t r = warp; r < n_reads; r += nwarps) {
        const uint8_t *s = seq + off[r];
        int L = (int)(off[r + 1] - off[r]);
        uint32_t *out = packed + woff[r];
        int nw = (L + 15) >> 4;
        for (int wi = lane; wi < nw; wi += 32) {
            uint32_t word = 0;
            int base = wi << 4;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                int b = min(base + t, L - 1);
                uint32_t c = s[b];
                bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
                if (!ok) { *bad_flag = 1; read_flag[r] = 1; }
                uint32_t code = ((c >> 1) & 3u) ^ ((c >> 2) & 1u);
                word |= code << (30 - 2 * t);
            }
            out[wi] = word;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t k1_get_kmer(const uint32_t *cp, int p, int k)
{
    int wi = p >> 4;
    int sh = (p & 15) << 1;
    uint32_t x = __funnelshift_l(cp[wi + 1], cp[wi], sh);
    return x >> (32 - 2 * k);
}

// Per-warp shared memory: km[lcap] u32 | cp[lcap/16 + 4] u32 | cs[lcap + 64] bytes
__host__ __device__ inline size_t k1_smem_per_warp(int lcap)
{
    return (size_t)lcap * 4 + ((size_t)lcap / 16 + 4) * 4 + (size_t)lcap + 64;
}

__global__ void __launch_bounds__(256)
k1_minimizers_kernel(const uint32_t *__restrict__ packed, const int64_t *__restrict__ woff,
                     const int64_t *__restrict__ off, const int64_t *__restrict__ moff,
                     Minimizer *__restrict__ mins, uint32_t *__restrict__ nmin,
                     uint32_t *__restrict__ lenc, int64_t n_reads, int k, int w, int lcap,
                     const int32_t *__restrict__ list, const int32_t *__restrict__ list_n)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warps_per_block = blockDim.x >> 5;
    const int wid = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *base = smem_raw + (size_t)wid * k1_smem_per_warp(lcap);
    uint32_t *km = reinterpret_cast<uint32_t *>(base);
    uint32_t *cp = km + lcap;
    uint8_t *cs = reinterpret_cast<uint8_t *>(cp + (lcap / 16 + 4));
    const int W = w - k + 1;

    // optional work list (reads the fast kernel handed over): entries [0, *list_n)
    const int64_t n_work = list ? (int64_t)*list_n : n_reads;
    for (int64_t wi_ = (int64_t)blockIdx.x * warps_per_block + wid; wi_ < n_work;
         wi_ += (int64_t)gridDim.x * warps_per_block) {
        const int64_t r = list ? (int64_t)list[wi_] : wi_;
        const int L = (int)(off[r + 1] - off[r]);
        const uint32_t *pk = packed + woff[r];
        Minimizer *out = mins + moff[r];

        // ---- phase A: homopolymer compression into one byte per kept base
        int nc = 0;
        uint32_t carry = 0;
        for (int b0 = 0; b0 < L; b0 += 32) {
            int b = b0 + (int)lane;
            bool valid = b < L;
            uint32_t word = valid ? __ldg(pk + (b >> 4)) : 0u;
            uint32_t code = (word >> (30 - 2 * (b & 15))) & 3u;
            uint32_t prev = __shfl_up_sync(NGSID_FULL_MASK, code, 1);
            if (lane == 0) prev = carry;
            bool keep = valid && (b == 0 || code != prev);
            uint32_t mask = __ballot_sync(NGSID_FULL_MASK, keep);
            if (keep) cs[nc + __popc(mask & lt_mask)] = (uint8_t)code;
            nc += __popc(mask);
            carry = __shfl_sync(NGSID_FULL_MASK, code, 31);
        }
        const int Lc = nc;
        // zero the tail so that packing can read whole words
        for (int t = lane; t < 64; t += 32) cs[Lc + t] = 0;
        __syncwarp();

        // ---- pack the compressed bases 2 bit/base (first base most significant)
        const int nwc = (Lc + 15) >> 4;
        const uint32_t *cs32 = reinterpret_cast<const uint32_t *>(cs);
        for (int wi = lane; wi < nwc + 2; wi += 32) {
            uint32_t word = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t v = cs32[4 * wi + q];
                word |= ((v * 0x40100401u) >> 24) << (24 - 8 * q);
            }
            cp[wi] = word;
        }
        __syncwarp();

        uint32_t n_out = 0;
        if (Lc >= w) {
            // ---- k-mer codes of every position
            const int nk = Lc - k + 1;
            for (int p = lane; p < nk; p += 32) km[p] = k1_get_kmer(cp, p, k);
            __syncwarp();
            // ---- leftmost minimum of every window of W k-mers; emit on change of position
            const int nwin = nk - W + 1;
            int carry_pos = -1;
            for (int p0 = 0; p0 < nwin; p0 += 32) {
                int p = p0 + (int)lane;
                bool valid = p < nwin;
                uint32_t best = 0xffffffffu;
                int bpos = -2;
                if (valid) {
                    best = km[p];
                    bpos = p;
                    for (int q = p + 1; q < p + W; ++q) {
                        uint32_t v = km[q];
                        if (v < best) { best = v; bpos = q; }
                    }
                }
                int prev = __shfl_up_sync(NGSID_FULL_MASK, bpos, 1);
                if (lane == 0) prev = carry_pos;
                bool emit = valid && (bpos != prev);
                uint32_t mask = __ballot_sync(NGSID_FULL_MASK, emit);
                if (emit) out[n_out + __popc(mask & lt_mask)] = make_uint2(best, (uint32_t)bpos);
                n_out += __popc(mask);
                carry_pos = __shfl_sync(NGSID_FULL_MASK, bpos, 31);
            }
        } else if (Lc >= k) {
            // ---- reference quirk for inputs shorter than w: the first window holds truncated
            // (possibly empty) suffix strings; one minimizer results (cluster.py:18-21).
            if (lane == 0) {
                int best_i = 0;
                for (int i = 1; i < W; ++i) {
                    int ta = min(max(Lc - i, 0), k), tb = min(max(Lc - best_i, 0), k);
                    int n = min(ta, tb);
                    int cmp = 0;
                    for (int t = 0; t < n && cmp == 0; ++t)
                        cmp = (int)cs[i + t] - (int)cs[best_i + t];
                    if (cmp == 0) cmp = ta - tb;
                    if (cmp < 0) best_i = i;
                }
                int t = min(max(Lc - best_i, 0), k);
                uint32_t code = 0;
                for (int q = 0; q < t; ++q) code = (code << 2) | cs[best_i + q];
                if (t < k) code |= (1u << 31) | (1u << (2 * t));
                out[0] = make_uint2(code, (uint32_t)best_i);
            }
            n_out = 1;
        }
        if (lane == 0) {
            nmin[r] = n_out;
            lenc[r] = (uint32_t)Lc;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// K0: quality statistics. Reference: modules/cluster.py:273-292 (per homopolymer run keep the
// quality char of lowest error probability; error rate = sum(count(c)*p(c))/len over the
// compressed qualities) and cluster.py:185-188 (same sum over the raw qualities / len(seq)).
// One warp per read; each lane owns a chunk of bases and the runs that START inside it.
__global__ void __launch_bounds__(256)
k0_quality_kernel(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual,
                  const int64_t *__restrict__ off, const double *__restrict__ phred_p,
                  const double *__restrict__ thr, double *__restrict__ errc,
                  double *__restrict__ erru, uint8_t *__restrict__ bucket, int64_t n_reads)
{
    __shared__ uint32_t hist[8][2][128];
    __shared__ double ptab[128];
    const int wid = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ptab[i] = phred_p[i];
    __syncthreads();
    uint32_t *hc = hist[wid][0], *hu = hist[wid][1];
    const int warps_per_block = blockDim.x >> 5;

    for (int64_t r = (int64_t)blockIdx.x * warps_per_block + wid; r < n_reads;
         r += (int64_t)gridDim.x * warps_per_block) {
        const int64_t o = off[r];
        const int L = (int)(off[r + 1] - o);
        const uint8_t *s = seq + o, *q = qual + o;
        for (int i = lane; i < 128; i += 32) { hc[i] = 0; hu[i] = 0; }
        __syncwarp();
        const int chunk = (L + 31) >> 5;
        const int beg = lane * chunk, end = min(L, beg + chunk);
        int nruns = 0;
        int i = beg;
        // skip the tail of a run that started in an earlier chunk (its owner follows it)
        if (i > 0) {
            uint8_t pb = s[i - 1];
            while (i < end && s[i] == pb) { atomicAdd(&hu[q[i] & 127], 1u); ++i; }
        }
        while (i < end) {
            uint8_t b = s[i];
            uint8_t best = q[i] & 127;
            double bp = ptab[best];
            atomicAdd(&hu[best], 1u);
            int j = i + 1;
            while (j < L && s[j] == b) {
                uint8_t c = q[j] & 127;
                if (j < end) atomicAdd(&hu[c], 1u);
                double p = ptab[c];
                if (p < bp) { bp = p; best = c; }
                ++j;
            }
            atomicAdd(&hc[best], 1u);
            ++nruns;
            i = j;   // may pass `end`; the next lane skips that tail (but counts its raw bins)
        }
        // total number of runs = compressed length
        for (int d = 16; d > 0; d >>= 1) nruns += __shfl_xor_sync(NGSID_FULL_MASK, nruns, d);
        __syncwarp();
        if (lane < 2) {
            const uint32_t *h = lane == 0 ? hc : hu;
            // Neumaier-compensated sum in ascending character order: Python >= 3.12 sum()
            double f = 0.0, c = 0.0;
            bool started = false;
            for (int ch = 0; ch < 128; ++ch) {
                uint32_t n = h[ch];
                if (n == 0) continue;
                double x = __dmul_rn((double)n, ptab[ch]);
                if (!started) { f = x; started = true; continue; }
                double t = __dadd_rn(f, x);
                if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dadd_rn(f, -t), x));
                else c = __dadd_rn(c, __dadd_rn(__dadd_rn(x, -t), f));
                f = t;
            }
            if (c != 0.0 && isfinite(c)) f = __dadd_rn(f, c);
            if (lane == 0) {
                double e = (nruns > 0) ? __ddiv_rn(f, (double)nruns) : 0.0;
                errc[r] = e;
                int bk = 0;
                for (int t = 0; t < 14; ++t) bk += (e >= thr[t]) ? 1 : 0;
                bucket[r] = (uint8_t)bk;
            } else {
                erru[r] = (L > 0) ? __ddiv_rn(f, (double)L) : 0.0;
            }
        }
        __syncwarp();
    }
}
