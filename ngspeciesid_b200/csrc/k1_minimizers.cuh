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
// K1 fast path (w - k + 1 == 8, k <= 13): one THREAD per read, all state in registers.
//   * 4 bases at a time go through a 1024-entry shared-memory table (previous base, packed byte)
//     -> (kept bases, count): homopolymer compression without a per-base branch;
//   * kept bases queue up in a 64-bit FIFO; every 8 of them form a block: rolling 2k-bit code per
//     slot, key = code << 4 | slot-in-block, and the sliding minimum over 8 k-mers is
//     min(suffix minimum of the previous block, prefix minimum of this block) (van Herk /
//     Gil-Werman), branch-free, leftmost on ties because the slot index sits in the low key bits;
//   * a minimizer is recorded whenever the window minimum changes; records go to a per-thread
//     ring in shared memory and leave for HBM in rows of 8 records = 64 bytes (two full sectors),
//     eight rows per warp-wide store, so DRAM sees whole sectors instead of scattered 8-byte stores;
//   * the loop is block-centric and warp-uniform: every iteration each lane refills its FIFO
//     (cheap, variable) and runs exactly one block step (expensive, always useful).
// Reads whose compressed length is < w (reference quirk, one truncated-window minimizer) are
// appended to `slow_list` and finished by k1_minimizers_kernel.
#define K1F_INF 0xffffffffu
#define K1F_THREADS 128
#define K1F_RING 16                 // records per thread ring
#define K1F_RSTRIDE 17              // ring stride in records (padding against bank conflicts)
#define K1F_ROW 8                   // records per flushed row (64 bytes = two full sectors)

struct K1FastState {
    uint32_t suf[9];
    uint32_t code;
    int base;          // slot index of the first slot of the next block
    uint32_t last_key; // key of the last recorded minimizer in the frame of the current block
    uint32_t n_out;
};

// One block of 8 compressed bases. Records (key, block base) pairs into the thread's ring.
template <bool PARTIAL>
__device__ __forceinline__ void k1_fast_block(K1FastState &S, uint32_t ch, int nvalid, int k, int w,
                                              uint32_t kmask, uint2 *__restrict__ ring)
{
    uint32_t key[8];
    uint32_t pm = K1F_INF;
    const bool warm = S.base + 7 < w - 1;          // no complete window ends in this block
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t b = (ch >> (14 - 2 * t)) & 3u;
        S.code = ((S.code << 2) | b) & kmask;
        const int slot = S.base + t;
        uint32_t kk = (S.code << 4) | (uint32_t)(8 + t);
        if (slot < k - 1) kk = K1F_INF;
        if (PARTIAL && t >= nvalid) kk = K1F_INF;
        key[t] = kk;
        pm = min(pm, kk);
        const uint32_t m = min(S.suf[t + 1], pm);
        bool emit = (m != S.last_key) && !warm && (slot >= w - 1);
        if (PARTIAL) emit = emit && (t < nvalid);
        if (emit) {
            ring[S.n_out & (K1F_RING - 1)] = make_uint2(m, (uint32_t)S.base);
            S.n_out++;
            S.last_key = m;
        }
    }
    uint32_t sm = K1F_INF;
#pragma unroll
    for (int t = 7; t >= 0; --t) {
        sm = min(sm, key[t]);
        S.suf[t] = sm - 8u;          // becomes "previous block": slot tags 0..7
    }
    // the same record seen from the next block's frame; a record from the previous block can no
    // longer be a window minimum, so it must not compare equal to anything
    S.last_key = (S.last_key & 8u) ? S.last_key - 8u : (K1F_INF - 1u);
    S.base += 8;
}

// Steady-state block (every slot >= w-1): no warm-up tests, k-mer codes by funnel shift out of
// (previous code : 8 new bases), and the record is stored unconditionally at ring[n_out] -- the
// index only advances when the window minimum changed, otherwise the next store overwrites it.
__device__ __forceinline__ void k1_fast_block_steady(K1FastState &S, uint32_t ch, uint32_t kmask,
                                                     uint2 *__restrict__ ring)
{
    const uint32_t lo = (S.code << 16) | ch, hi = S.code >> 16;
    uint32_t key[8];
    uint32_t pm = K1F_INF;
    const uint32_t base = (uint32_t)S.base;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t c = __funnelshift_r(lo, hi, 14 - 2 * t) & kmask;
        const uint32_t kk = (c << 4) | (uint32_t)(8 + t);
        key[t] = kk;
        pm = min(pm, kk);
        const uint32_t m = min(S.suf[t + 1], pm);
        ring[S.n_out & (K1F_RING - 1)] = make_uint2(m, base);
        S.n_out += (m != S.last_key) ? 1u : 0u;
        S.last_key = m;
    }
    S.code = lo & kmask;
    uint32_t sm = K1F_INF;
#pragma unroll
    for (int t = 7; t >= 0; --t) {
        sm = min(sm, key[t]);
        S.suf[t] = sm - 8u;
    }
    S.last_key = (S.last_key & 8u) ? S.last_key - 8u : (K1F_INF - 1u);
    S.base += 8;
}

// decode a ring record: key = code << 4 | tag, tag 0..7 = previous block, 8..15 = block at `base`
__device__ __forceinline__ Minimizer k1_fast_decode(uint2 rec, int k)
{
    const int slot = (int)rec.y - 8 + (int)(rec.x & 15u);
    return make_uint2(rec.x >> 4, (uint32_t)(slot - (k - 1)));
}

__global__ void __launch_bounds__(K1F_THREADS)
k1_fast_kernel(const uint32_t *__restrict__ packed, const int64_t *__restrict__ woff,
               const int64_t *__restrict__ off, const int64_t *__restrict__ moff,
               Minimizer *__restrict__ mins, uint32_t *__restrict__ nmin,
               uint32_t *__restrict__ lenc, int64_t n_reads, int k, int w,
               int32_t *__restrict__ slow_list, int32_t *__restrict__ slow_n)
{
    __shared__ uint16_t lut[1024];
    __shared__ uint2 rings[K1F_THREADS * K1F_RSTRIDE];
    for (int idx = threadIdx.x; idx < 1024; idx += blockDim.x) {
        uint32_t prev = idx >> 8, byte = idx & 255, bits = 0, cnt = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t b = (byte >> (6 - 2 * t)) & 3u;
            if (b != prev) { bits = (bits << 2) | b; cnt++; }
            prev = b;
        }
        lut[idx] = (uint16_t)((cnt << 8) | bits);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = r < n_reads;
    const int L = have ? (int)(off[r + 1] - off[r]) : 0;
    const uint2 *pk = reinterpret_cast<const uint2 *>(packed + (have ? woff[r] : 0));
    Minimizer *out = mins + (have ? moff[r] : 0);
    uint2 *ring = rings + threadIdx.x * K1F_RSTRIDE;
    const uint32_t kmask = (1u << (2 * k)) - 1u;

    K1FastState S;
#pragma unroll
    for (int t = 0; t < 9; ++t) S.suf[t] = K1F_INF;
    S.code = 0; S.base = 0; S.last_key = K1F_INF - 1u; S.n_out = 0;
    uint32_t fifo = 0;                          // <= 16 kept bases, newest in the low bits
    unsigned long long inbuf = 0;
    int avail = 0, inleft = 0;                  // bytes left in inbuf
    int bytes_left = L >> 2;                    // whole bytes (4 bases) of input still to consume
    int next_pair = 0;                          // next uint2 (8 bytes = 32 bases) to load
    const int n_pairs = (L + 31) >> 5;
    uint2 ahead = (have && n_pairs > 0) ? __ldg(pk) : make_uint2(0, 0);
    uint32_t prev = have ? ((ahead.x >> 30) ^ 1u) : 0u, n_fl = 0;
    bool tail_done = (L == 0);
    bool active = have;

#define K1F_LOAD_PAIR()                                                              \
    do {                                                                             \
        inbuf = ((unsigned long long)ahead.x << 32) | ahead.y;                       \
        inleft = 8;                                                                  \
        ++next_pair;                                                                 \
        if (next_pair < n_pairs) ahead = __ldg(pk + next_pair);                      \
    } while (0)
#define K1F_LUT_STEP()                                                               \
    do {                                                                             \
        if (inleft == 0) K1F_LOAD_PAIR();                                            \
        const uint32_t byte_ = (uint32_t)(inbuf >> 56);                              \
        inbuf <<= 8; --inleft; --bytes_left;                                         \
        const uint32_t e_ = lut[(prev << 8) | byte_];                                \
        const uint32_t c_ = e_ >> 8;                                                 \
        fifo = (fifo << (2 * c_)) | (e_ & 255u);                                     \
        avail += (int)c_;                                                            \
        prev = byte_ & 3u;                                                           \
    } while (0)

    while (__any_sync(NGSID_FULL_MASK, active)) {
        if (active) {
            // ---- refill: three straight-line table steps cover the average demand (8 kept bases per
            // block ~ 2.7 input bytes); the loop behind them only runs after long homopolymers
#pragma unroll
            for (int st = 0; st < 3; ++st)
                if (avail <= 12 && bytes_left > 0) K1F_LUT_STEP();
            while (avail < 8 && bytes_left > 0) K1F_LUT_STEP();
            if (avail < 8 && !tail_done) {
                const int rem = L & 3;          // the last 0..3 bases sit in the next byte
                if (rem) {
                    if (inleft == 0) K1F_LOAD_PAIR();
                    uint32_t byte = (uint32_t)(inbuf >> 56);
                    for (int t = 0; t < rem; ++t) {
                        const uint32_t b = (byte >> 6) & 3u;
                        byte <<= 2;
                        if (b != prev) { fifo = (fifo << 2) | b; avail++; }
                        prev = b;
                    }
                }
                tail_done = true;
            }
            if (avail >= 8) {
                const uint32_t ch = (fifo >> (2 * (avail - 8))) & 0xffffu;
                avail -= 8;
                if (S.base >= w) k1_fast_block_steady(S, ch, kmask, ring);
                else k1_fast_block<false>(S, ch, 8, k, w, kmask, ring);
            } else {
                if (avail > 0) {                // input exhausted: last partial block
                    const uint32_t ch = (fifo << (2 * (8 - avail))) & 0xffffu;
                    const int nv = avail;
                    avail = 0;
                    k1_fast_block<true>(S, ch, nv, k, w, kmask, ring);
                    S.base += nv - 8;            // base now equals the compressed length
                }
                active = false;
            }
        }
        // ---- flush full rows of 8 records; eight rows per round, 4 lanes x 16 bytes per row
        uint32_t fm = __ballot_sync(NGSID_FULL_MASK, have && (S.n_out - n_fl) >= K1F_ROW);
        while (fm) {
            // each group of 4 lanes serves the first of its own lanes that has a full row
            const uint32_t gm = (fm >> (lane & ~3)) & 15u;
            const int leader = gm ? ((lane & ~3) + __ffs(gm) - 1) : -1;
            const bool is_leader = (leader == lane);
            fm = __ballot_sync(NGSID_FULL_MASK, ((fm >> lane) & 1u) && !is_leader);
            const int src = leader >= 0 ? leader : 0;
            const uint32_t fl = __shfl_sync(NGSID_FULL_MASK, n_fl, src);
            const unsigned long long op = __shfl_sync(NGSID_FULL_MASK, (unsigned long long)out, src);
            if (leader >= 0) {
                const int e = (lane & 3) * 2;
                const uint2 *rp = rings + (threadIdx.x - lane + leader) * K1F_RSTRIDE + ((fl + e) & (K1F_RING - 1));
                const Minimizer a = k1_fast_decode(rp[0], k), b = k1_fast_decode(rp[1], k);
                reinterpret_cast<uint4 *>(reinterpret_cast<Minimizer *>(op) + fl)[lane & 3] = make_uint4(a.x, a.y, b.x, b.y);
            }
            if (is_leader) n_fl += K1F_ROW;
        }
    }
#undef K1F_LUT_STEP
#undef K1F_LOAD_PAIR
    if (!have) return;
    const int Lc = S.base;
    if (Lc < w) {
        // compressed read shorter than w (or than k): the generic kernel reproduces the quirk
        int i = atomicAdd(slow_n, 1);
        slow_list[i] = (int32_t)r;
        return;
    }
    // leftover records (< 16): written by the owning thread
    for (uint32_t i = n_fl; i < S.n_out; ++i) out[i] = k1_fast_decode(ring[i & (K1F_RING - 1)], k);
    nmin[r] = S.n_out;
    lenc[r] = (uint32_t)Lc;
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
