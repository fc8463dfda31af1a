// K1 "stream" kernel (w - k + 1 == 8, k <= 13): homopolymer compression + (k,w) minimizers.
// Reference semantics: modules/cluster.py:265 (compression), modules/cluster.py:16-39
// (get_kmer_minimizers): leftmost lexicographic minimum of every window of 8 k-mers of the
// compressed read, reported whenever its position changes.
//
// The kernel is bound by integer instruction issue, not by HBM, so it is organised around the
// instruction count per base:
//   phase A  thread per read   packed 2-bit words -> homopolymer-compressed 2-bit stream in shared
//                              memory. Four bases per table lookup; one funnel shift appends the
//                              kept bases of a byte (the table entry carries its own shift count).
//   phase B  thread per read   streaming sliding-window minimum over the compressed stream, 32 k-mers
//                              per step: key = left-aligned code | 6-bit position tag, van Herk
//                              prefix/suffix minima with 3-input minimum instructions (14 VIMNMX3 per
//                              8 windows), and the result is ONE BIT per minimizer position in a
//                              per-read bitmap (a minimizer record is fully determined by its
//                              position), so no compare/branch/store happens per window.
//   phase C  warp per 32 reads bitmap words -> ordered minimizer positions (popcount + warp scan +
//                              leading-zero extraction) staged in shared memory, then records
//                              (code, position) leave for HBM as consecutive 8-byte stores per read.
// All loops run to warp-uniform trip counts; lanes past the end of their read compute garbage in
// their own shared-memory region, and a per-read fix-up recomputes the last 7 windows exactly.
// Reads whose compressed length is < w + 8 (including the reference's "shorter than w" quirk) are
// handed to the generic kernel through `slow_list`.
//
// The per-thread phases are plain functions over pointers so that tests/k1_stream_host.cpp can run
// them on the CPU (K1S_HOST) against a naive window scan.
#pragma once
#include <stdint.h>

#ifdef K1S_HOST
#define K1S_DEV static inline
static inline uint32_t k1s_fsl(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? ((hi << s) | (lo >> (32u - s))) : hi;
}
static inline uint32_t k1s_fsr(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
}
static inline uint32_t k1s_clz(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
static inline uint32_t k1s_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
#else
#define K1S_DEV __device__ __forceinline__
__device__ __forceinline__ uint32_t k1s_fsl(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
__device__ __forceinline__ uint32_t k1s_fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
__device__ __forceinline__ uint32_t k1s_clz(uint32_t x) { return (uint32_t)__clz((int)x); }
__device__ __forceinline__ uint32_t k1s_min(uint32_t a, uint32_t b) { return min(a, b); }
#endif

#define K1S_THREADS 128
#define K1S_GROUP 4                 // reads per extraction group in phase C
#define K1S_WARP_EXTRA 32           // per warp: read starts of a group (K1S_GROUP + 1 words), 16-byte multiple
#define K1S_SLOTS 9                 // straight-line extraction slots per bitmap word
#define K1S_IPL 3                   // bitmap words per lane and extraction pass
#define K1S_SMEM_LIMIT (160 * 1024)  // above this the read set goes to the ring / generic kernels

#ifdef K1S_HOST
K1S_DEV uint32_t k1s_min3(uint32_t a, uint32_t b, uint32_t c) { return k1s_min(k1s_min(a, b), c); }
#else
K1S_DEV uint32_t k1s_min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }   // one VIMNMX3
#endif
// one bit per position, most significant bit first: bit(tag) = 0x80000000 >> (tag mod 32)
K1S_DEV uint32_t k1s_bit(uint32_t key) { return k1s_fsr(0x80000000u, 0u, key); }

// Table entry for (previous base, next 4 bases): kept bases left-aligned at bit 31, and 2 * count
// in the low bits, so that   out = funnelshift_l(e, out, e)   appends them in one instruction.
#ifdef K1S_HOST
static inline
#else
__host__ __device__ inline
#endif
uint32_t k1s_lut_entry(uint32_t idx)
{
    uint32_t prev = idx >> 8, byte = idx & 255u, bits = 0, cnt = 0;
    for (int t = 0; t < 4; ++t) {
        const uint32_t b = (byte >> (6 - 2 * t)) & 3u;
        if (b != prev) { bits = (bits << 2) | b; ++cnt; }
        prev = b;
    }
    return (cnt ? (bits << (32 - 2 * cnt)) : 0u) | (2u * cnt);
}

// Geometry of one thread's shared-memory region (32-bit words): stream | bitmap, odd stride so that
// lanes touching the same word index hit 32 different banks.
struct K1SGeom {
    int n_it_max;      // steps of 32 k-mers the longest possible read needs
    int sw, bw, rs;    // stream words, bitmap words, region stride
    int scap;          // staged records per extraction group (8 bytes each, even); overflow -> generic kernel
};
static inline K1SGeom k1s_geometry(int max_len, int k)
{
    K1SGeom g;
    // Regions are sized for a compressed length of 85 % of the longest raw read (sequencing reads
    // compress to ~75 %); a read that compresses less than that is handed to the generic kernel.
    int nk = (max_len * 17 + 19) / 20 - k + 1;
    if (nk < 1) nk = 1;
    g.n_it_max = (nk + 31) / 32;
    g.sw = 2 * g.n_it_max + 6;
    g.bw = g.n_it_max + 2;
    g.rs = g.sw + g.bw;
    if ((g.rs & 1) == 0) g.rs++;
    // records of a group are staged in shared memory before they leave as one bulk store per read.
    // Expected density is ~0.22 minimizers per k-mer of the ACTUAL read (regions are sized for the longest
    // read at 85 % compression); 0.195 per k-mer slot of the region keeps four blocks per SM at 800 bases
    // and leaves ~3 sigma of headroom; a group that does not fit goes to the generic kernel.
    g.scap = ((K1S_GROUP * g.n_it_max * 32 * 39 / 200) / 2) * 2;
    if (g.scap < 128) g.scap = 128;
    return g;
}
static inline size_t k1s_smem_bytes(const K1SGeom &g)
{
    // table | regions | per warp: staging + read starts
    return 1024 * 4 + ((size_t)K1S_THREADS * g.rs * 4 + 15) / 16 * 16 + (size_t)(K1S_THREADS / 32) * ((size_t)g.scap * 8 + K1S_WARP_EXTRA);
}

// ---------------------------------------------------------------------------------------------
// Phase A state: 64-bit left-aligned bit buffer of kept bases not yet written to the stream.
struct K1SCompress {
    unsigned long long buf;
    uint32_t nb;       // valid bits in buf (< 32 between words)
    uint32_t prevw;    // previous raw word (its last base is the run the next byte continues)
    int wp;            // stream words written
};

K1S_DEV void k1s_compress_init(K1SCompress &C, uint32_t first_word)
{
    C.buf = 0; C.nb = 0; C.wp = 0;
    C.prevw = (first_word >> 30) ^ 1u;      // differs from the first base: the first base is kept
}

// One raw word (16 bases, first base in the most significant bits). `live` is false for words past
// the end of the read: they are replaced by a run of the previous base and compress to nothing.
K1S_DEV void k1s_compress_word(K1SCompress &C, const uint32_t *lut, uint32_t w, bool live, uint32_t *st, int cap)
{
    const uint32_t fill = (C.prevw & 3u) * 0x55555555u;
    w = live ? w : fill;
    const char *lb = reinterpret_cast<const char *>(lut);
    const uint32_t e0 = *reinterpret_cast<const uint32_t *>(lb + (k1s_fsl(w, C.prevw, 10) & 0xffcu));
    const uint32_t e1 = *reinterpret_cast<const uint32_t *>(lb + ((w >> 14) & 0xffcu));
    const uint32_t e2 = *reinterpret_cast<const uint32_t *>(lb + ((w >> 6) & 0xffcu));
    const uint32_t e3 = *reinterpret_cast<const uint32_t *>(lb + ((w << 2) & 0xffcu));
    uint32_t out = k1s_fsl(e0, 0u, e0);
    out = k1s_fsl(e1, out, e1);
    out = k1s_fsl(e2, out, e2);
    out = k1s_fsl(e3, out, e3);
    const uint32_t s = (e0 + e1 + e2 + e3) & 63u;           // bits appended (<= 32)
    const uint32_t sh = 64u - C.nb - s;                     // >= 1
    C.buf |= ((unsigned long long)out << (sh - 1u)) << 1;
    C.nb += s;
    if (C.nb >= 32u) {
        if (C.wp < cap) st[C.wp] = (uint32_t)(C.buf >> 32);
        ++C.wp;
        C.buf <<= 32;
        C.nb -= 32u;
    }
    C.prevw = w;
}

// Writes the partial last word and the look-ahead words phase B and the fix-up read.
// Returns the compressed length in bases, or -1 when the stream does not fit `cap` words.
K1S_DEV int k1s_compress_finish(K1SCompress &C, uint32_t *st, int cap)
{
    if (C.wp + 4 > cap) return -1;
    st[C.wp] = (uint32_t)(C.buf >> 32);
    st[C.wp + 1] = 0u;
    st[C.wp + 2] = 0u;
    st[C.wp + 3] = 0u;
    return (int)((32u * (uint32_t)C.wp + C.nb) >> 1);
}

// ---------------------------------------------------------------------------------------------
// Phase B. Keys: (k-mer left-aligned in the top 2k bits) | 6-bit tag. Inside a step the 8 k-mers of
// the previous block carry tags 0..7 and the 32 k-mers of the step tags 8..39, so unsigned
// comparison of keys is (k-mer, position) lexicographic and the low 5 bits of the winning key
// select the bitmap bit directly.
struct K1SWin {
    uint32_t s2, s4, s6;       // suffix minima of the previous block from slots 2, 4, 6
    uint32_t a1, a3, a5, a7;   // odd keys of the previous block
};

// One block of 8 k-mers starting at step slot T0 (0, 8, 16, 24); lo/hi are the two stream words
// the k-mers start in (hi first). Returns the bits of the 8 windows ending in this block.
template <int T0, bool FIRST>
K1S_DEV uint32_t k1s_block(K1SWin &S, uint32_t hi, uint32_t lo, uint32_t topmask)
{
    uint32_t c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = T0 + j;
        const uint32_t x = ((t & 15) == 0) ? hi : k1s_fsl(lo, hi, 2 * (t & 15));
        c[j] = (x & topmask) | (uint32_t)(8 + t);
    }
    const uint32_t p2 = k1s_min3(c[0], c[1], c[2]);
    const uint32_t p4 = k1s_min3(p2, c[3], c[4]);
    const uint32_t p6 = k1s_min3(p4, c[5], c[6]);
    const uint32_t m7 = k1s_min(p6, c[7]);
    uint32_t bits;
    if (FIRST) {
        bits = k1s_bit(m7);                       // only the window ending at k-mer 7 is complete
    } else {
        const uint32_t m0 = k1s_min3(S.a1, S.s2, c[0]);
        const uint32_t m1 = k1s_min3(S.s2, c[0], c[1]);
        const uint32_t m2 = k1s_min3(S.a3, S.s4, p2);
        const uint32_t m3 = k1s_min3(S.s4, p2, c[3]);
        const uint32_t m4 = k1s_min3(S.a5, S.s6, p4);
        const uint32_t m5 = k1s_min3(S.s6, p4, c[5]);
        const uint32_t m6 = k1s_min(S.a7, p6);
        bits = k1s_bit(m0) | k1s_bit(m1) | k1s_bit(m2);
        bits |= k1s_bit(m3) | k1s_bit(m4);
        bits |= k1s_bit(m5) | k1s_bit(m6);
        bits |= k1s_bit(m7);
    }
    S.s6 = k1s_min(c[6], c[7]);
    S.s4 = k1s_min3(c[4], c[5], S.s6);
    S.s2 = k1s_min3(c[2], c[3], S.s4);
    S.a1 = c[1]; S.a3 = c[3]; S.a5 = c[5]; S.a7 = c[7];
    return bits;
}

// One step: k-mers [32 it, 32 it + 32). c0, c1, c2 = stream words 2 it, 2 it + 1, 2 it + 2.
// acc0 holds the bits of block 0 (tags 0..15), acc1 those of blocks 1..3 (tags 9..39, where 32..39
// alias onto 0..7 -- unambiguous because block 1..3 windows never reach tags below 9).
template <bool FIRST>
K1S_DEV void k1s_step(K1SWin &S, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t topmask,
                      uint32_t &acc0, uint32_t &acc1)
{
    acc0 = k1s_block<0, FIRST>(S, c0, c1, topmask);
    acc1 = k1s_block<8, false>(S, c0, c1, topmask);
    acc1 |= k1s_block<16, false>(S, c1, c2, topmask);
    acc1 |= k1s_block<24, false>(S, c1, c2, topmask);
    // the last block becomes "previous block": tags 32..39 -> 0..7
    S.s2 -= 32u; S.s4 -= 32u; S.s6 -= 32u;
    S.a1 -= 32u; S.a3 -= 32u; S.a5 -= 32u; S.a7 -= 32u;
}

// Streams n_it steps over the thread's compressed stream `st` and writes bitmap words bm[0, n_it).
// Bitmap bit (31 - b) of word i <-> k-mer position 32 i + b.
K1S_DEV void k1s_window_pass(const uint32_t *st, uint32_t *bm, int n_it, uint32_t topmask)
{
    K1SWin S;
    S.s2 = S.s4 = S.s6 = S.a1 = S.a3 = S.a5 = S.a7 = 0xffffffffu;
    uint32_t c0 = st[0], c1 = st[1], c2 = st[2];
    uint32_t acc0, acc1;
    k1s_step<true>(S, c0, c1, c2, topmask, acc0, acc1);
    uint32_t pending = k1s_fsl(acc1, acc0 | acc1, 8);
    for (int it = 1; it < n_it; ++it) {
        c0 = c2;
        c1 = st[2 * it + 1];
        c2 = st[2 * it + 2];
        k1s_step<false>(S, c0, c1, c2, topmask, acc0, acc1);
        bm[it - 1] = pending | (acc0 >> 24);              // tags 0..7: last 8 positions of the previous word
        pending = k1s_fsl(acc1, acc0 | acc1, 8);          // tags 8..31 | tags 32..39
    }
    bm[n_it - 1] = pending;
}

// Exact tail: windows ending at k-mers >= nk were computed from garbage. They can only have set
// bits at positions >= nwin = nk - 7; clear those, then recompute the 7 valid windows that can
// place a minimizer there (windows nwin-7 .. nwin-1, i.e. k-mers g .. g+13 with g = nwin - 7).
// Needs nwin >= 8. n_words = bitmap words the pass wrote (>= nwin / 32 + 1).
K1S_DEV void k1s_fix_tail(const uint32_t *st, uint32_t *bm, int n_words, int nwin, uint32_t topmask)
{
    const int wi = nwin >> 5, r = nwin & 31;
    bm[wi] &= r ? ~(0xffffffffu >> r) : 0u;
    for (int j = wi + 1; j < n_words; ++j) bm[j] = 0u;
    const int g = nwin - 7;
    const int gi = g >> 4;
    const uint32_t sh = 2u * (uint32_t)(g & 15);
    const uint32_t x0 = st[gi], x1 = st[gi + 1], x2 = st[gi + 2];
    const uint32_t h0 = k1s_fsl(x1, x0, sh), h1 = k1s_fsl(x2, x1, sh);
    uint32_t key[14];
#pragma unroll
    for (int j = 0; j < 14; ++j)
        key[j] = ((j == 0 ? h0 : k1s_fsl(h1, h0, 2 * j)) & topmask) | (uint32_t)j;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        uint32_t m = key[i];
#pragma unroll
        for (int j = 1; j < 8; ++j) m = k1s_min(m, key[i + j]);
        const int pos = g + (int)(m & 63u);
        bm[pos >> 5] |= 0x80000000u >> (pos & 31);
    }
}

#ifndef K1S_HOST
// shared-memory accesses by 32-bit shared-window address (keeps address arithmetic in 32 bits)
__device__ __forceinline__ uint32_t k1s_lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t k1s_lds16(uint32_t a)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return (uint32_t)v;
}
__device__ __forceinline__ unsigned long long k1s_lds64(uint32_t a)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void k1s_sts64(uint32_t a, unsigned long long v)
{
    asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void k1s_stg64(unsigned long long p, uint32_t x, uint32_t y)
{
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void k1s_sts32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void k1s_sts16(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory");
}

// compression table, filled once per device by the host (ngsid_ctx_create); every block copies it
__device__ uint32_t k1s_lut_dev[1024];

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K1S_THREADS)
k1_stream_kernel(const uint32_t *__restrict__ packed, const int64_t *__restrict__ woff,
                 const int64_t *__restrict__ off, const int64_t *__restrict__ moff,
                 Minimizer *__restrict__ mins, uint32_t *__restrict__ nmin,
                 uint32_t *__restrict__ lenc, int64_t n_reads, int k, int w, int sw, int bw, int rs, int scap,
                 int32_t *__restrict__ slow_list, int32_t *__restrict__ slow_n)
{
    extern __shared__ __align__(16) uint32_t k1s_smem[];
    uint32_t *lut = k1s_smem;
    uint32_t *regions = k1s_smem + 1024;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    // per-read metadata and the first two 16-byte chunks are requested before the table copy and
    // its barrier, so that their latency overlaps
    const int64_t r = (int64_t)blockIdx.x * K1S_THREADS + threadIdx.x;
    const bool have = r < n_reads;
    const int L = have ? (int)(off[r + 1] - off[r]) : 0;
    const uint4 *pk = reinterpret_cast<const uint4 *>(packed + (have ? woff[r] : 0));
    const int64_t my_moff = have ? moff[r] : 0;
    const int nw = (L + 15) >> 4;                 // raw words of this read
    const int nq = (nw + 3) >> 2;                 // 16-byte chunks
    uint4 cur = (nq > 0) ? __ldg(pk) : make_uint4(0, 0, 0, 0);
    uint4 nxt = (nq > 1) ? __ldg(pk + 1) : make_uint4(0, 0, 0, 0);
    for (int idx = threadIdx.x; idx < 1024; idx += K1S_THREADS) lut[idx] = k1s_lut_dev[idx];
    __syncthreads();
    uint32_t *st = regions + (size_t)threadIdx.x * rs;
    uint32_t *bm = st + sw;
    const uint32_t topmask = ~((1u << (32 - 2 * k)) - 1u);

    // ---- phase A: compression. 64 bases (one 16-byte load) per iteration, next load in flight.
    const int nq_max = __reduce_max_sync(NGSID_FULL_MASK, nq);
    K1SCompress C;
    k1s_compress_init(C, cur.x);
    for (int q = 0; q < nq_max; ++q) {
        uint4 nx2 = make_uint4(0, 0, 0, 0);                  // two loads in flight
        if (q + 2 < nq) nx2 = __ldg(pk + q + 2);
        const int wb = 4 * q;
        k1s_compress_word(C, lut, cur.x, wb + 0 < nw, st, sw);
        k1s_compress_word(C, lut, cur.y, wb + 1 < nw, st, sw);
        k1s_compress_word(C, lut, cur.z, wb + 2 < nw, st, sw);
        k1s_compress_word(C, lut, cur.w, wb + 3 < nw, st, sw);
        cur = nxt;
        nxt = nx2;
    }
    const int Lc = k1s_compress_finish(C, st, sw);        // -1: does not fit the region -> generic kernel

    // ---- phase B: window minima -> bitmap
    const int nk = Lc - k + 1, nwin = nk - 7;
    const bool ok = have && nwin >= 8 && ((nk + 31) >> 5) <= bw - 2;
    const int my_it = ok ? ((nk + 31) >> 5) : 1;
    const int n_it = __reduce_max_sync(NGSID_FULL_MASK, my_it);
    k1s_window_pass(st, bm, n_it, topmask);
    if (ok) {
        k1s_fix_tail(st, bm, n_it, nwin, topmask);
    } else {
        for (int j = 0; j < n_it; ++j) bm[j] = 0u;
    }
    __syncwarp();

    // ---- phase C: bitmaps -> records, K1S_GROUP reads at a time.
    // One warp instruction here serves 32 bitmap words of a few reads (phases A and B serve 32 reads), so this
    // part is written against the instruction count. A lane takes K1S_IPL consecutive bitmap words, a warp scan
    // of their popcounts gives every word the index of its first record, and the lane writes its records
    // (k-mer code re-read from the compressed stream, position) straight into the group's staging area in
    // shared memory, read after read (each read starts on a 16-byte boundary). The staged records of a read
    // then leave as ONE bulk copy shared -> global issued by the read's own lane (cp.async.bulk, the TMA
    // engine does the coalescing); the wait for the engine's reads of the staging area sits behind the next
    // group's popcount scan.
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(k1s_smem);
    const uint32_t rsb = (uint32_t)rs * 4u;
    const uint32_t s_wreg = s_base + 4096u + (uint32_t)(wid * 32) * rsb;          // this warp's 32 regions
    const uint32_t s_stage = s_base + 4096u + (((uint32_t)K1S_THREADS * rsb + 15u) & ~15u) + (uint32_t)wid * (8u * (uint32_t)scap + K1S_WARP_EXTRA);
    const uint32_t s_rstart = s_stage + 8u * (uint32_t)scap;
    const uint32_t bm_off = (uint32_t)sw * 4u;
    const int kshift = 32 - 2 * k;
    const int total_items = K1S_GROUP * n_it;
    const int q_l = (K1S_IPL * lane) / n_it, c_l = (K1S_IPL * lane) % n_it;        // first item of this lane in a pass
    const int dq = (32 * K1S_IPL) / n_it, dc = (32 * K1S_IPL) % n_it;              // advance per pass
    uint32_t my_n = 0;
    bool my_slow = have && !ok;
    bool pending = false;                                                          // bulk copies of the previous group in flight
    for (int g0 = 0; g0 < 32; g0 += K1S_GROUP) {
        if (!__any_sync(NGSID_FULL_MASK, have && lane >= g0)) break;
        // ---- pass 1: popcounts -> index of every word's first record, read starts
        uint32_t mw[3][K1S_IPL], ow[3][K1S_IPL];                                   // up to 3 passes of 96 words (n_it <= 72)
        int qw[3][K1S_IPL], cw[3][K1S_IPL];
        uint32_t run = 0;
        int qa = q_l, ca = c_l;
        int np_ = 0;
        for (int f0 = 0; f0 < total_items && np_ < 3; f0 += 32 * K1S_IPL, ++np_) {
            int q = qa, c = ca;
            uint32_t n = 0;
#pragma unroll
            for (int t = 0; t < K1S_IPL; ++t) {
                const bool v = f0 + K1S_IPL * lane + t < total_items;
                mw[np_][t] = v ? k1s_lds32(s_wreg + (uint32_t)(g0 + q) * rsb + bm_off + 4u * (uint32_t)c) : 0u;
                qw[np_][t] = q; cw[np_][t] = v ? c : -1;
                n += (uint32_t)__popc(mw[np_][t]);
                if (++c >= n_it) { c = 0; ++q; }
            }
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(NGSID_FULL_MASK, incl, d);
                if (lane >= d) incl += v;
            }
            uint32_t o = run + incl - n;
#pragma unroll
            for (int t = 0; t < K1S_IPL; ++t) {
                ow[np_][t] = o;
                if (cw[np_][t] == 0) k1s_sts32(s_rstart + 4u * (uint32_t)qw[np_][t], o);
                o += (uint32_t)__popc(mw[np_][t]);
            }
            run += __shfl_sync(NGSID_FULL_MASK, incl, 31);
            ca += dc; qa += dq;
            if (ca >= n_it) { ca -= n_it; ++qa; }
        }
        if (lane == 0) k1s_sts32(s_rstart + 4u * K1S_GROUP, run);
        __syncwarp();
        // per read of the group: first record (flat) and padding so that every read starts on 16 bytes
        uint32_t rs_[K1S_GROUP + 1], pad_[K1S_GROUP];
#pragma unroll
        for (int q = 0; q <= K1S_GROUP; ++q) rs_[q] = k1s_lds32(s_rstart + 4u * (uint32_t)q);
        uint32_t padsum = 0;
#pragma unroll
        for (int q = 0; q < K1S_GROUP; ++q) { pad_[q] = padsum; padsum += (rs_[q + 1] - rs_[q]) & 1u; }
        const bool overflow = run + padsum > (uint32_t)scap || total_items > 3 * 32 * K1S_IPL;
        if (lane >= g0 && lane < g0 + K1S_GROUP) {
            const int q = lane - g0;
            uint32_t nq = 0;
#pragma unroll
            for (int x = 0; x < K1S_GROUP; ++x) if (x == q) nq = rs_[x + 1] - rs_[x];
            my_n = nq;
            my_slow = my_slow || (have && overflow);
        }
        // the staging area is free once the engine has read the previous group's records
        if (pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pending = false; }
        __syncwarp();
        if (!overflow) {
            // ---- pass 2: records of every word into the staging area; the K1S_IPL words of a lane are
            // independent extraction chains inside one loop (one trip count for the warp: the most bits any
            // word has; a lane whose word has no bit left writes nothing)
            for (int pz = 0; pz < np_; ++pz) {
                uint32_t m[K1S_IPL], w0[K1S_IPL], w1[K1S_IPL], w2[K1S_IPL], dst[K1S_IPL], pbase[K1S_IPL];
                uint32_t most = 0;
#pragma unroll
                for (int t = 0; t < K1S_IPL; ++t) {
                    m[t] = cw[pz][t] < 0 ? 0u : mw[pz][t];
                    const int q = qw[pz][t], c = cw[pz][t] < 0 ? 0 : cw[pz][t];
                    uint32_t padq = 0;
#pragma unroll
                    for (int x = 0; x < K1S_GROUP; ++x) if (x == q) padq = pad_[x];
                    const uint32_t a_st = s_wreg + (uint32_t)(g0 + q) * rsb + 8u * (uint32_t)c;      // stream words 2c, 2c+1, 2c+2
                    w0[t] = k1s_lds32(a_st); w1[t] = k1s_lds32(a_st + 4u); w2[t] = k1s_lds32(a_st + 8u);
                    dst[t] = s_stage + 8u * (ow[pz][t] + padq);
                    pbase[t] = 32u * (uint32_t)c;
                    most = max(most, (uint32_t)__popc(m[t]));
                }
                most = __reduce_max_sync(NGSID_FULL_MASK, most);
                for (uint32_t so = 0; so < 8u * most; so += 8u) {
#pragma unroll
                    for (int t = 0; t < K1S_IPL; ++t) {
                        const uint32_t x = k1s_clz(m[t]);
                        if (m[t] != 0u) {
                            const uint32_t hi = x < 16u ? w0[t] : w1[t], lo = x < 16u ? w1[t] : w2[t];
                            const uint32_t code = k1s_fsl(lo, hi, 2u * x) >> kshift;
                            k1s_sts64(dst[t] + so, ((unsigned long long)(pbase[t] + x) << 32) | code);
                        }
                        m[t] &= k1s_fsr(0x7fffffffu, 0u, x);      // clears bit 31 - x (the bits above it are 0)
                    }
                }
            }
            __syncwarp();
            // ---- one bulk store per read, issued by the read's lane (16-byte multiples: an odd count copies one
            // stale record more into the slack of the read's slots; nmin says how many are valid)
            if (lane >= g0 && lane < g0 + K1S_GROUP && have && ok && my_n > 0) {
                const int q = lane - g0;
                uint32_t first = 0;
#pragma unroll
                for (int x = 0; x < K1S_GROUP; ++x) if (x == q) first = rs_[x] + pad_[x];
                const uint32_t bytes = ((my_n + 1u) & ~1u) * 8u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"((unsigned long long)(mins + my_moff)), "r"(s_stage + 8u * first), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                pending = true;
            }
        }
        __syncwarp();
    }
    if (pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (my_slow) slow_list[atomicAdd(slow_n, 1)] = (int32_t)r;
    else if (have) {
        nmin[r] = my_n;
        lenc[r] = (uint32_t)Lc;
    }
    (void)w;
}
#endif
