// Host build of the per-thread phases of the K1 stream kernel (ngspeciesid_b200/csrc/k1_stream.cuh)
// so that CPU tests can check them against a naive window scan. Test infrastructure only.
#define K1S_HOST
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../ngspeciesid_b200/csrc/k1_stream.cuh"

// seq: ACGT bytes. Runs phase A + B + tail fix-up the way one GPU thread does (n_it_extra extra
// steps emulate a warp whose longest read is longer than this one). Outputs the compressed length,
// and the minimizer (code, position) pairs read back from the bitmap.
extern "C" int k1s_host_minimizers(const char *seq, int L, int k, int n_it_extra, int garbage, int tight,
                                   uint32_t *out_code, uint32_t *out_pos, int cap, int *out_lc)
{
    static uint32_t lut[1024];
    static bool lut_ready = false;
    if (!lut_ready) { for (uint32_t i = 0; i < 1024; ++i) lut[i] = k1s_lut_entry(i); lut_ready = true; }
    // pack like k_pack_kernel: first base most significant, tail of the last word = last base
    const int nw = (L + 15) / 16;
    std::vector<uint32_t> pk(nw + 8, 0);
    for (int wi = 0; wi < nw; ++wi) {
        uint32_t word = 0;
        for (int t = 0; t < 16; ++t) {
            int b = wi * 16 + t;
            if (b >= L) b = L - 1;
            const char ch = seq[b];
            const uint32_t code = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3;
            word |= code << (30 - 2 * t);
        }
        pk[wi] = word;
    }
    K1SGeom g = k1s_geometry(tight ? L : 2 * L + 32 * n_it_extra + 64, k);
    std::vector<uint32_t> region(g.rs + 8, garbage ? 0xdeadbeefu : 0u);
    if (garbage) for (size_t i = 0; i < region.size(); ++i) region[i] = 0x9e3779b9u * (uint32_t)(i + 1 + garbage);
    uint32_t *st = region.data(), *bm = st + g.sw;
    K1SCompress C;
    k1s_compress_init(C, nw ? pk[0] : 0);
    const int nq = (nw + 3) / 4 + (n_it_extra ? 1 : 0);
    for (int q = 0; q < nq; ++q)
        for (int j = 0; j < 4; ++j) {
            const int wi = 4 * q + j;
            k1s_compress_word(C, lut, wi < nw ? pk[wi] : 0x12345678u, wi < nw, st, g.sw);
        }
    const int Lc = k1s_compress_finish(C, st, g.sw);
    *out_lc = Lc;
    if (Lc < 0) return -4;                        // does not fit: slow-list case
    const int nk = Lc - k + 1, nwin = nk - 7;
    if (nwin < 8) return -1;                      // slow-list case
    const uint32_t topmask = ~((1u << (32 - 2 * k)) - 1u);
    const int n_it = (nk + 31) / 32 + n_it_extra;
    if (n_it > g.n_it_max) return -3;
    k1s_window_pass(st, bm, n_it, topmask);
    k1s_fix_tail(st, bm, n_it, nwin, topmask);
    int n = 0;
    for (int wi = 0; wi < n_it; ++wi) {
        uint32_t m = bm[wi];
        while (m) {
            const uint32_t p = k1s_clz(m);
            m &= ~(0x80000000u >> p);
            const uint32_t pos = 32u * wi + p;
            if (n >= cap) return -2;
            const uint32_t x = k1s_fsl(st[(pos >> 4) + 1], st[pos >> 4], 2u * (pos & 15u));
            out_code[n] = x >> (32 - 2 * k);
            out_pos[n] = pos;
            ++n;
        }
    }
    return n;
}
