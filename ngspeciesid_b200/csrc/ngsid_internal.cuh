// Internal definitions shared by the kernels and the C ABI (single translation unit: ngsid_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/ngsid.h"

#define NGSID_NEG_INF (-(1 << 29))
#define NGSID_FULL_MASK 0xffffffffu

// ---- device buffer that only ever grows -------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grows to at least `bytes` and keeps the first `keep` bytes (copied on `stream`; the old block is
    // freed after the copy has finished)
    cudaError_t grow_keep(size_t bytes, size_t keep, cudaStream_t stream) {
        if (bytes <= cap) return cudaSuccess;
        void *np = nullptr;
        size_t want = bytes + bytes / 2 + 256;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) { e = cudaMalloc(&np, bytes); want = bytes; }
        if (e != cudaSuccess) return e;
        if (p && keep) {
            e = cudaMemcpyAsync(np, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (p) cudaFree(p);
        p = np; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Packed minimizer record: x = 2k-bit k-mer code (first base most significant), y = position in
// homopolymer-compressed coordinates.
typedef uint2 Minimizer;

struct ngsid_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t up_ev[2] = {nullptr, nullptr};      // pacing of the upload (two pieces in flight)
    cudaEvent_t pev[6][2] = {};
    bool pev_valid[6] = {false, false, false, false, false, false};
    float phase_acc[6] = {0, 0, 0, 0, 0, 0};
    std::vector<cudaEvent_t> ev_pool;     // pairs of events around launches inside a clustering pass
    std::vector<int> ev_kind;
    size_t ev_used = 0;
    std::string err;
    int64_t launches = 0;
    int64_t poa_cells = 0;            // DP cells of the last ngsid_poa_consensus call
    float poa_ms[3] = {-1.f, -1.f, -1.f};   // last K5 call: kernels + copies, host graph work, whole call

    // ---- uploaded reads
    int64_t n_reads = 0, total_bases = 0, total_words = 0;
    int max_len = 0;
    std::vector<int64_t> h_off;       // n+1 byte offsets
    std::vector<int64_t> h_woff;      // n+1 word offsets of the packed reads
    DevBuf d_seq, d_qual, d_off, d_packed, d_woff, d_flag, d_rflag;
    // reads with a base outside ACGT (exception path of K1, csrc/k1_exceptions.cuh) and the k-mers with such
    // a base that became minimizers: code = 1 << 30 | index into xkmers
    std::vector<int32_t> x_reads;
    std::vector<std::string> xkmers;
    std::unordered_map<std::string, uint32_t> xkmer_id;

    // ---- K1 results
    int k = 0, w = 0;
    bool have_min = false;
    int k4_shape = 0;                 // 0: choose per launch, 1: always one warp per pair, 2: always one block per pair
    int k4_tb = 0;                    // option 4: traceback 0 per launch by the number of pairs, 1 warp per pair, 2 thread per pair
    bool use_payload_k4 = false;      // option 2: the trace-free payload kernel instead of DP + traceback
    int k1_variant = 2;               // 2: stream kernel where it applies (default), 0: generic warp-per-read kernel for every (k,w)
    std::vector<int64_t> h_moff;      // n+1 offsets into d_mins (slack CSR: capacity per read)
    std::vector<uint32_t> h_nmin;     // mirror of d_nmin (filled lazily)
    bool h_nmin_valid = false;
    DevBuf d_moff, d_mins, d_nmin, d_lenc;

    // ---- K0 results
    bool have_q = false;
    DevBuf d_errc, d_erru, d_bucket, d_phred, d_thr;
    DevBuf d_ss_tab, d_ss_score, d_ss_err;       // sort-stage scores (row f.1)

    // ---- clustering scratch (see cluster_driver.cuh)
    DevBuf d_keys, d_heads, d_nodes, d_cursor, d_slot_read, d_slot_pos, d_slot_state;
    DevBuf d_order, d_accrank, d_dec, d_aux, d_via, d_list, d_scratch, d_params;
    DevBuf d_poa_dir, d_poa_arena, d_poa_meta, d_poa_h, d_poa_out, d_poa_len, d_poa_nodes, d_poa_err, d_job_off, d_lsrc, d_lbeg, d_llen;
    DevBuf d_trace, d_ends, d_auxseq, d_aoff, d_win, d_match, d_cols;
    DevBuf d_aovf, d_aovf_head;

    // ---- multi-GPU data plane (nccl_plane.cuh)
    void *nccl_comm = nullptr;
    bool nccl_owned = false;
    int nccl_rank = 0, nccl_nranks = 1;
    DevBuf d_cc_a, d_cc_b, d_cc_c;
    DevBuf d_req, d_reqn, d_acache, d_k4cnt, d_k4score, d_newslots, d_pa, d_pb, d_po, d_pm;
    DevBuf d_cl[6];                    // buffers a clustering pass borrows (spare table pair, second list, error words, prefetch tables): kept
                                       // between passes -- cudaMalloc / cudaFree synchronise the whole device, other contexts' streams included
};

#define CUDA_TRY(ctx, call)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                   \
            return (_e == cudaErrorMemoryAllocation) ? NGSID_ENOMEM : NGSID_ECUDA;             \
        }                                                                                      \
    } while (0)

#define KERNEL_CHECK(ctx)                                                                      \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(_e);              \
            return NGSID_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

static inline int fail(ngsid_ctx *ctx, int code, const std::string &msg) {
    ctx->err = msg;
    return code;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// Alignment scoring of bases outside ACGT: the reference scores with parasail.matrix_create("ACGT", 2, -2)
// (modules/cluster.py:131), where such a base is a mismatch against everything, itself included. Rows and
// columns map them to two different sentinels, so the DP's equality test never fires for them. (The block
// statistic compares the raw characters, modules/cluster.py:147, so the traceback reads the raw bases.)
__device__ __forceinline__ bool k4_is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
__device__ __forceinline__ uint32_t k4_row_base(uint32_t c) { return k4_is_acgt(c) ? c : 0xfeu; }
__device__ __forceinline__ uint8_t k4_col_base(uint32_t c) { return (uint8_t)(k4_is_acgt(c) ? c : 0xfdu); }
