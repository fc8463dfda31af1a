// K2 + K3: minimizer table (hash + postings), hit aggregation and the mapped-fraction decision.
// Reference semantics: cluster.get_all_hits (modules/cluster.py:43-62), cluster.get_best_cluster
// (modules/cluster.py:67-127), candidate order of get_best_cluster_block_align (:172-205) and
// the new-representative insert (:328-334). One warp per read.
#pragma once
#include "ngsid_internal.cuh"

#define NGSID_EMPTY_KEY 0xffffffffu
#define SLOT_DEAD 0
#define SLOT_VALID 1
#define SLOT_TENTATIVE 2

#define DEC_NEW (-1)
#define DEC_SKIP (-2)
#define DEC_NEED_ALIGN (-3)

#define ACACHE_N 8          // alignment results remembered per read

struct PostingNode { int32_t slot; int32_t next; };

struct MapTable {
    uint32_t *keys;         // open addressing, NGSID_EMPTY_KEY = free
    int32_t *heads;         // head of the posting list of that key
    PostingNode *nodes;
    uint32_t cap_mask;
};

struct DeviceClusterParams {
    int32_t k, min_shared, symmetric, pad;
    double min_fraction, mapped_threshold, aligned_threshold;
    int32_t max_gap[225];
};

__device__ __forceinline__ uint32_t hash_kmer(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ int32_t table_lookup(const MapTable &t, uint32_t key)
{
    uint32_t h = hash_kmer(key) & t.cap_mask;
    while (true) {
        uint32_t kk = t.keys[h];
        if (kk == key) return t.heads[h];
        if (kk == NGSID_EMPTY_KEY) return -1;
        h = (h + 1) & t.cap_mask;
    }
}

// ---- insert the minimizers of new representative slots [slot0, slot0 + n) ----------------------
// One warp per slot. A k-mer occurring at several positions of the representative is inserted
// once (the reference keeps a set of ids per k-mer, cluster.py:330-334).
__global__ void k2_insert_kernel(MapTable t, int32_t *__restrict__ node_cursor, int32_t node_cap,
                                 int32_t *__restrict__ err_flag,
                                 const int32_t *__restrict__ slot_read, int slot0, int n_slots,
                                 const Minimizer *__restrict__ mins,
                                 const int64_t *__restrict__ moff,
                                 const uint32_t *__restrict__ nmin)
{
    int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n_slots) return;
    uint32_t lane = lane_id();
    int slot = slot0 + warp;
    int read = slot_read[slot];
    const Minimizer *m = mins + moff[read];
    int n = (int)nmin[read];
    // The duplicate scan runs to a warp-uniform bound without an early exit: with a per-lane loop the
    // lanes leave it one by one and each of them then pays the latency of its three atomics alone
    // (measured: 0.28 ms per launch, 42 k warp-instructions per representative).
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + (int)lane;
        const bool have = j < n;
        const uint32_t key = have ? m[j].x : 0u;
        bool dup = false;
        const int qend = min(n, j0 + 32) - 1;
        for (int q = 0; q < qend; ++q) {
            const uint32_t kq = m[q].x;
            dup |= (q < j) && (kq == key);
        }
        __syncwarp();
        if (!have || dup) continue;
        uint32_t h = hash_kmer(key) & t.cap_mask;
        while (true) {
            uint32_t old = atomicCAS(&t.keys[h], NGSID_EMPTY_KEY, key);
            if (old == NGSID_EMPTY_KEY || old == key) break;
            h = (h + 1) & t.cap_mask;
        }
        int32_t node = atomicAdd(node_cursor, 1);
        if (node >= node_cap) { *err_flag = 3; continue; }      // host sizes the pool before the launch: never expected
        t.nodes[node].slot = slot;
        t.nodes[node].next = atomicExch(&t.heads[h], node);
    }
}

// ---- grow: re-insert every key of the old table into a larger one -------------------------------
__global__ void k2_rehash_kernel(const uint32_t *__restrict__ old_keys,
                                 const int32_t *__restrict__ old_heads, uint32_t old_cap,
                                 MapTable t)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    uint32_t key = old_keys[i];
    if (key == NGSID_EMPTY_KEY) return;
    uint32_t h = hash_kmer(key) & t.cap_mask;
    while (true) {
        uint32_t old = atomicCAS(&t.keys[h], NGSID_EMPTY_KEY, key);
        if (old == NGSID_EMPTY_KEY) break;
        h = (h + 1) & t.cap_mask;
    }
    t.heads[h] = old_heads[i];
}

// ---- alignment request emitted by the map kernel -------------------------------------------------
struct AlignRequest {
    int32_t pos;        // position in `order` of the read that asked
    int32_t slot;       // candidate representative slot
    int32_t read_a;     // s1 = the read
    int32_t read_b;     // s2 = the representative
    int32_t open;       // gap open penalty (cluster.py:189-196)
    int32_t match_id;   // floor((1 - err_sum) * k)  (cluster.py:198)
};

struct AlignCacheEntry { int32_t slot; int32_t passed; };   // slot < 0 = free
// Results beyond the ACACHE_N inline entries of a read (more than ACACHE_N failed candidates over
// all rounds of a pass: the reference tries every tied candidate, cluster.py:174-205) go to a
// per-read chain of pool nodes.
struct AlignCacheOvf { int32_t slot; int32_t passed; int32_t next; };

struct MapArgs {
    MapTable table;
    const DeviceClusterParams *params;
    const int32_t *list;            // positions (into order) to evaluate
    int n_list;
    const int32_t *order;           // position -> read index
    const int32_t *slot_read;       // slot -> read index
    const int32_t *slot_pos;        // slot -> position in order (-1 for initial representatives)
    const uint8_t *slot_state;
    int n_slots;
    const Minimizer *mins;
    const int64_t *moff;
    const uint32_t *nmin;
    const uint32_t *lenc;
    const uint8_t *bucket;
    const double *erru;
    const uint32_t *acc_rank;
    uint32_t *scratch;              // per warp: cnt[scap] | spos[scap] | touched[scap]
    int scap;
    AlignCacheEntry *acache;        // per position: ACACHE_N entries
    const int32_t *aovf_head;       // per position: first overflow node or -1
    int acache_inline;              // inline entries in use (ACACHE_N; tests lower it to reach the chain)
    const AlignCacheOvf *aovf;
    int32_t *dec;                   // per position: representative read index / DEC_*
    uint8_t *via;                   // per position: 0 new, 1 map, 2 align
    AlignRequest *req;
    int32_t *req_n;
    int32_t *err_flag;
    // Alignment prefetch for the tentative representatives of a tile (see ngsid_cluster): results
    // of pairs (tentative read, slot) aligned ahead of the order-dependent resolution.
    // spec_mat[row * spec_cols + slot] = -1 unknown / 0 failed / 1 passed, row = spec_u[pos - spec_lo].
    const int32_t *spec_u;          // nullptr: no prefetch table
    int8_t *spec_mat;
    int spec_lo, spec_hi, spec_cols;
    int spec_mode;                  // 1: this launch only emits the prefetch requests
};

// cached alignment result of (position, slot): -1 unknown, 0 failed, 1 passed
__device__ __forceinline__ int acache_lookup(const MapArgs &A, int pos, int slot)
{
    const AlignCacheEntry *ac = A.acache + (size_t)pos * ACACHE_N;
    int cached = -1;
#pragma unroll
    for (int e = 0; e < ACACHE_N; ++e)
        if (ac[e].slot == slot) cached = ac[e].passed;
    if (cached < 0 && ac[A.acache_inline - 1].slot >= 0)
        for (int o = A.aovf_head[pos]; o >= 0; o = A.aovf[o].next)
            if (A.aovf[o].slot == slot) { cached = A.aovf[o].passed; break; }
    return cached;
}

// gap open penalty and match_id of a (read, representative) pair (cluster.py:185-198)
__device__ __forceinline__ AlignRequest make_align_request(const MapArgs &A, int pos, int slot, int read, int k)
{
    const int rep_read = A.slot_read[slot];
    double es = __dadd_rn(A.erru[read], A.erru[rep_read]);
    int go = (es <= 0.01) ? 5 : (es <= 0.04) ? 4 : (es <= 0.1) ? 3 : 2;
    int mid = (int)floor(__dmul_rn(__dadd_rn(1.0, -es), (double)k));
    AlignRequest rq = {pos, slot, read, rep_read, go, mid};
    return rq;
}

// rank key of cluster.py:79 / :174, descending: (n_hits, sum of positions, accession string)
__device__ __forceinline__ bool key_greater(uint32_t n1, uint32_t s1, uint32_t a1,
                                            uint32_t n2, uint32_t s2, uint32_t a2)
{
    if (n1 != n2) return n1 > n2;
    if (s1 != s2) return s1 > s2;
    return a1 > a2;
}

#define VISITED_BIT 0x80000000u
#define K2_SMEM_SLOTS 192      // representatives up to which the map kernel counts hits in shared memory

__global__ void __launch_bounds__(256) k2_map_kernel(MapArgs A)
{
    const DeviceClusterParams &P = *A.params;
    const uint32_t lane = lane_id();
    const int warps_per_block = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    // per-warp hit counters: in shared memory while the table holds few representatives (amplicon
    // data: tens), so that the counting atomics of pass 1 do not pay an L2 round trip per posting
    // node; in the global scratch otherwise
    __shared__ uint32_t s_scr[8][3 * K2_SMEM_SLOTS];
    const bool scr_shared = A.n_slots <= K2_SMEM_SLOTS;
    const int scr_stride = scr_shared ? K2_SMEM_SLOTS : A.scap;
    uint32_t *cnt = scr_shared ? s_scr[threadIdx.x >> 5] : A.scratch + (size_t)gwarp * 3 * A.scap;
    uint32_t *spos = cnt + scr_stride;
    uint32_t *touched = spos + scr_stride;
    if (scr_shared) {
        for (int x = (int)lane; x < 3 * K2_SMEM_SLOTS; x += 32) cnt[x] = 0u;
        __syncwarp();
    }
    __shared__ uint32_t s_ntouched[8];
    uint32_t *ntouched = &s_ntouched[threadIdx.x >> 5];

    for (int li = gwarp; li < A.n_list; li += nwarps) {
        const int pos = A.list[li];
        const int read = A.order[pos];
        const int nm = (int)A.nmin[read];
        const int len_c = (int)A.lenc[read];
        if (len_c < P.k) {                      // cluster.py:266-268
            if (lane == 0) { A.dec[pos] = DEC_SKIP; A.via[pos] = 0; }
            continue;
        }
        const Minimizer *m = A.mins + A.moff[read];
        if (lane == 0) *ntouched = 0;
        __syncwarp();

        // ---- pass 1: hits per candidate slot (cluster.py:43-62)
        for (int j = lane; j < nm; j += 32) {
            Minimizer mj = m[j];
            int32_t node = table_lookup(A.table, mj.x);
            while (node >= 0) {
                PostingNode pn = A.table.nodes[node];
                int s = pn.slot;
                const uint8_t stt = A.slot_state[s];
                if ((stt == SLOT_VALID || (A.spec_mode && stt == SLOT_TENTATIVE)) && A.slot_pos[s] < pos) {
                    uint32_t old = atomicAdd(&cnt[s], 1u);
                    atomicAdd(&spos[s], mj.y);
                    if (old == 0) touched[atomicAdd(ntouched, 1u)] = (uint32_t)s;
                }
                node = pn.next;
            }
        }
        __threadfence_block();
        __syncwarp();
        const int nt = (int)*ntouched;

        if (A.spec_mode) {
            // Prefetch pass: whatever subset of the tentative slots survives, an alignment of this
            // read is only ever asked for a slot whose hit count equals the top count of that
            // world, which is >= min_shared and >= the best count among the certain slots. Ask for
            // every (read, slot) pair above that bound that is not known yet.
            uint32_t topc = 0;
            for (int t = lane; t < nt; t += 32) {
                const int s = (int)touched[t];
                if (A.slot_state[s] == SLOT_VALID) topc = max(topc, cnt[s]);
            }
            for (int d = 16; d > 0; d >>= 1) topc = max(topc, __shfl_xor_sync(NGSID_FULL_MASK, topc, d));
            const uint32_t thr = max(topc, (uint32_t)max(P.min_shared, 1));
            const int row = A.spec_u[pos - A.spec_lo];
            for (int t = lane; t < nt; t += 32) {
                const int s = (int)touched[t];
                bool ask = row >= 0 && s < A.spec_cols && cnt[s] >= thr;
                if (ask && acache_lookup(A, pos, s) >= 0) ask = false;
                if (ask && A.spec_mat[(size_t)row * A.spec_cols + s] < 0)
                    A.req[atomicAdd(A.req_n, 1)] = make_align_request(A, pos, s, read, P.k);
            }
            for (int t = lane; t < nt; t += 32) {
                const int s = (int)touched[t];
                cnt[s] = 0; spos[s] = 0;
            }
            __syncwarp();
            continue;
        }

        int decision = DEC_NEW;
        int via = 0;
        uint32_t top = 0;
        for (int t = lane; t < nt; t += 32) top = max(top, cnt[touched[t]]);
        for (int d = 16; d > 0; d >>= 1) top = max(top, __shfl_xor_sync(NGSID_FULL_MASK, top, d));

        if (nt > 0 && (int)top >= P.min_shared) {
            const double cut = __dmul_rn(P.min_fraction, (double)top);
            const int b_read = A.bucket[read];
            // ---- mapping test in rank order (cluster.py:84-125)
            while (true) {
                // best unvisited candidate that still qualifies
                uint32_t bn = 0, bs = 0, ba = 0; int bslot = -1;
                for (int t = lane; t < nt; t += 32) {
                    int s = (int)touched[t];
                    uint32_t c = cnt[s];
                    if (c & VISITED_BIT) continue;
                    if ((double)c < cut || (int)c < P.min_shared) continue;
                    uint32_t sp = spos[s], ar = A.acc_rank[A.slot_read[s]];
                    if (bslot < 0 || key_greater(c, sp, ar, bn, bs, ba)) { bn = c; bs = sp; ba = ar; bslot = s; }
                }
                for (int d = 16; d > 0; d >>= 1) {
                    uint32_t on = __shfl_xor_sync(NGSID_FULL_MASK, bn, d);
                    uint32_t os = __shfl_xor_sync(NGSID_FULL_MASK, bs, d);
                    uint32_t oa = __shfl_xor_sync(NGSID_FULL_MASK, ba, d);
                    int oslot = __shfl_xor_sync(NGSID_FULL_MASK, bslot, d);
                    if (oslot >= 0 && (bslot < 0 || key_greater(on, os, oa, bn, bs, ba))) {
                        bn = on; bs = os; ba = oa; bslot = oslot;
                    }
                }
                if (bslot < 0) break;
                if (lane == 0) cnt[bslot] |= VISITED_BIT;
                __syncwarp();
                // hit list of (read, bslot) in minimizer order; gaps bridged when <= max_gap
                const int rep_read = A.slot_read[bslot];
                const int gmax = P.max_gap[b_read * 15 + A.bucket[rep_read]];
                int carry_idx = -1; uint32_t carry_pos = 0; int total = 0;
                for (int j0 = 0; j0 < nm; j0 += 32) {
                    int j = j0 + (int)lane;
                    bool hit = false; uint32_t pj = 0;
                    if (j < nm) {
                        Minimizer mj = m[j];
                        pj = mj.y;
                        int32_t node = table_lookup(A.table, mj.x);
                        while (node >= 0) {
                            PostingNode pn = A.table.nodes[node];
                            if (pn.slot == bslot) { hit = true; break; }
                            node = pn.next;
                        }
                    }
                    uint32_t hm = __ballot_sync(NGSID_FULL_MASK, hit);
                    uint32_t below = hm & ((1u << lane) - 1u);
                    int src = below ? (31 - __clz(below)) : 0;
                    uint32_t ppos = __shfl_sync(NGSID_FULL_MASK, pj, src);
                    int pidx = j0 + src;
                    if (!below) { ppos = carry_pos; pidx = carry_idx; }
                    int contrib = 0;
                    if (hit && (j - pidx - 1) <= gmax) contrib = (int)pj - (int)ppos;
                    for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(NGSID_FULL_MASK, contrib, d);
                    total += contrib;
                    if (hm) {
                        int last = 31 - __clz(hm);
                        carry_idx = j0 + last;
                        carry_pos = __shfl_sync(NGSID_FULL_MASK, pj, last);
                    }
                }
                if ((nm - carry_idx - 1) <= gmax) total += len_c - (int)carry_pos;
                double ratio = __ddiv_rn((double)total, (double)len_c);
                if (P.symmetric) {
                    double r2 = __ddiv_rn((double)total, (double)A.lenc[rep_read]);
                    ratio = fmin(ratio, r2);
                }
                if (ratio > P.mapped_threshold) { decision = rep_read; via = 1; break; }
            }

            // ---- alignment stage: candidates tied for the top hit count, rank order
            // (cluster.py:174-182); results of earlier K4 rounds come from the per-read cache.
            if (decision == DEC_NEW) {
                // clear the visited marks of the tied candidates, then walk them in order
                for (int t = lane; t < nt; t += 32) {
                    int s = (int)touched[t];
                    cnt[s] &= ~VISITED_BIT;
                }
                __syncwarp();
                while (true) {
                    uint32_t bs = 0, ba = 0; int bslot = -1;
                    for (int t = lane; t < nt; t += 32) {
                        int s = (int)touched[t];
                        uint32_t c = cnt[s];
                        if (c != top) continue;          // visited ones carry the mark bit
                        uint32_t sp = spos[s], ar = A.acc_rank[A.slot_read[s]];
                        if (bslot < 0 || key_greater(top, sp, ar, top, bs, ba)) { bs = sp; ba = ar; bslot = s; }
                    }
                    for (int d = 16; d > 0; d >>= 1) {
                        uint32_t os = __shfl_xor_sync(NGSID_FULL_MASK, bs, d);
                        uint32_t oa = __shfl_xor_sync(NGSID_FULL_MASK, ba, d);
                        int oslot = __shfl_xor_sync(NGSID_FULL_MASK, bslot, d);
                        if (oslot >= 0 && (bslot < 0 || key_greater(top, os, oa, top, bs, ba))) {
                            bs = os; ba = oa; bslot = oslot;
                        }
                    }
                    if (bslot < 0) break;                 // every tied candidate failed -> new
                    if (lane == 0) cnt[bslot] |= VISITED_BIT;
                    __syncwarp();
                    int cached = acache_lookup(A, pos, bslot);   // -1 unknown, 0 failed, 1 passed
                    if (cached < 0 && A.spec_u && pos >= A.spec_lo && pos < A.spec_hi && bslot < A.spec_cols) {
                        const int row = A.spec_u[pos - A.spec_lo];
                        if (row >= 0) cached = (int)A.spec_mat[(size_t)row * A.spec_cols + bslot];
                    }
                    if (cached == 1) { decision = A.slot_read[bslot]; via = 2; break; }
                    if (cached == 0) continue;
                    // not aligned yet: ask for it (one candidate per round, like the reference)
                    if (lane == 0) A.req[atomicAdd(A.req_n, 1)] = make_align_request(A, pos, bslot, read, P.k);
                    decision = DEC_NEED_ALIGN;
                    break;
                }
            }
        }
        // ---- clean the scratch for the next read of this warp
        for (int t = lane; t < nt; t += 32) {
            int s = (int)touched[t];
            cnt[s] = 0; spos[s] = 0;
        }
        if (lane == 0) { A.dec[pos] = decision; A.via[pos] = (uint8_t)via; }
        __syncwarp();
    }
}

// ---- apply K4 results: turn window counts into pass/fail cache entries ---------------------------
// alignment_ratio = count / len(s1) >= aligned_threshold (cluster.py:167-168, 200-203)
__global__ void k2_apply_align_kernel(const AlignRequest *__restrict__ req, int n_req,
                                      const int32_t *__restrict__ k4cnt,
                                      const int64_t *__restrict__ off,
                                      const DeviceClusterParams *__restrict__ params,
                                      AlignCacheEntry *__restrict__ acache,
                                      int32_t *__restrict__ aovf_head, AlignCacheOvf *__restrict__ aovf,
                                      int32_t *__restrict__ aovf_cursor, int32_t aovf_cap, int inline_n,
                                      int32_t *__restrict__ list_out, int32_t *err_flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_req) return;
    AlignRequest rq = req[i];
    double n1 = (double)(off[rq.read_a + 1] - off[rq.read_a]);
    double n2 = (double)(off[rq.read_b + 1] - off[rq.read_b]);
    atomicAdd(reinterpret_cast<unsigned long long *>(err_flag + 2), (unsigned long long)(n1 * n2));
    double ratio = __ddiv_rn((double)k4cnt[i], n1);
    if (params->symmetric) ratio = fmin(ratio, __ddiv_rn((double)k4cnt[i], n2));
    int passed = ratio >= params->aligned_threshold ? 1 : 0;
    AlignCacheEntry *ac = acache + (size_t)rq.pos * ACACHE_N;
    int e = 0;
    while (e < inline_n && ac[e].slot >= 0) ++e;
    if (e == inline_n) {
        // one request per read and round, so the chain of a position has a single writer
        const int32_t o = atomicAdd(aovf_cursor, 1);
        if (o >= aovf_cap) { *err_flag = 2; return; }
        aovf[o].slot = rq.slot; aovf[o].passed = passed; aovf[o].next = aovf_head[rq.pos];
        aovf_head[rq.pos] = o;
    } else {
        ac[e].slot = rq.slot;
        ac[e].passed = passed;
    }
    list_out[i] = rq.pos;
}

// the same for the prefetched pairs: one cell of the prefetch table per request
__global__ void k2_apply_spec_kernel(const AlignRequest *__restrict__ req, int n_req,
                                     const int32_t *__restrict__ k4cnt, const int64_t *__restrict__ off,
                                     const DeviceClusterParams *__restrict__ params,
                                     const int32_t *__restrict__ spec_u, int spec_lo, int spec_cols,
                                     int8_t *__restrict__ spec_mat, int32_t *err_flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_req) return;
    AlignRequest rq = req[i];
    double n1 = (double)(off[rq.read_a + 1] - off[rq.read_a]);
    double n2 = (double)(off[rq.read_b + 1] - off[rq.read_b]);
    atomicAdd(reinterpret_cast<unsigned long long *>(err_flag + 2), (unsigned long long)(n1 * n2));
    double ratio = __ddiv_rn((double)k4cnt[i], n1);
    if (params->symmetric) ratio = fmin(ratio, __ddiv_rn((double)k4cnt[i], n2));
    spec_mat[(size_t)spec_u[rq.pos - spec_lo] * spec_cols + rq.slot] = ratio >= params->aligned_threshold ? 1 : 0;
}

__global__ void k2_fill_acache_kernel(AlignCacheEntry *ac, int32_t *aovf_head, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ac[i].slot = -1; ac[i].passed = 0; }
    if (i < n / ACACHE_N) aovf_head[i] = -1;
}

__global__ void k2_iota_kernel(int32_t *list, int first, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = first + i;
}

// ---- stand-alone hit table: cluster.get_all_hits (modules/cluster.py:43-62) for (read, representative slot)
// pairs: hit count and sum of hit positions, what the ranking of cluster.py:79 uses. One warp per read.
__global__ void k2_hits_kernel(MapTable t, const int32_t *__restrict__ reads, int n_reads, int n_slots,
                               const Minimizer *__restrict__ mins, const int64_t *__restrict__ moff,
                               const uint32_t *__restrict__ nmin, const int32_t *__restrict__ slot_read,
                               uint32_t *__restrict__ out_cnt, uint32_t *__restrict__ out_possum)
{
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n_reads) return;
    const uint32_t lane = lane_id();
    const int read = reads[warp];
    const Minimizer *m = mins + moff[read];
    const int nm = (int)nmin[read];
    for (int j = lane; j < nm; j += 32) {
        const Minimizer mj = m[j];
        int32_t node = table_lookup(t, mj.x);
        while (node >= 0) {
            const PostingNode pn = t.nodes[node];
            if (slot_read[pn.slot] != read) {                    // a read does not hit itself (cluster.py:52)
                atomicAdd(&out_cnt[(size_t)warp * n_slots + pn.slot], 1u);
                atomicAdd(&out_possum[(size_t)warp * n_slots + pn.slot], mj.y);
            }
            node = pn.next;
        }
    }
}
