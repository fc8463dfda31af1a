/*
 * ngsid.h -- C ABI of libngsid.so: the B200 (sm_100a) implementation of NGSpeciesID's
 * data-parallel hot paths. Plain pointers and sizes only; every buffer is caller-owned
 * (numpy / ctypes on the Python side); every entry returns 0 on success and a negative
 * NGSID_E* code on failure, with a message available from ngsid_last_error().
 *
 * One ngsid_ctx per GPU. A context is not thread-safe; distinct contexts are independent.
 * There is no CPU fallback behind any entry: without a CUDA device ngsid_ctx_create fails.
 *
 * The reference (ksahlin/NGSpeciesID v0.3.1) has no FFI: its hot path is plain Python calling
 * parasail / spoa / racon. Each entry below names the reference function(s) whose arithmetic it
 * replaces (file:line relative to the reference root); INTEGRATION.md shows the ctypes stub a
 * maintainer would add to modules/cluster.py and modules/consensus.py to bind them.
 */
#ifndef NGSID_H
#define NGSID_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngsid_ctx ngsid_ctx;

enum {
    NGSID_OK = 0,
    NGSID_EINVAL = -1,      /* bad argument (k/w out of range, null pointer, ...)            */
    NGSID_ECUDA = -2,       /* CUDA runtime error (message has the cudaError string)         */
    NGSID_ENOMEM = -3,      /* device or host allocation failed                              */
    NGSID_EUNSUPPORTED = -4,/* input outside what this build handles (non-ACGT base, k > 15) */
    NGSID_ESTATE = -5       /* call order violated (e.g. cluster before minimizers)          */
};

/* ---- context ------------------------------------------------------------------------------ */
int ngsid_ctx_create(int device_id, ngsid_ctx **out);
void ngsid_ctx_destroy(ngsid_ctx *ctx);
const char *ngsid_last_error(const ngsid_ctx *ctx);   /* valid until the next call on ctx */
int ngsid_version(void);                               /* ABI version, currently 1 */
/* Number of kernels this context has launched since creation / since the last reset. */
int64_t ngsid_launch_count(const ngsid_ctx *ctx);
/* DP cells (graph rows x layer bases, summed over layers and jobs) of the last ngsid_poa_consensus. */
int64_t ngsid_poa_cells(const ngsid_ctx *ctx);
void ngsid_reset_launch_count(ngsid_ctx *ctx);
int ngsid_sync(ngsid_ctx *ctx);
/* Device time (CUDA events on the context's stream) of the most recent run of a phase, in ms:
 * which = 0 pack (inside upload), 1 K1 minimizers, 2 K0 quality stats, 3 whole clustering pass,
 * 4 sum of K4 launches inside the last clustering pass, 5 sum of map launches inside it;
 * last ngsid_poa_consensus call (host clock): 6 kernels + copies, 7 host graph work, 8 whole call.
 * Returns a negative value when that phase has not run.                                        */
float ngsid_phase_ms(ngsid_ctx *ctx, int which);
/* Tuning / test switches. option 1 selects the K1 kernel: 0 (default) the stream kernel for
 * w-k+1 == 8, k <= 13 and the generic warp-per-read kernel otherwise; 1 the generic kernel for
 * every (k, w).
 * option 2: value != 0 selects the trace-free payload variant of K4.
 * option 3 selects the shape of the K4 DP kernel: 0 (default) per launch by the number of pairs,
 * 1 always one warp per pair, 2 always one thread block per pair (pipelined strips).
 * option 4 selects the K4 traceback kernel: 0 (default) per launch by the number of pairs,
 * 1 always one warp per pair, 2 always one thread per pair (no window breaking points).         */
int ngsid_set_option(ngsid_ctx *ctx, int option, int value);

/* Page-locked host memory for the arrays ngsid_upload_reads takes (no context needed): an upload from it runs at
 * the full PCIe rate and the host layer can fill it in place. The reference has no counterpart: it hands reads over
 * as Python str (modules/cluster.py:207). */
int ngsid_pinned_alloc(void **out, int64_t bytes);
int ngsid_pinned_free(void *p);

/* ---- read upload ---------------------------------------------------------------------------
 * seq / qual: ASCII bases and PHRED+33 qualities of n_reads reads, concatenated; offsets has
 * n_reads+1 entries (offsets[0] == 0). Copies host->device and packs bases 2 bit/base on the device
 * (A,C,G,T = 0..3; first base in the most significant bits of each 32-bit word; every read starts
 * on a word boundary). Reads with any other character (N, IUPAC codes, lower case) are legal, as in
 * the reference, which compares raw characters: their minimizers come from an exact exception path
 * (see ngsid_kmer_string), alignments score such a base as a mismatch and the block statistic
 * compares the raw characters.
 * Replaces nothing in the reference (its reads are Python str); it is the layout the kernels read.
 * A new upload discards all per-read results of the previous one.                              */
int ngsid_upload_reads(ngsid_ctx *ctx, const uint8_t *seq, const uint8_t *qual,
                       const int64_t *offsets, int64_t n_reads);

/* ---- K1: homopolymer compression + (k,w) minimizers ---------------------------------------
 * Replaces: modules/cluster.py:265 (groupby compression) and cluster.get_kmer_minimizers
 * (modules/cluster.py:16-39) for every uploaded read. Results stay on the device.
 * k <= 15 (2k-bit code in a u32), k <= w <= 100.                                               */
int ngsid_minimizers(ngsid_ctx *ctx, int k, int w);
/* Timed variant for benchmarks: runs K1 `iters` times back to back on the context's stream and
 * returns the average kernel duration (CUDA events on that stream) in milliseconds.            */
int ngsid_minimizers_timed(ngsid_ctx *ctx, int k, int w, int iters, float *avg_ms);
/* Copy back results for reads [begin, end): len_c and counts have end-begin entries; kmer/pos
 * receive the concatenated minimizers (capacity `cap` entries; *n_total gets the number
 * written or needed). kmer is the 2k-bit code, first base most significant; a code with bit 31
 * set encodes a truncated suffix (reads whose compressed length is < w, reference quirk).      */
int ngsid_get_minimizers(ngsid_ctx *ctx, int64_t begin, int64_t end, uint32_t *len_c,
                         uint32_t *counts, uint32_t *kmer, uint32_t *pos, int64_t cap,
                         int64_t *n_total);

/* The string behind a minimizer code of the exception path: a k-mer with a base outside ACGT has the code
 * 1 << 30 | index into a per-context dictionary (equal strings, equal codes; valid until the next
 * upload). Returns the length, the string is NUL terminated.                                       */
int ngsid_kmer_string(ngsid_ctx *ctx, uint32_t code, char *out, int cap);

/* ---- K0: quality statistics ----------------------------------------------------------------
 * Replaces: modules/cluster.py:273-292 (per-run best quality, compressed error rate) and the
 * per-read part of cluster.py:185-188 (mean capped error probability of the raw qualities).
 * phred_p: 128 doubles, the caller's PHRED char -> error probability table
 * (min(10**(-(c-33)/10), 0.79433), computed by the caller so libm differences cannot matter).
 * bucket_thresholds: 14 doubles, smallest double x for which round(x, 2) >= (b+2)/100,
 * b = 0..13 (the caller derives them with its own round()); gives the 15 buckets of
 * cluster.p_shared_minimizer_empirical (cluster.py:356-366).
 * Sums use Neumaier compensation in ascending character order (what Python >= 3.12's sum()
 * does), so the doubles equal the reference's.                                                 */
int ngsid_quality_stats(ngsid_ctx *ctx, const double *phred_p, const double *bucket_thresholds);
int ngsid_get_quality_stats(ngsid_ctx *ctx, int64_t begin, int64_t end, double *err_compressed,
                            double *err_raw, uint8_t *bucket);

/* ---- K2+K3+K4: the greedy clustering pass ---------------------------------------------------
 * Replaces: cluster.reads_to_clusters (modules/cluster.py:207-353) including get_all_hits
 * (:43-62), get_best_cluster (:67-127), get_best_cluster_block_align (:172-205) and
 * parasail_block_alignment (:130-169).                                                         */
typedef struct {
    int32_t k, w;
    int32_t min_shared;                 /* --min_shared          (5)   */
    int32_t symmetric;                  /* --symmetric_map_align_thresholds */
    double min_fraction;                /* --min_fraction        (0.8) */
    double mapped_threshold;            /* --mapped_threshold    (0.7) */
    double aligned_threshold;           /* --aligned_threshold   (0.4) */
    /* max_gap[b1*15+b2]: largest g with (1-p_emp[b1][b2])^g (left-to-right product from 1)
     * not < min_prob_no_hits; derived by the caller from the probability table
     * (modules/p_minimizers_shared.py, NGSpeciesID:72-77) and --min_prob_no_hits.              */
    int32_t max_gap[225];
    int32_t tile_reads;                 /* speculation tile size, 0 = default                  */
    int32_t reserved[7];
} ngsid_cluster_params;

typedef struct {
    int64_t n_processed, n_new_reps, n_mapped, n_aln_called, n_aln_passed, n_alignments;
    int64_t n_tiles, n_chain_steps, n_surprises, n_map_launch_reads, align_cells;
    int64_t reserved[5];
} ngsid_cluster_stats;

/* order:      n_order read indices (into the uploaded set) in processing order.
 * init_reps:  n_init read indices whose minimizers pre-populate the table (merge rounds of
 *             modules/parallelize.py: the lower batch's representatives); may be NULL/0.
 * acc_rank:   per uploaded read, rank of its accession string (incl. score suffix) in
 *             ascending lexicographic order -- the tie-break of cluster.py:79.
 * out_assign: per entry of `order`: the read index of the representative it joined, or -1 when
 *             it became a representative itself, or -2 when skipped (compressed length < k).
 * out_via:    optional (may be NULL): 0 = new, 1 = mapping, 2 = alignment.                     */
int ngsid_cluster(ngsid_ctx *ctx, const ngsid_cluster_params *params, const int32_t *order,
                  int64_t n_order, const int32_t *init_reps, int64_t n_init,
                  const uint32_t *acc_rank, int32_t *out_assign, uint8_t *out_via,
                  ngsid_cluster_stats *stats);

/* ---- K2 alone: the hit table ------------------------------------------------------------------
 * Replaces: cluster.get_all_hits (modules/cluster.py:43-62) for every (read, representative) pair: a
 * table is built from the minimizers of `reps` (a k-mer once per representative, cluster.py:330-334),
 * every minimizer of reads[i] is looked up (a read does not hit itself) and out_count[i * n_reps + r] /
 * out_possum[i * n_reps + r] receive the number of hits and the sum of the hit positions -- the keys of
 * the ranking at cluster.py:79. Inspection / test entry; ngsid_cluster does the same inside its pass. */
int ngsid_hit_counts(ngsid_ctx *ctx, const int32_t *reps, int64_t n_reps, const int32_t *reads, int64_t n,
                     uint32_t *out_count, uint32_t *out_possum);

/* ---- K4 alone: semi-global block alignment statistic ----------------------------------------
 * Replaces: cluster.parasail_block_alignment (modules/cluster.py:130-169): parasail
 * sg_trace_scan_16 (match 2, mismatch -2, gap open `open`, extend 1), CIGAR expansion and the
 * k-column window count. Pairs index uploaded reads: s1 = read_a (rows), s2 = read_b (columns).
 * out_count[i] = number of windows with >= match_id[i] matches; out_score[i] (optional) = score.
 * The caller divides by len(s1) / len(s2) (cluster.py:167-168).                                */
int ngsid_sg_block_align(ngsid_ctx *ctx, const int32_t *read_a, const int32_t *read_b,
                         const int32_t *open, const int32_t *match_id, int64_t n_pairs, int k,
                         int32_t *out_count, int32_t *out_score);

/* ---- K4 paths: alignment score / identity / window breaking points ---------------------------
 * Replaces: consensus.parasail_alignment + highest_aln_identity (modules/consensus.py:58-73,
 * 129-145: identity = equal columns / all columns of the semi-global alignment, open 3) and the
 * read-to-draft mapping that run_racon obtains from minimap2 (modules/consensus.py:121), in the
 * form racon consumes it: per 500-base window of the target the first/last aligned (=/X) column.
 * a[i] is the row sequence (s1), b[i] the column sequence (s2); an index >= 0 names an uploaded
 * read, an index < 0 names auxiliary sequence -index-1 of (aux_seq, aux_off) (host buffers, e.g.
 * consensus strings). out_win (optional) receives 16 windows x 4 int32 per pair:
 * (q_first, q_last_exclusive, t_first, t_last_exclusive), -1 when the window has no aligned column. */
int ngsid_sg_align_paths(ngsid_ctx *ctx, const int32_t *a, const int32_t *b, const int32_t *open,
                         int64_t n_pairs, const uint8_t *aux_seq, const int64_t *aux_off, int64_t n_aux,
                         int window, int32_t *out_score, int32_t *out_match, int32_t *out_cols,
                         int32_t *out_win);

/* ---- sort stage (the step in front of the clustering path) --------------------------------
 * Replaces the arithmetic of modules/get_sorted_fastq_for_cluster.py:23-33,140-152 for the uploaded
 * reads: out_score[r] = (1 - E[erroneous k-mers] / n) * n, n = len - k + 1, computed with the
 * reference's operation order in IEEE double (bit-identical), with phred_p_capped[c] =
 * min(10^(-(c-33)/10), 0.79433); out_err_rate[r] = sum_c count(c) * phred_p_uncapped[c] / len
 * (terms in ascending character order, Python's compensated sum). Filtering (len < 2k, compressed
 * length < k, mean quality threshold), the stable descending sort and the score suffix of the
 * read names stay on the host (ngspeciesid_b200/modules/get_sorted_fastq_for_cluster.py).      */
int ngsid_sort_scores(ngsid_ctx *ctx, int k, const double *phred_p_capped, const double *phred_p_uncapped,
                      double *out_score, double *out_err_rate);

/* ---- FASTA/FASTQ ingest (host code, no context, no GPU) --------------------------------------
 * Replaces: modules/help_functions.py:13-42 (readfq, lh3's generator) for a whole file held in
 * memory, with the line model of Python's text mode 'r' that the reference opens its input with
 * (get_sorted_fastq_for_cluster.py:126, NGSpeciesID:54): "\n", "\r\n" and "\r" end a line.
 * Record i: name = buf[name_off[i], +name_len[i]) (the whole header line after '@' / '>'),
 * sequence = seq_out[seq_off[i], seq_off[i+1]), quality = qual_out[qual_off[i], qual_off[i+1])
 * if has_qual[i] (0 = FASTA record / quality cut short by the end of the file: the reference
 * yields None). seq_out / qual_out need `len` bytes, the per-record arrays cap_records (+1 for the
 * offsets). With seq_out == NULL only *n_records is computed. Returns NGSID_EINVAL (and the needed
 * count in *n_records) when cap_records is too small. The reference's quirks are kept: the last
 * line of a file without a final newline loses its last character.                              */
int ngsid_fastq_parse(const uint8_t *buf, int64_t len, int64_t cap_records,
                      uint8_t *seq_out, uint8_t *qual_out, int64_t *name_off, int32_t *name_len,
                      int64_t *seq_off, int64_t *qual_off, uint8_t *has_qual, int64_t *n_records);

/* ---- K5: partial-order-alignment consensus ---------------------------------------------------
 * Replaces: the spoa call of consensus.run_spoa (modules/consensus.py:83-92: local alignment,
 * match 5, mismatch -4, linear gap -2, quality weights, heaviest-bundle consensus) and the
 * per-window POA inside racon (consensus.run_racon, modules/consensus.py:107-126: global,
 * 3 / -5 / -4, backbone without weight, coverage-trimmed consensus).
 * Job j consists of layers [job_off[j], job_off[j+1]) added in that order; layer l is bases
 * [layer_begin[l], layer_begin[l]+layer_len[l]) of uploaded read layer_src[l] (weights =
 * quality - 33) or, when layer_src[l] < 0, of auxiliary sequence -layer_src[l]-1 (weight 0).
 * All DP cells and the traceback run on the GPU, one launch per layer step over every job that
 * still has a layer (one thread block per job); the graph of a job -- adding the alignment and
 * spoa's depth-first topological re-sort, O(V + E) pointer chasing -- is kept on host threads
 * inside the library. Graphs grow as needed: max_nodes > 0 makes a larger graph an error
 * (NGSID_EUNSUPPORTED), 0 = no bound. Layers are limited to 4095 bases.
 * out_seq holds n_jobs rows of out_stride bytes; out_len[j] = consensus length.                   */
typedef struct {
    int32_t mode;        /* 0 local (spoa -l 0), 1 global (racon windows) */
    int32_t match, mismatch, gap;
    int32_t trim;        /* racon's coverage trimming of the consensus ends */
    int32_t max_nodes;
    int32_t reserved[2];
} ngsid_poa_params;

int ngsid_poa_consensus(ngsid_ctx *ctx, const ngsid_poa_params *params, int64_t n_jobs,
                        const int64_t *job_off, const int32_t *layer_src, const int32_t *layer_begin,
                        const int32_t *layer_len, const uint8_t *aux_seq, const int64_t *aux_off,
                        int64_t n_aux, uint8_t *out_seq, int64_t out_stride, int32_t *out_len,
                        int32_t *out_nodes);

/* The same with racon's treatment of window layers that do not span their window (racon window.cpp
 * generate_consensus, behind consensus.run_racon, modules/consensus.py:107-126): layer l with
 * layer_sub_begin[l] >= 0 is aligned only to the sub-graph between the backbone positions
 * layer_sub_begin[l] .. layer_sub_end[l] (inclusive; the backbone is the job's first layer) -- the nodes
 * reached from backbone node `end` over in-edges and aligned nodes with id >= `begin` (spoa
 * Graph::subgraph) -- and then added to the whole graph; -1 = the whole graph. Both arrays NULL =
 * ngsid_poa_consensus. */
int ngsid_poa_consensus_sub(ngsid_ctx *ctx, const ngsid_poa_params *params, int64_t n_jobs,
                            const int64_t *job_off, const int32_t *layer_src, const int32_t *layer_begin,
                            const int32_t *layer_len, const int32_t *layer_sub_begin, const int32_t *layer_sub_end,
                            const uint8_t *aux_seq, const int64_t *aux_off,
                            int64_t n_aux, uint8_t *out_seq, int64_t out_stride, int32_t *out_len,
                            int32_t *out_nodes);

/* ---- multi-GPU data plane (one process per GPU, NCCL over NVLink / NVSwitch) --------------------
 * Replaces: the exchange of the reference's --t N mode -- modules/parallelize.py:153-187, where the
 * process pool returns (clusters, representatives, minimizer_database) of every batch to the parent
 * as pickles -- and the per-cluster read files of the consensus step (modules/consensus.py:249-278,
 * 186-246). NCCL is loaded at run time (libnccl.so.2, or $NGSID_NCCL_LIB); a context without a
 * communicator behaves as a world of one rank, so the same driver code runs on one GPU.
 *
 * ngsid_nccl_unique_id: rank 0 creates the id (returns its size, <= cap; 128 bytes) and hands it
 * to the other ranks out of band (file, MPI, torch.distributed, ...). ngsid_nccl_init: collective. */
int ngsid_nccl_unique_id(uint8_t *out, int64_t cap);
int ngsid_nccl_init(ngsid_ctx *ctx, const uint8_t *unique_id, int rank, int nranks);
int ngsid_nccl_finalize(ngsid_ctx *ctx);
/* A further context on the same GPU borrows the communicator of `owner` (which must outlive it). */
int ngsid_nccl_share(ngsid_ctx *ctx, ngsid_ctx *owner);
/* Variable-size all-gather of host bytes (accessions, consensus strings): recv = rank 0's bytes |
 * rank 1's | ...; counts[r] = bytes of rank r (nranks entries). With recv == NULL only the counts
 * are exchanged (every rank must then make the same second call with a buffer).                   */
int ngsid_allgather_bytes(ngsid_ctx *ctx, const uint8_t *send, int64_t n_send, uint8_t *recv,
                          int64_t recv_cap, int64_t *counts);
/* In-place all-reduce of a host array of int32 / int64 (elem_bytes 4 / 8); op 0 = sum, 1 = max.  */
int ngsid_allreduce(ngsid_ctx *ctx, void *buf, int64_t n, int elem_bytes, int op);
/* Collective. Every rank names its surviving representatives (read indices of ctx, processing
 * order). Afterwards dst (another context on the same GPU) holds the representatives of ALL ranks,
 * rank 0's first, as an uploaded read set together with their K1 results (minimizer records,
 * counts, compressed lengths) and K0 results (error rates, buckets), moved device to device:
 * ngsid_cluster(dst, ...) runs the merge rounds of modules/parallelize.py:196-215 on them without
 * recomputing anything. counts[r] = representatives that came from rank r.                       */
int ngsid_gather_representatives(ngsid_ctx *ctx, const int32_t *reps, int64_t n_reps, ngsid_ctx *dst,
                                 int64_t *counts);
/* Collective all-to-all of reads (bases + qualities): read read_idx[i] of ctx goes to rank dest[i]
 * with the caller's 64-bit tag[i]; entries grouped by ascending dest. Afterwards dst holds the
 * reads this rank received as an uploaded read set (rank 0's sends first, in the order listed),
 * out_tag their tags (capacity tag_cap), recv_counts[r] the number that came from rank r.         */
int ngsid_exchange_reads(ngsid_ctx *ctx, const int32_t *read_idx, const int32_t *dest, const int64_t *tag,
                         int64_t n_send, ngsid_ctx *dst, int64_t *out_tag, int64_t tag_cap, int64_t *recv_counts);
/* Doubles the read set: read n + i = reverse complement of read i, qualities reversed (the
 * polishing step aligns the reads of a reverse-complement-merged centre in both orientations,
 * modules/consensus.py:148-183).                                                                 */
int ngsid_append_revcomp(ngsid_ctx *ctx);
/* Host copy of reads [begin, end) of a context (offsets: end-begin+1 entries, first = 0).         */
int ngsid_download_reads(ngsid_ctx *ctx, int64_t begin, int64_t end, uint8_t *seq, uint8_t *qual, int64_t *offsets);

#ifdef __cplusplus
}
#endif
#endif /* NGSID_H */
