// K0s: the sort key of the stage in front of the clustering path (SURVEY.md section 8 f, rank 1).
// Reference: modules/get_sorted_fastq_for_cluster.py:23-33 (expected_number_of_erroneous_kmers) and
// :150-152 (score = (1 - E[erroneous k-mers] / n) * n with n = len - k + 1).
// The reference walks the read once, carrying the probability that the current window of k bases is
// error free:  cur *= (1 - p_new) / (1 - p_leaving);  total += cur  -- one rounding per operation in
// IEEE double. The order of the operations is the result, so a read is one thread and every
// operation is an explicit round-to-nearest intrinsic (the library is built with --fmad=false).
#pragma once
#include "ngsid_internal.cuh"

__global__ void __launch_bounds__(128)
k0s_sortscore_kernel(const uint8_t *__restrict__ qual, const int64_t *__restrict__ off,
                     const double *__restrict__ phred_p, int k, double *__restrict__ score, int64_t n_reads)
{
    __shared__ double ptab[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ptab[i] = phred_p[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint8_t *q = qual + off[r];
    const int L = (int)(off[r + 1] - off[r]);
    if (L < k || k < 1) { score[r] = 0.0; return; }       // never scored by the reference (len < 2k is skipped)
    double cur = 1.0;
    for (int i = 0; i < k; ++i) cur = __dmul_rn(cur, __dsub_rn(1.0, ptab[q[i] & 127]));
    double total = cur;
    for (int i = k; i < L; ++i) {
        const double ratio = __ddiv_rn(__dsub_rn(1.0, ptab[q[i] & 127]), __dsub_rn(1.0, ptab[q[i - k] & 127]));
        cur = __dmul_rn(cur, ratio);
        total = __dadd_rn(total, cur);
    }
    const double n = (double)(L - k + 1);
    const double exp_err = __dsub_rn(n, total);
    const double p_no_err = __dsub_rn(1.0, __ddiv_rn(exp_err, n));
    score[r] = __dmul_rn(p_no_err, n);
}
