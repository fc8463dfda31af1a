// Store-pattern microbenchmark for the K1 output (sm_100a): 2 M "reads", each owning a slack region
// of 731 records of 8 bytes of which 118 are written (the ONT workload of bench.py).
//   coal   one warp instruction writes 32 consecutive records of ONE read (the record pass of
//          k1_stream_kernel)
//   t8     every thread writes the records of ITS read, one 8-byte store per record
//   t16    ... one 16-byte store per two records
//   t32    ... two 16-byte stores per four records (one full 32-byte sector per thread)
// Blocks of 128 threads with 46 KB of dynamic shared memory (the occupancy of the real kernel).
// Reports the effective write bandwidth of each pattern.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define NREC 118
#define SLOTS 731
// coal with `slots` = 731 (the slack layout of this build), 736 (128-byte aligned regions) and 118
// (dense: no slack) separates the cost of the slack layout from the cost of the store pattern;
// coal16d = dense with 16 bytes per lane (512 bytes per warp instruction).

__global__ void __launch_bounds__(128) k_coal(uint2 *out, long long n_reads, int slots)
{
    const int lane = threadIdx.x & 31;
    const long long r0 = ((long long)blockIdx.x * 128 + (threadIdx.x & ~31));
    for (int q = 0; q < 32; ++q) {
        const long long r = r0 + q;
        if (r >= n_reads) break;
        uint2 *p = out + r * slots;
        for (int i = lane; i < NREC; i += 32) p[i] = make_uint2((uint32_t)r, (uint32_t)i);
    }
}
__global__ void __launch_bounds__(128) k_coal16d(uint4 *out, long long n_reads)
{
    const int lane = threadIdx.x & 31;
    const long long r0 = ((long long)blockIdx.x * 128 + (threadIdx.x & ~31));
    for (int q = 0; q < 32; ++q) {
        const long long r = r0 + q;
        if (r >= n_reads) break;
        uint4 *p = out + r * (NREC / 2);
        for (int i = lane; i < NREC / 2; i += 32) p[i] = make_uint4((uint32_t)r, (uint32_t)i, (uint32_t)r, (uint32_t)i);
    }
}
__global__ void __launch_bounds__(128) k_t8(uint2 *out, long long n_reads)
{
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= n_reads) return;
    uint2 *p = out + r * SLOTS;
#pragma unroll 2
    for (int i = 0; i < NREC; ++i) p[i] = make_uint2((uint32_t)r, (uint32_t)i);
}
__global__ void __launch_bounds__(128) k_t16(uint4 *out, long long n_reads)
{
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= n_reads) return;
    uint4 *p = out + r * (SLOTS + 1) / 2;      // 16-byte aligned regions
#pragma unroll 2
    for (int i = 0; i < NREC / 2; ++i) p[i] = make_uint4((uint32_t)r, (uint32_t)i, (uint32_t)r, (uint32_t)i + 1);
}
__global__ void __launch_bounds__(128) k_t32(uint4 *out, long long n_reads)
{
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= n_reads) return;
    uint4 *p = out + r * ((SLOTS + 3) / 4 * 2);  // 32-byte aligned regions
#pragma unroll 2
    for (int i = 0; i < NREC / 4; ++i) {
        p[2 * i] = make_uint4((uint32_t)r, (uint32_t)i, (uint32_t)r, (uint32_t)i + 1);
        p[2 * i + 1] = make_uint4((uint32_t)r, (uint32_t)i + 2, (uint32_t)r, (uint32_t)i + 3);
    }
}
// like t8, but each thread also does ~12 dependent integer instructions per record (the extraction)
__global__ void __launch_bounds__(128) k_t8alu(uint2 *out, long long n_reads, uint32_t seed)
{
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= n_reads) return;
    uint2 *p = out + r * SLOTS;
    uint32_t x = seed + (uint32_t)r;
#pragma unroll 2
    for (int i = 0; i < NREC; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) x = __funnelshift_l(x, x ^ seed, 5) + (x & 0x55u);
        p[i] = make_uint2(x, (uint32_t)i);
    }
}

template <typename F> static float time_it(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    const long long n = 2000000;
    const size_t bytes = (size_t)n * (SLOTS + 3) * 8 + 4096;
    void *buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, bytes);
    const int blocks = (int)((n + 127) / 128);
    const size_t smem = 46 * 1024;
    cudaFuncSetAttribute(k_coal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_t8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_t16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_t32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_t8alu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double gb = (double)n * NREC * 8 / 1e9;
    float ms;
    cudaFuncSetAttribute(k_coal16d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int sl[3] = {731, 736, 118};
    for (int v = 0; v < 3; ++v) {
        ms = time_it([&] { k_coal<<<blocks, 128, smem>>>((uint2 *)buf, n, sl[v]); });
        printf("coal slots=%d   %.3f ms  %.0f GB/s\n", sl[v], ms, gb / ms * 1e3);
    }
    ms = time_it([&] { k_coal16d<<<blocks, 128, smem>>>((uint4 *)buf, n); });
    printf("coal16d %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_it([&] { k_t8<<<blocks, 128, smem>>>((uint2 *)buf, n); });
    printf("t8     %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_it([&] { k_t16<<<blocks, 128, smem>>>((uint4 *)buf, n); });
    printf("t16    %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_it([&] { k_t32<<<blocks, 128, smem>>>((uint4 *)buf, n); });
    printf("t32    %.3f ms  %.0f GB/s\n", ms, (double)n * (NREC / 4 * 4) * 8 / 1e9 / ms * 1e3);
    ms = time_it([&] { k_t8alu<<<blocks, 128, smem>>>((uint2 *)buf, n, 12345u); });
    printf("t8alu  %.3f ms  %.0f GB/s  (12 dependent integer instructions per record)\n", ms, gb / ms * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
