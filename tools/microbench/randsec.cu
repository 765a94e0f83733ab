// randsec.cu — random-access HBM microbenchmark (measurement tool, not product code).
// Explores: load width per probe (32/64/128 B), L2 fetch granularity limit, loads in flight per thread,
// CTA size, cache operators.  Prints loads/s and useful GB/s; run under ncu to get dram__bytes_read.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ void ld32(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
    if (OP == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (OP == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (OP == 2) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (OP == 4) { asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p)); b = c = d = 0; }
    if (OP == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// SECT = sectors read per probe (1, 2 or 4 -> 32, 64, 128 bytes), U = independent probes in flight per thread
template <int SECT, int U, int OP>
__global__ void probe(const uint64_t *__restrict__ buf, uint64_t n_units, uint64_t iters, uint64_t seed, uint64_t *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ULL ^ seed;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t v[U][SECT][4];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
            uint64_t s = __umul64hi(x, n_units);
#pragma unroll
            for (int q = 0; q < SECT; ++q) ld32<OP>(buf + (s * SECT + q) * 4, v[j][q][0], v[j][q][1], v[j][q][2], v[j][q][3]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
            for (int q = 0; q < SECT; ++q) acc += v[j][q][0] ^ v[j][q][1] ^ v[j][q][2] ^ v[j][q][3];
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

template <int SECT, int U, int OP>
void run(const char *name, const uint64_t *buf, uint64_t bytes, int threads, int ctas_per_sm, int sms, uint64_t *sink)
{
    uint64_t n_units = bytes / (32 * SECT);
    int grid = sms * ctas_per_sm;
    uint64_t total_threads = (uint64_t)grid * threads;
    uint64_t iters = (1ULL << 28) / (total_threads * U) + 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        probe<SECT, U, OP><<<grid, threads>>>(buf, n_units, iters, 1234567ULL * (r + 1), sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double probes = (double)total_threads * U * iters;
    printf("%-34s sect=%d U=%2d thr=%4d cta/sm=%2d : %7.2f Gprobe/s  %8.1f GB/s useful  (%s)\n", name, SECT, U, threads, ctas_per_sm,
           probes / best / 1e6, probes * 32.0 * SECT / best / 1e6, cudaGetErrorString(e));
    fflush(stdout);
}

int main(int argc, char **argv)
{
    double gibf = argc > 1 ? atof(argv[1]) : 16;
    uint64_t gib = (uint64_t)gibf;
    int gran = argc > 2 ? atoi(argv[2]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
        printf("set L2 fetch granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), v);
    } else {
        size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
        printf("default L2 fetch granularity %zu\n", v);
    }
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint64_t bytes = (uint64_t)(gibf * (1ULL << 30)) & ~4095ULL;
    uint64_t *buf, *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    printf("%s, %d SMs, buffer %llu GiB\n", prop.name, sms, (unsigned long long)gib);
    int mode = argc > 3 ? atoi(argv[3]) : 0;
    if (mode == 0) {
        run<1, 8, 0>("nc.noalloc", buf, bytes, 256, 2, sms, sink);
        run<1, 8, 0>("nc.noalloc", buf, bytes, 256, 4, sms, sink);
        run<1, 8, 0>("nc.noalloc", buf, bytes, 256, 8, sms, sink);
        run<1, 4, 0>("nc.noalloc", buf, bytes, 256, 8, sms, sink);
        run<1, 16, 0>("nc.noalloc", buf, bytes, 256, 4, sms, sink);
        run<1, 8, 0>("nc.noalloc", buf, bytes, 1024, 2, sms, sink);
        run<1, 2, 0>("nc.noalloc", buf, bytes, 1024, 2, sms, sink);
        run<1, 8, 1>("plain", buf, bytes, 256, 8, sms, sink);
        run<1, 8, 2>("cg", buf, bytes, 256, 8, sms, sink);
        run<1, 8, 3>("nc.noalloc.evict_first", buf, bytes, 256, 8, sms, sink);
        run<2, 8, 0>("nc.noalloc 64B", buf, bytes, 256, 4, sms, sink);
        run<2, 4, 0>("nc.noalloc 64B", buf, bytes, 256, 8, sms, sink);
        run<4, 4, 0>("nc.noalloc 128B", buf, bytes, 256, 4, sms, sink);
        run<4, 2, 0>("nc.noalloc 128B", buf, bytes, 256, 8, sms, sink);
        run<4, 4, 0>("nc.noalloc 128B", buf, bytes, 256, 8, sms, sink);
    } else if (mode == 2) {  // size / SM-count sweep
        for (double f : {1.0 / 16, 0.25, 1.0, 4.0, 16.0, 64.0}) {
            uint64_t b = (uint64_t)(f * (1ULL << 30));
            if (b > bytes) break;
            printf("buffer %.3f GiB: ", f);
            run<1, 8, 0>("nc.noalloc", buf, b, 256, 8, sms, sink);
        }
        for (int s : {16, 37, 74, 111, 148}) {
            printf("SMs %3d: ", s);
            run<1, 8, 0>("nc.noalloc", buf, bytes, 256, 8, s, sink);
        }
        printf("LDG.64 instead of LDG.256: ");
        run<1, 8, 4>("u64", buf, bytes, 256, 8, sms, sink);
    } else {  // short list for ncu
        run<1, 8, 0>("nc.noalloc", buf, bytes, 256, 8, sms, sink);
        run<2, 4, 0>("nc.noalloc 64B", buf, bytes, 256, 8, sms, sink);
        run<4, 2, 0>("nc.noalloc 128B", buf, bytes, 256, 8, sms, sink);
    }
    return 0;
}
