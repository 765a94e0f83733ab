// randline.cu — second random-access microbenchmark (measurement tool, not product code).
// Questions it answers for the minimizer-clustered table design:
//   T1  baseline: one random 32-B sector per lane                      (the r1 "random-sector roofline")
//   T2  the same, but all lanes of a warp stay inside one random 2-MB page   -> is the cap translation-bound?
//   T3  one random 128-B line per 4 lanes (4 x LDG.256 in ONE warp instruction)  -> cost per line when coalesced
//   T4  one random 128-B line per cp.async.bulk (TMA 1-D bulk copy into shared memory)
//   T5  sector 0 of a random line, wait for it, then sectors 1..3       -> are later sectors of a filled line L2 hits?
//   T6  four passes over the same 2^18 random lines, pass p reads sector p (L2 flushed before pass 0)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ uint64_t rnd(uint64_t &x)
{
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
    return x;
}

// T1 / T2: MODE 0 = anywhere, MODE 1 = inside a per-warp random 2-MB page chosen per iteration
template <int U, int MODE>
__global__ void k_sector(const uint64_t *__restrict__ buf, uint64_t n_sectors, uint64_t iters, uint64_t seed, uint64_t *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ULL ^ seed;
    uint64_t wx = ((t >> 5) + 1) * 0xD1B54A32D192ED03ULL ^ seed;
    const uint64_t n_pages = n_sectors >> 16; // 2 MB = 65536 sectors
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t v[U][4];
        uint64_t page = 0;
        if (MODE == 1) page = __umul64hi(rnd(wx), n_pages) << 16;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            uint64_t r = rnd(x);
            uint64_t s = MODE == 1 ? page + (r >> 48) : __umul64hi(r, n_sectors);
            ld32(buf + s * 4, v[j][0], v[j][1], v[j][2], v[j][3]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) acc += v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

// T3: groups of 4 lanes read the 4 sectors of one random line in one instruction; U lines in flight per group
template <int U>
__global__ void k_line4(const uint64_t *__restrict__ buf, uint64_t n_lines, uint64_t iters, uint64_t seed, uint64_t *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = ((t >> 2) + 1) * 0x9E3779B97F4A7C15ULL ^ seed; // same stream for the 4 lanes of a group
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t v[U][4];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            uint64_t l = __umul64hi(rnd(x), n_lines);
            ld32(buf + l * 16 + (t & 3) * 4, v[j][0], v[j][1], v[j][2], v[j][3]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) acc += v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

// T4: TMA 1-D bulk copies of one random 128-B line each into shared memory.  Each warp owns LINES slots and one
// mbarrier; ALL = 0: lane 0 issues the LINES copies, ALL = 1: every lane issues LINES/32 copies.
template <int LINES, int ALL>
__global__ void k_bulk(const uint64_t *__restrict__ buf, uint64_t n_lines, uint64_t iters, uint64_t seed, uint64_t *sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t *slots = smem + (size_t)warp * LINES * 128;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + (size_t)nw * LINES * 128) + warp;
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t slots_a = (uint32_t)__cvta_generic_to_shared(slots);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ULL ^ seed;
    uint64_t acc = 0;
    uint32_t phase = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(LINES * 128) : "memory");
        __syncwarp();
        if (ALL) {
#pragma unroll
            for (int j = 0; j < LINES / 32; ++j) {
                uint64_t l = __umul64hi(rnd(x), n_lines);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(
                                 slots_a + (j * 32 + lane) * 128),
                             "l"(buf + l * 16), "r"(mbar_a)
                             : "memory");
            }
        } else if (lane == 0) {
#pragma unroll 4
            for (int j = 0; j < LINES; ++j) {
                uint64_t l = __umul64hi(rnd(x), n_lines);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(slots_a + j * 128),
                             "l"(buf + l * 16), "r"(mbar_a)
                             : "memory");
            }
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done)
                         : "r"(mbar_a), "r"(phase)
                         : "memory");
        }
        phase ^= 1;
        acc += reinterpret_cast<const uint64_t *>(slots)[lane * (LINES * 16 / 32)];
        __syncwarp();
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

// T5: sector 0 first; sectors 1..3 of the same line only after sector 0 has arrived
template <int U>
__global__ void k_dep(const uint64_t *__restrict__ buf, uint64_t n_lines, uint64_t iters, uint64_t seed, uint64_t *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ULL ^ seed;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t v[U][4], l[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            l[j] = __umul64hi(rnd(x), n_lines);
            ld32(buf + l[j] * 16, v[j][0], v[j][1], v[j][2], v[j][3]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            // the buffer holds 0x0101..; the dependency is real but never changes the address
            uint64_t z = (v[j][0] ^ v[j][1]) & 1ULL; // == 0
            acc += v[j][2] ^ v[j][3];
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                uint64_t a, b, c, d;
                ld32(buf + (l[j] + z) * 16 + q * 4, a, b, c, d);
                acc += a ^ b ^ c ^ d;
            }
        }
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

// T6: pass p reads sector p of line perm(i), i < n_sel; lines spread over the whole buffer
__global__ void k_pass(const uint64_t *__restrict__ buf, uint64_t n_lines, uint64_t n_sel, int sector, uint64_t *sink)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (; i < n_sel; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ULL;
        uint64_t l = __umul64hi(rnd(x), n_lines);
        uint64_t a, b, c, d;
        ld32(buf + l * 16 + sector * 4, a, b, c, d);
        acc += a ^ b ^ c ^ d;
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b)
{
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

#define RUN(name, unit_per_thread_iter, launch)                                                     \
    do {                                                                                            \
        float best = 1e30f;                                                                         \
        for (int r = 0; r < 4; ++r) {                                                               \
            uint64_t seed = 1234567ULL * (r + 1);                                                   \
            (void)seed;                                                                             \
            cudaEventRecord(e0);                                                                    \
            launch;                                                                                 \
            cudaEventRecord(e1);                                                                    \
            float ms = time_ms(e0, e1);                                                             \
            if (r > 0 && ms < best) best = ms;                                                      \
        }                                                                                           \
        cudaError_t e = cudaGetLastError();                                                         \
        double units = (double)(unit_per_thread_iter);                                              \
        printf("%-44s %8.3f ms  %8.2f G/s  (%s)\n", name, best, units / best / 1e6, cudaGetErrorString(e)); \
        fflush(stdout);                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    double gibf = argc > 1 ? atof(argv[1]) : 16;
    int only = argc > 2 ? atoi(argv[2]) : 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint64_t bytes = (uint64_t)(gibf * (1ULL << 30)) & ~((2ULL << 20) - 1);
    uint64_t *buf, *sink, *flush;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    const size_t flush_bytes = 512ULL << 20;
    cudaMalloc(&flush, flush_bytes);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%s, %d SMs, buffer %.1f GiB\n", prop.name, sms, gibf);
    const uint64_t n_sectors = bytes / 32, n_lines = bytes / 128;
    const int thr = 256, grid = sms * 8;
    const uint64_t total_threads = (uint64_t)grid * thr;
    const uint64_t N = 1ULL << 28;

    if (!only || only == 1) {
        uint64_t it = N / (total_threads * 8) + 1;
        RUN("T1 random sector, U=8 [sectors]", total_threads * 8 * it, (k_sector<8, 0><<<grid, thr>>>(buf, n_sectors, it, seed, sink)));
    }
    if (!only || only == 2) {
        uint64_t it = N / (total_threads * 8) + 1;
        RUN("T2 sectors inside a per-warp 2MB page, U=8", total_threads * 8 * it, (k_sector<8, 1><<<grid, thr>>>(buf, n_sectors, it, seed, sink)));
        it = N / (total_threads * 16) + 1;
        RUN("T2 sectors inside a per-warp 2MB page, U=16", total_threads * 16 * it, (k_sector<16, 1><<<grid, thr>>>(buf, n_sectors, it, seed, sink)));
    }
    if (!only || only == 3) {
        uint64_t it = N / (total_threads * 8) + 1;
        RUN("T3 line per 4 lanes, U=8 [lines]", total_threads / 4 * 8 * it, (k_line4<8><<<grid, thr>>>(buf, n_lines, it, seed, sink)));
        it = N / (total_threads * 4) + 1;
        RUN("T3 line per 4 lanes, U=4 [lines]", total_threads / 4 * 4 * it, (k_line4<4><<<grid, thr>>>(buf, n_lines, it, seed, sink)));
        it = N / (total_threads * 16) + 1;
        RUN("T3 line per 4 lanes, U=16 [lines]", total_threads / 4 * 16 * it, (k_line4<16><<<grid, thr>>>(buf, n_lines, it, seed, sink)));
    }
    if (!only || only == 4) {
        {
            constexpr int LINES = 32;
            const int bthr = 128, nw = bthr / 32;
            const size_t sm = (size_t)nw * LINES * 128 + nw * 8;
            for (int cps : {4, 8, 12}) {
                const int g = sms * cps;
                uint64_t it = (N / 4) / ((uint64_t)g * nw * LINES) + 1;
                cudaFuncSetAttribute(k_bulk<LINES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaFuncSetAttribute(k_bulk<LINES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                char nm[96];
                snprintf(nm, sizeof nm, "T4 bulk 128B, lane0 issues 32, %d CTA/SM [lines]", cps);
                RUN(nm, (uint64_t)g * nw * LINES * it, (k_bulk<LINES, 0><<<g, bthr, sm>>>(buf, n_lines, it, seed, sink)));
                snprintf(nm, sizeof nm, "T4 bulk 128B, all lanes issue 1, %d CTA/SM [lines]", cps);
                RUN(nm, (uint64_t)g * nw * LINES * it, (k_bulk<LINES, 1><<<g, bthr, sm>>>(buf, n_lines, it, seed, sink)));
            }
        }
        {
            constexpr int LINES = 64;
            const int bthr = 128, nw = bthr / 32;
            const size_t sm = (size_t)nw * LINES * 128 + nw * 8;
            const int cps = 6, g = sms * cps;
            uint64_t it = (N / 4) / ((uint64_t)g * nw * LINES) + 1;
            cudaFuncSetAttribute(k_bulk<LINES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            RUN("T4 bulk 128B, all lanes issue 2, 6 CTA/SM [lines]", (uint64_t)g * nw * LINES * it, (k_bulk<LINES, 1><<<g, bthr, sm>>>(buf, n_lines, it, seed, sink)));
        }
    }
    if (!only || only == 5) {
        uint64_t it = (N / 4) / (total_threads * 4) + 1;
        RUN("T5 sector0 then sectors1-3 (dependent), U=4 [lines]", total_threads * 4 * it, (k_dep<4><<<grid, thr>>>(buf, n_lines, it, seed, sink)));
        it = (N / 4) / (total_threads * 8) + 1;
        RUN("T5 sector0 then sectors1-3 (dependent), U=8 [lines]", total_threads * 8 * it, (k_dep<8><<<grid, thr>>>(buf, n_lines, it, seed, sink)));
    }
    if (!only || only == 6) {
        for (uint64_t n_sel : {1ULL << 18, 1ULL << 20}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(flush, rep, flush_bytes); // evict the lines from L2
                cudaDeviceSynchronize();
                for (int p = 0; p < 4; ++p) {
                    cudaEventRecord(e0);
                    k_pass<<<sms * 8, 256>>>(buf, n_lines, n_sel, p, sink);
                    cudaEventRecord(e1);
                    float ms = time_ms(e0, e1);
                    printf("T6 n_sel=2^%d rep %d pass %d (sector %d): %8.4f ms  %8.2f G/s\n", n_sel == (1ULL << 18) ? 18 : 20, rep, p, p, ms,
                           n_sel / ms / 1e6);
                }
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("final: %s\n", cudaGetErrorString(e));
    return 0;
}
