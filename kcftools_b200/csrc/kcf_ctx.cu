// kcf_ctx.cu — context, error plumbing, pinned host memory, and K2: the FASTA -> 2-bit packer.
//
// K2 replaces FastaIndex.getSequence (FastaIndex.java:122-182) + the validity / upper-casing logic of
// Fasta.getKmersList (Fasta.java:96-104, 132-134): base p of a sequence lives at byte
// (p / lineBases) * lineWidth + p % lineBases of the mapped slice (FastaIndex.java:147-152); a base is
// valid iff it is one of ACGTacgt.  The whole sequence is converted once; windows then address bases.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include "kcf_internal.cuh"

static std::mutex g_err_mu;
static std::string g_init_err;

int kcf_fail(kcf_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else {
        std::lock_guard<std::mutex> lk(g_err_mu);
        g_init_err = buf;
    }
    return code;
}

extern "C" const char *kcf_version(void) { return "kcf-b200 0.1.0 (sm_100a)"; }

extern "C" const char *kcf_last_error(kcf_ctx *ctx)
{
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    return g_init_err.c_str();
}

extern "C" int kcf_init(int device, kcf_ctx **out)
{
    if (!out) return KCF_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return kcf_fail(nullptr, KCF_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return kcf_fail(nullptr, KCF_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return kcf_fail(nullptr, KCF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10)
        return kcf_fail(nullptr, KCF_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    kcf_ctx *ctx = new kcf_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->desc_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ctx->raw_free[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->h2d_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->plan_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_flags, 64 * sizeof(uint32_t));
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e != cudaSuccess) {
        int rc = kcf_fail(nullptr, KCF_ERR_CUDA, "context setup: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return KCF_OK;
}

// device memory of cleared sequences and destroyed plans is kept and handed out again: a host that re-screens (or a cohort run that swaps
// references) does not pay cudaMalloc / cudaFree per sequence
void *kcf_pool_get(kcf_ctx *ctx, size_t bytes)
{
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); ++i)
        if (ctx->pool[i].bytes >= bytes && ctx->pool[i].bytes <= bytes + bytes / 8 + 4096 &&
            (best < 0 || ctx->pool[i].bytes < ctx->pool[best].bytes))
            best = (int)i;
    if (best >= 0) {
        void *p = ctx->pool[best].p;
        ctx->pool.erase(ctx->pool.begin() + best);
        return p;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        // give the pooled blocks back to the driver and try once more
        (void)cudaGetLastError(); // the failed allocation must not surface as the "last error" of a later launch check
        cudaStreamSynchronize(ctx->stream);
        for (auto &b : ctx->pool) cudaFree(b.p);
        ctx->pool.clear();
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
    }
    return p;
}

void kcf_pool_trim(kcf_ctx *ctx)
{
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->pool) cudaFree(b.p);
    ctx->pool.clear();
}

void kcf_pool_put(kcf_ctx *ctx, void *p, size_t bytes)
{
    if (p) ctx->pool.push_back({p, bytes});
}

extern "C" int kcf_ref_clear(kcf_ctx *ctx)
{
    if (!ctx) return KCF_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->seqs) {
        const size_t n_words = (size_t)((s.len + 31) / 32 + 2);
        if (s.codes) ctx->pool.push_back({s.codes, n_words * 8});
        if (s.valid) ctx->pool.push_back({s.valid, n_words * 4});
    }
    ctx->seqs.clear();
    ctx->seqs_dirty = true;
    ctx->seqs_uploaded = 0;
    ++ctx->ref_generation; // plans built before this point address recycled memory: kcf_plan_run refuses them
    return KCF_OK;
}

extern "C" void kcf_shutdown(kcf_ctx *ctx)
{
    if (!ctx) return;
    kcf_ref_clear(ctx);
    for (auto &b : ctx->pool) cudaFree(b.p);
    if (ctx->d_seqs) cudaFree(ctx->d_seqs);
    if (ctx->h_seqs) cudaFreeHost(ctx->h_seqs);
    for (int i = 0; i < 2; ++i) {
        if (ctx->d_raw[i]) cudaFree(ctx->d_raw[i]);
        if (ctx->raw_free[i]) cudaEventDestroy(ctx->raw_free[i]);
        if (ctx->h2d_done[i]) cudaEventDestroy(ctx->h2d_done[i]);
    }
    if (ctx->plan_ready) cudaEventDestroy(ctx->plan_ready);
    for (int i = 0; i < 8; ++i) {
        if (ctx->ing_h[i]) cudaFreeHost(ctx->ing_h[i]);
        if (ctx->ing_d[i]) cudaFree(ctx->ing_d[i]);
        if (ctx->ing_free[i]) cudaEventDestroy(ctx->ing_free[i]);
        if (ctx->ing_copied[i]) cudaEventDestroy(ctx->ing_copied[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->ing_proved[i]) cudaEventDestroy(ctx->ing_proved[i]);
        if (ctx->ing_inserted[i]) cudaEventDestroy(ctx->ing_inserted[i]);
    }
    if (ctx->ing_stream) cudaStreamDestroy(ctx->ing_stream);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
    for (int i = 0; i < 4; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->desc_stream) cudaStreamDestroy(ctx->desc_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void *kcf_stream(kcf_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int kcf_host_alloc(kcf_ctx *ctx, uint64_t n_bytes, void **out)
{
    if (!ctx || !out) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    KCF_CUDA(ctx, cudaHostAlloc(out, n_bytes ? n_bytes : 1, cudaHostAllocDefault));
    return KCF_OK;
}

extern "C" void kcf_host_free(kcf_ctx *ctx, void *p)
{
    (void)ctx;
    if (p) cudaFreeHost(p);
}

extern "C" int kcf_set_profiling(kcf_ctx *ctx, int on)
{
    if (!ctx) return KCF_ERR_ARG;
    ctx->profiling = on;
    return KCF_OK;
}

extern "C" int kcf_last_kernel_ms(kcf_ctx *ctx, float *screen_ms, float *finalize_ms)
{
    if (!ctx) return KCF_ERR_ARG;
    if (!ctx->ev_valid) return kcf_fail(ctx, KCF_ERR_ARG, "no profiled run yet (kcf_set_profiling(ctx, 1) then kcf_plan_run)");
    KCF_CUDA(ctx, cudaEventSynchronize(ctx->ev[2]));
    float a = 0, b = 0;
    KCF_CUDA(ctx, cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
    KCF_CUDA(ctx, cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
    if (screen_ms) *screen_ms = a;
    if (finalize_ms) *finalize_ms = b;
    return KCF_OK;
}

// ---- K2 ----------------------------------------------------------------------------------------
// One CTA converts PACK_BASES consecutive bases.  The raw bytes they occupy (bases + line terminators)
// are staged in shared memory with 16-byte loads from a 16-byte aligned window, then each thread turns
// 32 bases into the two bit planes of their codes and one validity word.
#define PACK_THREADS 256
#define PACK_BASES (PACK_THREADS * 32)

__global__ void __launch_bounds__(PACK_THREADS)
kcf_pack_kernel(const uint8_t *__restrict__ raw, uint64_t n_bytes_padded, uint32_t line_bases, uint32_t line_width,
                uint64_t seq_len, uint32_t *__restrict__ codes, uint32_t *__restrict__ valid, uint32_t smem_bytes)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    const uint64_t p0 = (uint64_t)blockIdx.x * PACK_BASES;
    if (p0 >= seq_len) return;
    const uint64_t p1 = min(p0 + (uint64_t)PACK_BASES, seq_len); // exclusive
    const uint64_t b0 = (p0 / line_bases) * line_width + p0 % line_bases;
    const uint64_t b1 = ((p1 - 1) / line_bases) * line_width + (p1 - 1) % line_bases + 1;
    const uint64_t a0 = b0 & ~15ULL;
    // cooperative 128-bit loads of [a0, b1) (the device buffer is padded to a multiple of 16 bytes)
    const uint32_t n16 = (uint32_t)((b1 - a0 + 15) >> 4);
    const uint4 *src = reinterpret_cast<const uint4 *>(raw + a0);
    uint4 *dst = reinterpret_cast<uint4 *>(s_raw);
    for (uint32_t i = threadIdx.x; i < n16; i += PACK_THREADS)
        if (a0 + 16ULL * i < n_bytes_padded && 16u * i + 16u <= smem_bytes) dst[i] = src[i];
    __syncthreads();
    const uint64_t p = p0 + 32ULL * threadIdx.x;
    if (p >= seq_len) return;
    uint32_t line = (uint32_t)(p / line_bases);
    uint32_t col = (uint32_t)(p % line_bases);
    uint32_t off = (uint32_t)((uint64_t)line * line_width + col - a0);
    uint32_t pl0 = 0, pl1 = 0, v = 0;
    const uint32_t nb = (uint32_t)min((uint64_t)32, seq_len - p);
    for (uint32_t j = 0; j < nb; ++j) {
        uint32_t b = s_raw[off];
        uint32_t u = b & 0xDFu; // Character.toUpperCase for ASCII letters (Fasta.java:98)
        uint32_t ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T'); // Fasta.java:132-134
        uint32_t c = (((b >> 1) & 3u) ^ ((b >> 2) & 1u)) & (0u - ok);    // A0 C1 G2 T3 (Kmer.java:286-294)
        pl0 |= (c & 1u) << j;
        pl1 |= (c >> 1) << j;
        v |= ok << j;
        ++col;
        ++off;
        if (col == line_bases) { // FastaIndex.java:169 — skip the line terminator
            col = 0;
            off += line_width - line_bases;
        }
    }
    const uint64_t w = p >> 5;
    reinterpret_cast<uint2 *>(codes)[w] = make_uint2(pl0, pl1);
    valid[w] = v;
}

extern "C" int kcf_ref_add_async(kcf_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, uint32_t line_bases, uint32_t line_width,
                                 uint64_t seq_len, int *seq_id_out)
{
    if (!ctx || (!bytes && n_bytes)) return KCF_ERR_ARG;
    if (seq_len >= (1ULL << 31)) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "sequence length %llu exceeds the reference's int range", (unsigned long long)seq_len);
    if (seq_len > 0 && (line_bases == 0 || line_width < line_bases))
        return kcf_fail(ctx, KCF_ERR_ARG, "bad .faidx line geometry (lineBases=%u lineWidth=%u)", line_bases, line_width);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (seq_len > 0) {
        uint64_t last = ((seq_len - 1) / line_bases) * (uint64_t)line_width + (seq_len - 1) % line_bases;
        if (last >= n_bytes) return kcf_fail(ctx, KCF_ERR_FASTA, "sequence bytes end before base %llu", (unsigned long long)(seq_len - 1));
    }
    KcfSeqHost s;
    s.len = seq_len;
    s.n_bytes = n_bytes;
    s.line_bases = line_bases;
    s.line_width = line_width;
    const uint64_t n_words = (seq_len + 31) / 32 + 2; // +2: the tile loader may read one word past the end
    s.codes = (uint32_t *)kcf_pool_get(ctx, n_words * 8);
    s.valid = (uint32_t *)kcf_pool_get(ctx, n_words * 4);
    if (!s.codes || !s.valid) {
        if (s.codes) cudaFree(s.codes);
        if (s.valid) cudaFree(s.valid);
        return kcf_fail(ctx, KCF_ERR_NOMEM, "device memory for a %llu-base sequence", (unsigned long long)seq_len);
    }
    // the pack kernel writes every word that holds a base; only the slack words need clearing
    const uint64_t full = seq_len / 32;
    cudaMemsetAsync(s.codes + 2 * full, 0, (n_words - full) * 8, ctx->stream);
    cudaMemsetAsync(s.valid + full, 0, (n_words - full) * 4, ctx->stream);
    if (seq_len > 0) {
        const int b = ctx->raw_next;
        ctx->raw_next ^= 1;
        const uint64_t padded = (n_bytes + 15) & ~15ULL;
        if (ctx->d_raw_cap[b] < padded + 16) {
            cudaEventSynchronize(ctx->raw_free[b]); // the pack kernel that last read this buffer
            if (ctx->d_raw[b]) cudaFree(ctx->d_raw[b]);
            ctx->d_raw[b] = nullptr;
            ctx->d_raw_cap[b] = 0;
            const uint64_t cap = padded + padded / 16 + 16;
            cudaError_t e = cudaMalloc(&ctx->d_raw[b], cap);
            if (e != cudaSuccess) {
                ctx->pool.push_back({s.codes, n_words * 8});
                ctx->pool.push_back({s.valid, n_words * 4});
                return kcf_fail(ctx, KCF_ERR_NOMEM, "cudaMalloc(raw): %s", cudaGetErrorString(e));
            }
            ctx->d_raw_cap[b] = cap;
        }
        cudaStreamWaitEvent(ctx->copy_stream, ctx->raw_free[b], 0);
        cudaError_t e = cudaMemcpyAsync(ctx->d_raw[b], bytes, n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->h2d_done[b], ctx->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->h2d_done[b], 0);
        if (e != cudaSuccess) {
            ctx->pool.push_back({s.codes, n_words * 8});
            ctx->pool.push_back({s.valid, n_words * 4});
            return kcf_fail(ctx, KCF_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
        }
        // shared bytes: the raw span of PACK_BASES bases plus alignment slack
        const uint64_t lines = PACK_BASES / line_bases + 2;
        uint64_t smem = PACK_BASES + lines * (line_width - line_bases) + 48;
        smem = (smem + 15) & ~15ULL;
        if (smem > 200 * 1024) {
            ctx->pool.push_back({s.codes, n_words * 8});
            ctx->pool.push_back({s.valid, n_words * 4});
            return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "line geometry needs %llu B of shared memory", (unsigned long long)smem);
        }
        if (smem > 48 * 1024) cudaFuncSetAttribute(kcf_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const unsigned grid = (unsigned)((seq_len + PACK_BASES - 1) / PACK_BASES);
        kcf_pack_kernel<<<grid, PACK_THREADS, smem, ctx->stream>>>(ctx->d_raw[b], padded, line_bases, line_width, seq_len,
                                                                   s.codes, s.valid, (uint32_t)smem);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ctx->raw_free[b], ctx->stream);
        if (e != cudaSuccess) {
            cudaStreamSynchronize(ctx->stream);
            ctx->pool.push_back({s.codes, n_words * 8});
            ctx->pool.push_back({s.valid, n_words * 4});
            return kcf_fail(ctx, KCF_ERR_CUDA, "pack kernel: %s", cudaGetErrorString(e));
        }
    }
    ctx->seqs.push_back(s);
    ctx->seqs_dirty = true;
    if (seq_id_out) *seq_id_out = (int)ctx->seqs.size() - 1;
    return KCF_OK;
}

extern "C" int kcf_ref_sync(kcf_ctx *ctx)
{
    if (!ctx) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KCF_OK;
}

extern "C" int kcf_ref_add(kcf_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, uint32_t line_bases, uint32_t line_width,
                           uint64_t seq_len, int *seq_id_out)
{
    int rc = kcf_ref_add_async(ctx, bytes, n_bytes, line_bases, line_width, seq_len, seq_id_out);
    if (rc != KCF_OK) return rc;
    return kcf_ref_sync(ctx); // the caller may reuse `bytes` after return
}
