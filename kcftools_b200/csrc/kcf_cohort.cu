// kcf_cohort.cu — the windows x samples matrix that getVariations results feed, resident in HBM:
//
//   cohort      Plugins/Cohort.java:71-119            columns = samples; filled straight from device results (kcf_plan, no
//                                                     KCF text in between) or from rows parsed by the host's KCF reader
//   scores      Data/Window.java:42-83 -> Data.java:41-67, 95-107   every cell's score recomputed from its integers and
//                                                     the header weights, as the reference does when it reads a KCF
//   findIBS     Plugins/FindIBS.java:118-158          block numbering per sample over the windows in traversal order
//   kcf2gt      Plugins/KCFToGenotypeTable.java:116-133, 159-172     allele codes and the per-window filter
//
// Layout: cells[sample][window] (one 40-byte kcf_cell_t each) — a sample's column is contiguous, so appending a sample is
// one strided pass over a plan's rows and the findIBS walk of one sample is a coalesced stream; total_kmers / eff_len are
// per window.  All of it is small next to the screening path (3e5 windows x 64 samples = 0.8 GB); these kernels are
// bandwidth-trivial and exist so that the numbers never leave the device between getVariations and the tables.
#include <algorithm>
#include <vector>
#include "kcf_internal.cuh"

struct kcf_cohort {
    kcf_ctx *ctx = nullptr;
    uint64_t n_windows = 0;
    uint32_t n_samples = 0;
    int32_t *d_total = nullptr, *d_eff = nullptr; // per window
    kcf_cell_t *d_cells = nullptr;                // [n_samples][n_windows]
    uint32_t *d_flags = nullptr;                  // [0] window totals differ between samples, [1] a score was computed
    std::vector<uint8_t> filled;                  // per sample
    bool totals_set = false;
};

enum { CF_MISMATCH = 0, CF_SCORE_USED = 1, CF_COUNT = 4 };

// Data.computeScore (Data.java:95-107): IEEE double, left to right, no fused multiply-add
__device__ __forceinline__ double kcf_cell_score(int32_t obs, int32_t total, int32_t eff, int32_t inner, int32_t left, int32_t right, double wi,
                                                 double wt, double wr, uint32_t *flags)
{
    if (obs == 0 || total == 0 || eff == 0) return 0.0;
    atomicOr(&flags[CF_SCORE_USED], 1u);
    const double e = (double)eff;
    const double ta = __dmul_rn(wr, __ddiv_rn((double)obs, (double)total));
    const double tb = __dmul_rn(wi, __dsub_rn(1.0, __ddiv_rn((double)inner, e)));
    const double tc = __dmul_rn(wt, __dsub_rn(1.0, __ddiv_rn((double)(left + right), e)));
    return __dmul_rn(__dadd_rn(__dadd_rn(ta, tb), tc), 100.0);
}

// one sample column from the rows of a plan (Cohort.java:78-97 without the text round trip)
__global__ void kcf_cohort_from_plan_kernel(const kcf_result_t *__restrict__ res, uint64_t n, kcf_cell_t *__restrict__ cells,
                                            int32_t *__restrict__ total, int32_t *__restrict__ eff, int set_totals, uint32_t *flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const kcf_result_t r = res[i];
    if (set_totals) {
        total[i] = r.total_kmers;
        eff[i] = r.eff_len;
    } else if (total[i] != r.total_kmers || eff[i] != r.eff_len) {
        atomicOr(&flags[CF_MISMATCH], 1u); // the samples were not screened over the same windows
    }
    kcf_cell_t c;
    c.obs = r.obs;
    c.variations = r.variations;
    c.inner = r.inner;
    c.left = r.left;
    c.right = r.right;
    c.ibs = -1; // "N"
    c.kmer_count = r.kmer_count_sum;
    c.score = r.score;
    cells[i] = c;
}

__global__ void kcf_cohort_score_kernel(kcf_cell_t *__restrict__ cells, const int32_t *__restrict__ total, const int32_t *__restrict__ eff,
                                        uint64_t n_windows, uint32_t n_samples, double wi, double wt, double wr, uint32_t *flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_windows * n_samples) return;
    const uint64_t w = i % n_windows;
    kcf_cell_t *c = cells + i;
    c->score = kcf_cell_score(c->obs, total[w], eff[w], c->inner, c->left, c->right, wi, wt, wr, flags);
}

// FindIBS.java:118-158.  One warp per sample walks the windows in traversal order (order[p] = window index, chrom[p] =
// ordinal of its chromosome), 32 positions per step.  The reference's state machine
//     IBS window: first ever -> block 1; else block++ when numNA > minConsecutive or the previous IBS window sits on
//                 another chromosome; numNA = 0           non-IBS window: numNA++ (numNA restarts at 0 per chromosome)
// is a prefix sum over the IBS windows of "previous IBS window exists and (other chromosome or more than minConsecutive
// windows in between)": ballots give every lane its predecessor, popc gives the running block number.
__global__ void __launch_bounds__(128) kcf_cohort_ibs_kernel(kcf_cell_t *__restrict__ cells, uint64_t n_windows, uint32_t n_samples,
                                                             const uint32_t *__restrict__ order, const uint32_t *__restrict__ chrom, uint64_t n,
                                                             int detect_var, int32_t min_consecutive, double cutoff)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_samples) return;
    kcf_cell_t *col = cells + (uint64_t)s * n_windows;
    long long prev_pos = -1; // traversal position of the last IBS window so far (warp uniform)
    uint32_t prev_chrom = 0;
    int32_t block = 0; // block number of the last IBS window so far
    for (uint64_t p0 = 0; p0 < n; p0 += 32) {
        const uint64_t p = p0 + lane;
        const bool in = p < n;
        const uint32_t w = in ? order[p] : 0u, ch = in ? chrom[p] : 0u;
        const double sc = in ? col[w].score : 0.0;
        const bool is = in && (detect_var ? sc < cutoff : sc >= cutoff);
        const uint32_t bits = __ballot_sync(0xffffffffu, is);
        // predecessor of this lane's window among the IBS windows
        const uint32_t below = bits & ((1u << lane) - 1u);
        const int pl = below ? 31 - __clz(below) : -1; // lane of the predecessor inside this step
        const uint32_t pch_lane = __shfl_sync(0xffffffffu, ch, pl < 0 ? 0 : pl);
        const long long ppos = pl >= 0 ? (long long)(p0 + (uint32_t)pl) : prev_pos;
        const uint32_t pch = pl >= 0 ? pch_lane : prev_chrom;
        const bool inc = is && ppos >= 0 && (pch != ch || (long long)p - ppos - 1 > (long long)min_consecutive);
        const uint32_t incb = __ballot_sync(0xffffffffu, inc);
        // the first IBS window ever opens block 1 (FindIBS.java:143-146): every IBS lane of that step counts from there
        const int32_t base = block + (prev_pos < 0 ? 1 : 0);
        if (in) col[w].ibs = is ? base + (int32_t)__popc(incb & ((2u << lane) - 1u)) : -1;
        if (bits) {
            const int last = 31 - __clz(bits);
            block += (int32_t)__popc(incb) + ((prev_pos < 0) ? 1 : 0);
            prev_pos = (long long)(p0 + (uint32_t)last);
            prev_chrom = __shfl_sync(0xffffffffu, ch, last);
        }
    }
}

// KCFToGenotypeTable.java:116-133 (allele codes), :159-172 (badWindow).  One thread per window.
__global__ void kcf_cohort_gt_kernel(const kcf_cell_t *__restrict__ cells, uint64_t n_windows, uint32_t n_samples, double score_a, double score_b,
                                     double score_n, double min_maf, double max_missing, int8_t *__restrict__ alleles, uint8_t *__restrict__ bad)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    int c0 = 0, c1 = 0, c2 = 0, cn = 0;
    for (uint32_t s = 0; s < n_samples; ++s) {
        const double sc = cells[(uint64_t)s * n_windows + w].score;
        int a;
        if (sc >= score_a) { a = 0; ++c0; }
        else if (sc >= score_b) { a = 2; ++c2; }
        else if (sc <= score_n) { a = -1; ++cn; }
        else { a = 1; ++c1; }
        alleles[w * n_samples + s] = (int8_t)a;
    }
    const int n = (int)n_samples, valid = n - cn;
    const bool b = (c0 == n || c1 == n || c2 == n || cn == n) ||
                   (valid > 0 && ((double)c0 <= __dmul_rn(min_maf, (double)valid) || (double)c2 <= __dmul_rn(min_maf, (double)valid))) ||
                   ((double)cn >= __dmul_rn(max_missing, (double)n) || (double)(cn + c1) >= __dmul_rn(max_missing, (double)n));
    bad[w] = b ? 1 : 0;
}

extern "C" void kcf_cohort_destroy(kcf_cohort *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cudaFree(c->d_total);
    cudaFree(c->d_eff);
    cudaFree(c->d_cells);
    cudaFree(c->d_flags);
    delete c;
}

extern "C" int kcf_cohort_create(kcf_ctx *ctx, uint64_t n_windows, uint32_t n_samples, const int32_t *total_kmers, const int32_t *eff_len,
                                 kcf_cohort **out)
{
    if (!ctx || !out || n_samples == 0) return KCF_ERR_ARG;
    if ((total_kmers == nullptr) != (eff_len == nullptr)) return KCF_ERR_ARG;
    *out = nullptr;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    kcf_cohort *c = new kcf_cohort();
    c->ctx = ctx;
    c->n_windows = n_windows;
    c->n_samples = n_samples;
    c->filled.assign(n_samples, 0);
    const uint64_t nw = std::max<uint64_t>(n_windows, 1);
    cudaError_t e = cudaMalloc(&c->d_total, nw * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_eff, nw * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cells, nw * n_samples * sizeof(kcf_cell_t));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flags, CF_COUNT * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_flags, 0, CF_COUNT * 4, ctx->stream);
    if (e == cudaSuccess && total_kmers && n_windows) {
        e = cudaMemcpyAsync(c->d_total, total_kmers, n_windows * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_eff, eff_len, n_windows * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // the caller's arrays may go away
        c->totals_set = true;
    }
    if (e != cudaSuccess) {
        kcf_cohort_destroy(c);
        return kcf_fail(ctx, e == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, "kcf_cohort_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return KCF_OK;
}

extern "C" int kcf_cohort_add_plan(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, uint64_t window_offset, kcf_plan *plan)
{
    if (!ctx || !c || !plan || c->ctx != ctx || plan->ctx != ctx || sample >= c->n_samples) return KCF_ERR_ARG;
    if (!plan->ran) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_cohort_add_plan before kcf_plan_run");
    if (window_offset + plan->n_wins > c->n_windows) return kcf_fail(ctx, KCF_ERR_ARG, "plan rows [%llu, %llu) outside the cohort's %llu windows",
                                                                     (unsigned long long)window_offset, (unsigned long long)(window_offset + plan->n_wins),
                                                                     (unsigned long long)c->n_windows);
    if (plan->n_wins == 0) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    // the first sample to arrive defines TOTAL_KMERS / EFFLEN of its windows unless kcf_cohort_create was given them
    const bool any = std::any_of(c->filled.begin(), c->filled.end(), [](uint8_t f) { return f != 0; });
    const int set_totals = (!c->totals_set && (!any || c->filled[sample] == 1)) ? 1 : 0;
    kcf_cohort_from_plan_kernel<<<(unsigned)((plan->n_wins + 255) / 256), 256, 0, ctx->stream>>>(
        plan->d_out, plan->n_wins, c->d_cells + (uint64_t)sample * c->n_windows + window_offset, c->d_total + window_offset, c->d_eff + window_offset,
        set_totals, c->d_flags);
    KCF_CUDA(ctx, cudaGetLastError());
    c->filled[sample] = 1;
    return KCF_OK;
}

extern "C" int kcf_cohort_set_sample(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, const kcf_cell_t *cells)
{
    if (!ctx || !c || c->ctx != ctx || sample >= c->n_samples || (!cells && c->n_windows)) return KCF_ERR_ARG;
    if (!c->totals_set) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_cohort_set_sample needs the window totals given to kcf_cohort_create");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (c->n_windows) {
        KCF_CUDA(ctx, cudaMemcpyAsync(c->d_cells + (uint64_t)sample * c->n_windows, cells, c->n_windows * sizeof(kcf_cell_t), cudaMemcpyHostToDevice, ctx->stream));
        KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    c->filled[sample] = 2;
    return KCF_OK;
}

static int kcf_cohort_ready(kcf_ctx *ctx, kcf_cohort *c)
{
    for (uint32_t s = 0; s < c->n_samples; ++s)
        if (!c->filled[s]) return kcf_fail(ctx, KCF_ERR_ARG, "cohort sample %u has no data yet", s);
    return KCF_OK;
}

extern "C" int kcf_cohort_scores(kcf_ctx *ctx, kcf_cohort *c, const double w[3])
{
    if (!ctx || !c || !w || c->ctx != ctx) return KCF_ERR_ARG;
    int rc = kcf_cohort_ready(ctx, c);
    if (rc != KCF_OK) return rc;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = c->n_windows * c->n_samples;
    uint32_t flags[CF_COUNT] = {0};
    if (n) {
        kcf_cohort_score_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(c->d_cells, c->d_total, c->d_eff, c->n_windows, c->n_samples, w[0], w[1],
                                                                                     w[2], c->d_flags);
        KCF_CUDA(ctx, cudaGetLastError());
    }
    KCF_CUDA(ctx, cudaMemcpyAsync(flags, c->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[CF_MISMATCH]) return kcf_fail(ctx, KCF_ERR_ARG, "Windows mismatch found: the samples of the cohort were not screened over the same windows");
    // Data.java:101-103 — evaluated only for cells that reach the formula, left to right in double
    if (flags[CF_SCORE_USED] && w[0] + w[1] + w[2] != 1.0) return kcf_fail(ctx, KCF_ERR_WEIGHTS, "Weights should sum to 1.0");
    return KCF_OK;
}

extern "C" int kcf_cohort_find_ibs(kcf_ctx *ctx, kcf_cohort *c, const uint32_t *order, const uint32_t *chrom, uint64_t n, int detect_var,
                                   int32_t min_consecutive, float score_cutoff)
{
    if (!ctx || !c || c->ctx != ctx || (n && (!order || !chrom))) return KCF_ERR_ARG;
    int rc = kcf_cohort_ready(ctx, c);
    if (rc != KCF_OK) return rc;
    for (uint64_t i = 0; i < n; ++i)
        if (order[i] >= c->n_windows) return kcf_fail(ctx, KCF_ERR_ARG, "order[%llu] = %u outside the cohort's windows", (unsigned long long)i, order[i]);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t *d = nullptr;
    KCF_CUDA(ctx, cudaMalloc(&d, std::max<uint64_t>(n, 1) * 8));
    cudaError_t e = cudaMemcpyAsync(d, order, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, chrom, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n) {
        const unsigned threads = 128, warps = threads / 32;
        kcf_cohort_ibs_kernel<<<(c->n_samples + warps - 1) / warps, threads, 0, ctx->stream>>>(c->d_cells, c->n_windows, c->n_samples, d, d + n, n, detect_var,
                                                                                              min_consecutive, (double)score_cutoff);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "kcf_cohort_find_ibs: %s", cudaGetErrorString(e));
    return KCF_OK;
}

extern "C" int kcf_cohort_genotypes(kcf_ctx *ctx, kcf_cohort *c, double score_a, double score_b, double score_n, double min_maf, double max_missing,
                                    int8_t *alleles_out, uint8_t *bad_out)
{
    if (!ctx || !c || c->ctx != ctx || (c->n_windows && (!alleles_out || !bad_out))) return KCF_ERR_ARG;
    int rc = kcf_cohort_ready(ctx, c);
    if (rc != KCF_OK) return rc;
    if (c->n_windows == 0) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    int8_t *d_al = nullptr;
    uint8_t *d_bad = nullptr;
    const uint64_t na = c->n_windows * c->n_samples;
    KCF_CUDA(ctx, cudaMalloc(&d_al, na));
    cudaError_t e = cudaMalloc(&d_bad, c->n_windows);
    if (e == cudaSuccess) {
        kcf_cohort_gt_kernel<<<(unsigned)((c->n_windows + 127) / 128), 128, 0, ctx->stream>>>(c->d_cells, c->n_windows, c->n_samples, score_a, score_b, score_n,
                                                                                             min_maf, max_missing, d_al, d_bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(alleles_out, d_al, na, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bad_out, d_bad, c->n_windows, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_al);
    cudaFree(d_bad);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "kcf_cohort_genotypes: %s", cudaGetErrorString(e));
    return KCF_OK;
}

extern "C" int kcf_cohort_fetch(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, kcf_cell_t *cells_out, int32_t *total_kmers_out, int32_t *eff_len_out)
{
    if (!ctx || !c || c->ctx != ctx || sample >= c->n_samples) return KCF_ERR_ARG;
    if (!c->filled[sample]) return kcf_fail(ctx, KCF_ERR_ARG, "cohort sample %u has no data yet", sample);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t flags[CF_COUNT] = {0};
    if (c->n_windows) {
        if (cells_out) KCF_CUDA(ctx, cudaMemcpyAsync(cells_out, c->d_cells + (uint64_t)sample * c->n_windows, c->n_windows * sizeof(kcf_cell_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (total_kmers_out) KCF_CUDA(ctx, cudaMemcpyAsync(total_kmers_out, c->d_total, c->n_windows * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (eff_len_out) KCF_CUDA(ctx, cudaMemcpyAsync(eff_len_out, c->d_eff, c->n_windows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KCF_CUDA(ctx, cudaMemcpyAsync(flags, c->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[CF_MISMATCH]) return kcf_fail(ctx, KCF_ERR_ARG, "Windows mismatch found: the samples of the cohort were not screened over the same windows");
    return KCF_OK;
}
