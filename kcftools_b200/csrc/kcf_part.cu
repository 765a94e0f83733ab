// kcf_part.cu — screening against a database that is partitioned over several GPUs (placement 1, SURVEY §8e: a
// database beyond one GPU's HBM is cut by home line, 1/world per rank).
//
// A rank cannot answer its own k-mers any more, so the hot path splits at the probe:
//
//   requester   kcf_xchg_extract   screening front half (stage, canonical k-mer, minimizer -> global home line) for a
//                                  batch of tiles, then the k-mers grouped by the rank that owns their home line
//   host        all-to-all of the (key, home) arrays        <- the one real exchange step of this path (NCCL)
//   owner       kcf_xchg_lookup    probe of the local slice for every received k-mer (consecutive k-mers of a sender
//                                  stay consecutive, so neighbours still share their line requests)
//   host        all-to-all of the counts back
//   requester   kcf_xchg_fold      counts back to positions, hit bitmaps, gap summaries per tile
//   requester   kcf_plan_finalize  K5 as in the replicated path
//
// The host moves caller-owned device buffers between ranks (torch.distributed / NCCL in kcftools_b200/partitioned.py);
// this file neither knows nor links a communication library.
#include <algorithm>
#include <vector>
#include "kcf_internal.cuh"
#include "kcf_lookup.cuh"
#include "kcf_gap.cuh"

#define KCF_MAX_WORLD 64

// ---- requester: group the extracted k-mers by owner -------------------------------------------------------------
__global__ void __launch_bounds__(256) kcf_part_count_kernel(const uint32_t *__restrict__ homes, uint64_t n, uint64_t n_lines, uint32_t world,
                                                             unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int s_cnt[KCF_MAX_WORLD];
    if (threadIdx.x < KCF_MAX_WORLD) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = homes[i];
        if (h != 0xFFFFFFFFu) atomicAdd(&s_cnt[kcf_line_owner(h, n_lines, world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// A CTA moves KCF_SCATTER_CHUNK consecutive positions: it counts them per owner in shared memory, reserves one range per
// owner with ONE global atomic each, and then every warp step places its lanes (the lanes bound for the same owner take
// consecutive slots), so neighbouring k-mers stay neighbours on the owner's side and the global cursors see two
// atomics per 4096 positions instead of one per warp step.
#define KCF_SCATTER_CHUNK 4096
__global__ void __launch_bounds__(256) kcf_part_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ homes,
                                                               uint64_t n, uint64_t n_lines, uint32_t world, unsigned long long *__restrict__ cursor,
                                                               unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ homes_out,
                                                               uint32_t *__restrict__ src_out)
{
    __shared__ unsigned int s_cnt[KCF_MAX_WORLD], s_off[KCF_MAX_WORLD];
    __shared__ unsigned long long s_base[KCF_MAX_WORLD];
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t c0 = (uint64_t)blockIdx.x * KCF_SCATTER_CHUNK; c0 < n; c0 += (uint64_t)gridDim.x * KCF_SCATTER_CHUNK) {
        if (threadIdx.x < KCF_MAX_WORLD) s_cnt[threadIdx.x] = s_off[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t c1 = min(c0 + (uint64_t)KCF_SCATTER_CHUNK, n);
        for (uint64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const uint32_t h = homes[i];
            if (h != 0xFFFFFFFFu) atomicAdd(&s_cnt[kcf_line_owner(h, n_lines, world)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
        for (uint64_t i0 = c0 + (threadIdx.x & ~31u); i0 < c1; i0 += blockDim.x) {
            const uint64_t i = i0 + lane;
            const uint32_t h = i < c1 ? homes[i] : 0xFFFFFFFFu;
            const uint32_t owner = h != 0xFFFFFFFFu ? kcf_line_owner(h, n_lines, world) : 0xFFFFFFFFu;
            uint32_t todo = __ballot_sync(0xffffffffu, owner != 0xFFFFFFFFu);
            while (todo) {
                const uint32_t leader = __ffs(todo) - 1;
                const uint32_t o = __shfl_sync(0xffffffffu, owner, leader);
                const uint32_t same = __ballot_sync(0xffffffffu, owner == o);
                unsigned int off = 0;
                if (lane == leader) off = atomicAdd(&s_off[o], (unsigned int)__popc(same));
                off = __shfl_sync(0xffffffffu, off, leader);
                if (owner == o) {
                    const uint64_t dst = s_base[o] + off + __popc(same & ((1u << lane) - 1u));
                    keys_out[dst] = keys[i];
                    homes_out[dst] = h;
                    src_out[dst] = (uint32_t)i;
                }
                todo &= ~same;
            }
        }
        __syncthreads();
    }
}

// ---- owner: probe the local slice -----------------------------------------------------------------------------------
// Threads own consecutive received k-mers = consecutive k-mers of the sender, so the lanes of a warp mostly ask for the
// same few home lines and the coalescer merges them, as in the replicated kernel.
template <int S>
__global__ void __launch_bounds__(256) kcf_part_lookup_kernel(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                                              KcfTableGeom g, const unsigned long long *__restrict__ keys,
                                                              const uint32_t *__restrict__ homes, uint64_t n, uint32_t *__restrict__ counts)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    const uint32_t home = homes[i];
    const uint8_t *L = table + (uint64_t)kcf_line_wrap(home, 0, g) * KCF_LINE_BYTES;
    uint32_t c = 0;
    if (!(KCF_KEY_IN_LINES(key) && kcf_probe_line<S>(L, key, c))) {
        c = 0;
        if (kcf_filter_pass(L, key, g))
            c = kcf_probe_lines(table, stash, g, key, home, kcf_mask_from_word31(__ldg(reinterpret_cast<const uint32_t *>(L) + 31)), 1);
    }
    counts[i] = c;
}

// ---- requester: counts back to positions, then the gap summaries -----------------------------------------------------
__global__ void __launch_bounds__(256) kcf_part_unscatter_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ src, uint64_t n,
                                                                 uint32_t *__restrict__ cnt_by_pos)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt_by_pos[src[i]] = counts[i];
}

// one warp per tile of KCF_TILE positions: 64 words of 32 positions, two per lane, reduced in order
__global__ void __launch_bounds__(128) kcf_part_fold_kernel(const uint32_t *__restrict__ cnt_by_pos, const uint32_t *__restrict__ okw,
                                                            const uint32_t *__restrict__ start, uint64_t n_tiles, uint32_t k, int32_t min_count,
                                                            KcfGap *__restrict__ tile_sum)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    constexpr int WORDS = KCF_TILE / 32; // 64
    unsigned long long sum = 0;
    uint32_t hw0 = 0, hw1 = 0;
    for (int wd = 0; wd < WORDS; ++wd) {
        const uint64_t pos = t * KCF_TILE + 32ULL * wd + lane;
        const uint32_t vw = okw[t * WORDS + wd];
        const bool ok = (vw >> lane) & 1u;
        const uint32_t c = ok ? cnt_by_pos[pos] : 0u;
        const bool hit = ok && (int32_t)c >= min_count; // Java int compare (GetVariants.java:224)
        if (hit) sum += c;
        const uint32_t hb = __ballot_sync(0xffffffffu, hit);
        if ((uint32_t)(wd & 31) == lane) {
            if (wd < 32) hw0 = hb;
            else hw1 = hb;
        }
    }
    const KcfGap a = kcf_gap_fold_warp(hw0, okw[t * WORDS + lane], start[t * WORDS + lane], lane, k);
    const KcfGap b = kcf_gap_fold_warp(hw1, okw[t * WORDS + 32 + lane], start[t * WORDS + 32 + lane], lane, k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if (lane == 0) {
        KcfGap r = kcf_gap_combine(a, b, k);
        r.sum = sum;
        tile_sum[t] = r;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int kcf_xchg_reserve(kcf_ctx *ctx, kcf_plan *plan, uint64_t positions)
{
    if (plan->x_cap >= positions && plan->x_keys) return KCF_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(plan->x_keys);
    cudaFree(plan->x_homes);
    cudaFree(plan->x_okw);
    cudaFree(plan->x_start);
    cudaFree(plan->x_cnt);
    plan->x_keys = nullptr;
    plan->x_homes = plan->x_okw = plan->x_start = plan->x_cnt = nullptr;
    plan->x_cap = 0;
    KCF_CUDA(ctx, cudaMalloc(&plan->x_keys, positions * 8));
    KCF_CUDA(ctx, cudaMalloc(&plan->x_homes, positions * 4));
    KCF_CUDA(ctx, cudaMalloc(&plan->x_okw, positions / 32 * 4 + 4));
    KCF_CUDA(ctx, cudaMalloc(&plan->x_start, positions / 32 * 4 + 4));
    KCF_CUDA(ctx, cudaMalloc(&plan->x_cnt, positions * 4));
    if (!plan->x_cursor) KCF_CUDA(ctx, cudaMalloc(&plan->x_cursor, 2 * KCF_MAX_WORLD * sizeof(unsigned long long)));
    plan->x_cap = positions;
    return KCF_OK;
}

extern "C" int kcf_xchg_extract(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, int world,
                                void *d_keys_out, void *d_homes_out, void *d_src_out, uint64_t cap, uint64_t *send_counts)
{
    if (!ctx || !db || !plan || !send_counts || plan->ctx != ctx || db->ctx != ctx) return KCF_ERR_ARG;
    if (world < 1 || world > KCF_MAX_WORLD) return kcf_fail(ctx, KCF_ERR_ARG, "world %d outside 1..%d", world, KCF_MAX_WORLD);
    if (world != db->part_world) return kcf_fail(ctx, KCF_ERR_ARG, "database was opened as slice %d of %d, not of %d", db->part_rank, db->part_world, world);
    if (plan->k != db->info.kmer_length) return kcf_fail(ctx, KCF_ERR_ARG, "plan built for k=%d, database has k=%d", plan->k, db->info.kmer_length);
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    for (int r = 0; r < world; ++r) send_counts[r] = 0;
    if (tile_begin >= tile_end) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t npos = (tile_end - tile_begin) * KCF_TILE;
    if (npos >= (1ULL << 32)) return kcf_fail(ctx, KCF_ERR_ARG, "batch of %llu positions: keep batches below 2^32", (unsigned long long)npos);
    int rc = kcf_xchg_reserve(ctx, plan, npos);
    if (rc != KCF_OK) return rc;
    KCF_CUDA(ctx, cudaMemsetAsync(plan->x_homes, 0xFF, npos * 4, ctx->stream));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->x_okw, 0, npos / 32 * 4, ctx->stream));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->x_start, 0, npos / 32 * 4, ctx->stream));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->x_cursor, 0, 2 * KCF_MAX_WORLD * sizeof(unsigned long long), ctx->stream));
    rc = kcf_launch_screen(ctx, db, plan, 1, tile_begin, tile_end, nullptr, true, nullptr, nullptr);
    if (rc != KCF_OK) return rc;
    const unsigned grid = (unsigned)std::min<uint64_t>((npos + 255) / 256, (uint64_t)ctx->sm_count * 8);
    kcf_part_count_kernel<<<grid, 256, 0, ctx->stream>>>(plan->x_homes, npos, db->geom.n_lines, (uint32_t)world, plan->x_cursor);
    unsigned long long h_counts[KCF_MAX_WORLD];
    KCF_CUDA(ctx, cudaMemcpyAsync(h_counts, plan->x_cursor, world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long offs[KCF_MAX_WORLD], total = 0;
    for (int r = 0; r < world; ++r) {
        offs[r] = total;
        total += h_counts[r];
        send_counts[r] = h_counts[r];
    }
    if (total > cap) return kcf_fail(ctx, KCF_ERR_ARG, "send buffers hold %llu records, the batch has %llu", (unsigned long long)cap, total);
    if (total && (!d_keys_out || !d_homes_out || !d_src_out)) return KCF_ERR_ARG;
    // the scatter cursors start at each owner's offset (second half of x_cursor)
    KCF_CUDA(ctx, cudaMemcpyAsync(plan->x_cursor + KCF_MAX_WORLD, offs, world * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    if (total)
        kcf_part_scatter_kernel<<<grid, 256, 0, ctx->stream>>>(plan->x_keys, plan->x_homes, npos, db->geom.n_lines, (uint32_t)world,
                                                               plan->x_cursor + KCF_MAX_WORLD, (unsigned long long *)d_keys_out,
                                                               (uint32_t *)d_homes_out, (uint32_t *)d_src_out);
    KCF_CUDA(ctx, cudaGetLastError());
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // offs[] leaves scope; the caller hands the buffers to its communication library
    return KCF_OK;
}

extern "C" int kcf_xchg_lookup(kcf_ctx *ctx, kcf_db *db, const void *d_keys, const void *d_homes, uint64_t n, void *d_counts_out)
{
    if (!ctx || !db || db->ctx != ctx || (n && (!d_keys || !d_homes || !d_counts_out))) return KCF_ERR_ARG;
    if (n == 0) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((n + 255) / 256);
    const unsigned long long *kp = (const unsigned long long *)d_keys;
    const uint32_t *hp = (const uint32_t *)d_homes;
    if (db->geom.S == 13) kcf_part_lookup_kernel<13><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, kp, hp, n, (uint32_t *)d_counts_out);
    else if (db->geom.S == 12) kcf_part_lookup_kernel<12><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, kp, hp, n, (uint32_t *)d_counts_out);
    else kcf_part_lookup_kernel<10><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, kp, hp, n, (uint32_t *)d_counts_out);
    KCF_CUDA(ctx, cudaGetLastError());
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KCF_OK;
}

extern "C" int kcf_xchg_fold(kcf_ctx *ctx, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, const void *d_counts_back, const void *d_src,
                             uint64_t n, int32_t min_count)
{
    if (!ctx || !plan || plan->ctx != ctx || (n && (!d_counts_back || !d_src))) return KCF_ERR_ARG;
    if (min_count < 1) return kcf_fail(ctx, KCF_ERR_ARG, "Minimum kmer count should be at least 1");
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    if (tile_begin >= tile_end) return KCF_OK;
    const uint64_t npos = (tile_end - tile_begin) * KCF_TILE;
    if (!plan->x_cnt || plan->x_cap < npos) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_xchg_fold without the matching kcf_xchg_extract");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n) kcf_part_unscatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d_counts_back, (const uint32_t *)d_src, n, plan->x_cnt);
    const uint64_t nt = tile_end - tile_begin;
    kcf_part_fold_kernel<<<(unsigned)((nt * 32 + 127) / 128), 128, 0, ctx->stream>>>(plan->x_cnt, plan->x_okw, plan->x_start, nt, (uint32_t)plan->k,
                                                                                    min_count, plan->d_tile_sum + tile_begin);
    KCF_CUDA(ctx, cudaGetLastError());
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the caller may reuse its buffers
    return KCF_OK;
}

// ---- scan placement: every rank walks every tile, probes what it owns, bitmaps are reduced over ranks ------------------
// The second way to screen against a partitioned table (the first is the k-mer exchange above).  The 2-bit reference is
// small next to the table (0.375 B per base), so it is replicated; a rank runs the whole screening front half over ALL
// tiles but probes only the k-mers whose home line lies in its slice (ownership goes by minimizer, so the owned k-mers
// still come in runs that share a line).  What crosses NVLink is one hit BIT per position and one Σcount per tile —
// a sum-reduction (the owners' bitmaps are disjoint, so + is OR) — instead of 12-16 bytes per k-mer each way.

// one warp per tile: 64 words of 32 positions, two per lane, reduced in order
__global__ void __launch_bounds__(128) kcf_scan_fold_kernel(const uint32_t *__restrict__ hit, const uint32_t *__restrict__ okw,
                                                            const uint32_t *__restrict__ start, const unsigned long long *__restrict__ sums,
                                                            uint64_t n_tiles, uint32_t k, KcfGap *__restrict__ tile_sum)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    constexpr int WORDS = KCF_TILE / 32; // 64
    const uint64_t w0 = t * WORDS + lane, w1 = w0 + 32;
    // a hit bit is only ever set where a k-mer ends; the mask keeps a corrupted reduction from inventing k-mers
    const KcfGap a = kcf_gap_fold_warp(hit[w0] & okw[w0], okw[w0], start[w0], lane, k);
    const KcfGap b = kcf_gap_fold_warp(hit[w1] & okw[w1], okw[w1], start[w1], lane, k);
    if (lane == 0) {
        KcfGap r = kcf_gap_combine(a, b, k);
        r.sum = sums[t];
        tile_sum[t] = r;
    }
}

static int kcf_scan_reserve(kcf_ctx *ctx, kcf_plan *plan, uint64_t words)
{
    if (plan->s_cap >= words && plan->s_okw) return KCF_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(plan->s_okw);
    cudaFree(plan->s_start);
    plan->s_okw = plan->s_start = nullptr;
    plan->s_cap = 0;
    KCF_CUDA(ctx, cudaMalloc(&plan->s_okw, words * 4));
    KCF_CUDA(ctx, cudaMalloc(&plan->s_start, words * 4));
    plan->s_cap = words;
    return KCF_OK;
}

extern "C" int kcf_scan_owned(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, int32_t min_count,
                              void *d_hit_out, void *d_sum_out)
{
    if (!ctx || !db || !plan || plan->ctx != ctx || db->ctx != ctx) return KCF_ERR_ARG;
    if (min_count < 1) return kcf_fail(ctx, KCF_ERR_ARG, "Minimum kmer count should be at least 1"); // GetVariants.java:383-385
    if (plan->k != db->info.kmer_length) return kcf_fail(ctx, KCF_ERR_ARG, "plan built for k=%d, database has k=%d", plan->k, db->info.kmer_length);
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    if (tile_begin >= tile_end) return KCF_OK;
    if (!d_hit_out || !d_sum_out) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nt = tile_end - tile_begin, words = nt * (KCF_TILE / 32);
    int rc = kcf_scan_reserve(ctx, plan, words);
    if (rc != KCF_OK) return rc;
    // chunks past the end of a window are never visited: their words must read "no k-mer, no hit"
    KCF_CUDA(ctx, cudaMemsetAsync(d_hit_out, 0, words * 4, ctx->stream));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->s_okw, 0, words * 4, ctx->stream));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->s_start, 0, words * 4, ctx->stream));
    rc = kcf_launch_screen(ctx, db, plan, min_count, tile_begin, tile_end, nullptr, false, (uint32_t *)d_hit_out, (unsigned long long *)d_sum_out);
    if (rc != KCF_OK) return rc;
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the caller hands the buffers to its communication library
    return KCF_OK;
}

extern "C" int kcf_scan_fold(kcf_ctx *ctx, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, const void *d_hit, const void *d_sum)
{
    if (!ctx || !plan || plan->ctx != ctx) return KCF_ERR_ARG;
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    if (tile_begin >= tile_end) return KCF_OK;
    if (!d_hit || !d_sum) return KCF_ERR_ARG;
    const uint64_t nt = tile_end - tile_begin;
    if (!plan->s_okw || plan->s_cap < nt * (KCF_TILE / 32)) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_scan_fold without the matching kcf_scan_owned");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    kcf_scan_fold_kernel<<<(unsigned)((nt * 32 + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t *)d_hit, plan->s_okw, plan->s_start,
                                                                                    (const unsigned long long *)d_sum, nt, (uint32_t)plan->k,
                                                                                    plan->d_tile_sum + tile_begin);
    KCF_CUDA(ctx, cudaGetLastError());
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the caller may reuse its buffers
    return KCF_OK;
}
