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

// =========================================================================================================================
// k-mer exchange over peer memory (KcfXgDev, kcf_internal.cuh): send -> [barrier] -> answer -> [barrier] -> fold, all
// stream-ordered; the barriers are the caller's (a tiny NCCL collective on this context's stream, kcftools_b200/
// partitioned.py), nothing else crosses the host.  What moves: 16 bytes out and 16 bytes back (1-byte counters; 48
// otherwise) per RUN of up to 11 k-mers sharing a home line — about 5 bytes per k-mer — (world - 1) / world of it over
// NVLink, written by the kernels themselves into the peers' memory.
// =========================================================================================================================
struct kcf_xg {
    kcf_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    uint64_t batch_positions = 0, cap = 0;
    uint32_t cbytes = 4;
    uint64_t n_lines = 0;          // of the whole table (all slices)
    uint8_t *block = nullptr;      // exported: inbox runs | runs per sender | back
    uint64_t block_bytes = 0, off_keys = 0, off_homes = 0, off_count = 0, off_back = 0; // off_keys: the run entries
    uint8_t *local = nullptr;      // pos_slot | okw | start | cursor | flags
    uint8_t *peer[KCF_XG_MAX_WORLD] = {nullptr};
    bool opened[KCF_XG_MAX_WORLD] = {false};
    bool connected = false;
    KcfXgDev dev{};
    // pipelined use (kcf_xg_pipeline): sends run on their own stream with a capped grid, beside the answers of the batch before
    cudaStream_t send_stream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_sent = nullptr;
    uint32_t send_ctas = 0;
};

// tell every owner how many runs this rank appended to its inbox region
__global__ void kcf_xg_publish_kernel(KcfXgDev X)
{
    const uint32_t o = threadIdx.x;
    if (o < X.world) {
        const unsigned int n = X.cursor[o];
        *X.in_count[o] = n < X.cap ? n : (uint32_t)X.cap;
    }
    __threadfence_system();
}

// owner: a warp takes 32 runs of a sender's region and spreads their k-mers over its lanes — lane l of a step looks up the
// l-th k-mer of the block, found through the prefix sums of the run lengths — so that, as in the replicated kernel, the
// lanes that share a run share its home line inside ONE load instruction and the coalescer fetches the line once.  (One
// thread per run asked for the line's sectors one after the other: 84 ms per c4s step; a quad per run with a whole-line
// prefetch: 50 ms — the later loads did not hit L1.)  The rest follows the replicated kernel too: filter and mask word
// travel with the key words; a k-mer whose home line names other lines goes to a per-warp queue that is searched one item
// per lane with whole-line probes (searching them where they turn up, slot by slot, ran 40 % of the kernel's instructions
// at two active lanes: profiles/r2w_answer_*).  The counts of a run (one byte each while the database's counts fit a byte)
// make the run's slot in that sender's BACK region; slots leave the SM whole, through shared memory: 512 contiguous bytes
// of the requester's memory per store instruction.
#define KCF_XG_QCAP 64
#ifndef KCF_XG_ANSWER_BLOCKS
#define KCF_XG_ANSWER_BLOCKS 6 // resident CTAs per SM the registers are held to (48 warps)
#endif
struct KcfXgItem {
    unsigned long long key;
    uint32_t home;
    uint32_t info; // byte offset of the count in the warp's staged slots << 16 | home mask (bit 0 cleared)
};

template <int S>
__global__ void __launch_bounds__(256, KCF_XG_ANSWER_BLOCKS) kcf_xg_answer_kernel(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash, const __grid_constant__ KcfTableGeom g, const __grid_constant__ KcfXgDev X)
{
    constexpr uint32_t CB = S == 13 ? 1u : 4u, STRIDE = CB == 1 ? 16u : 48u; // bytes per count / per run slot (= X.cbytes, X.stride)
    __shared__ __align__(16) uint8_t stage_all[8][32 * STRIDE];
    __shared__ __align__(16) KcfXgItem queue_all[8][KCF_XG_QCAP];
    uint8_t *stage = stage_all[threadIdx.x >> 5];
    KcfXgItem *queue = queue_all[threadIdx.x >> 5];
    const uint32_t s = blockIdx.y; // sender
    const uint32_t n = X.my_count[s];
    const uint4 *runs = X.my_runs + (uint64_t)s * X.cap;
    uint8_t *back = X.back[s];
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    auto put = [&](uint32_t at, uint32_t c) {
        if (CB == 1) stage[at] = (uint8_t)c;
        else *reinterpret_cast<uint32_t *>(stage + at) = c;
    };
    for (uint64_t i0 = warp * 32; i0 < n; i0 += n_warps * 32) {
        uint64_t p0 = 0, p1 = 0;
        uint32_t len = 0, home = 0;
        if (i0 + lane < n) {
            const uint4 e = __ldg(runs + i0 + lane);
            if ((e.x & e.y) != 0xFFFFFFFFu) // else: the unused rest of a sender warp's slab
                kcf_xg_unpack_run(((uint64_t)e.y << 32) | e.x, ((uint64_t)e.w << 32) | e.z, p0, p1, len, home);
        }
        uint32_t incl = len; // prefix sums of the run lengths over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        const uint32_t excl = incl - len, total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t qn = 0; // queue length (warp uniform)
        auto flush_queue = [&]() {
            __syncwarp();
#pragma unroll 1
            for (uint32_t t = lane; t < qn; t += 32) {
                const KcfXgItem it = queue[t];
                uint32_t m2 = it.info & 0x7FFEu, c2 = 0;
                bool found = false;
                while (m2 && !found) {
                    const uint32_t d = __ffs(m2) - 1;
                    m2 &= m2 - 1;
                    found = kcf_probe_line<S>(table + (uint64_t)kcf_line_wrap(it.home, d, g) * KCF_LINE_BYTES, it.key, c2);
                }
                if (!found) c2 = (it.info & (1u << KCF_STASH_BIT)) ? kcf_stash_find(stash, g, it.key) : 0u;
                put(it.info >> 16, c2);
            }
            qn = 0;
            __syncwarp();
        };
#pragma unroll 1
        for (uint32_t t0 = 0; t0 < total; t0 += 32) {
            const uint32_t kidx = t0 + lane;
            const bool active = kidx < total;
            uint32_t r = 0; // the run holding k-mer kidx: the last lane whose exclusive prefix is <= kidx
#pragma unroll
            for (uint32_t step = 16; step > 0; step >>= 1) {
                const uint32_t ex = __shfl_sync(0xffffffffu, excl, r + step);
                if (ex <= kidx) r += step;
            }
            const uint32_t j = kidx - __shfl_sync(0xffffffffu, excl, r);
            const uint64_t q0 = __shfl_sync(0xffffffffu, p0, r), q1 = __shfl_sync(0xffffffffu, p1, r);
            const uint32_t hm = __shfl_sync(0xffffffffu, home, r);
            const uint32_t at = r * STRIDE + j * CB;
            uint64_t key = 0;
            uint32_t mask = 0;
            bool pending = false;
            if (active) {
                uint32_t f0 = (uint32_t)(q0 >> j) & g.km, f1 = (uint32_t)(q1 >> j) & g.km;
                if (g.both_strands) kcf_plane_canonical(f0, f1, kcf_plane_rc(f0, g.k, g.km), kcf_plane_rc(f1, g.k, g.km), f0, f1);
                key = ((uint64_t)f1 << 32) | f0;
                const uint8_t *L = table + (uint64_t)kcf_line_wrap(hm, 0, g) * KCF_LINE_BYTES;
                // filter and mask words travel with the key words (a miss needs them; asked for after the compare they would
                // cost a dependent round trip)
                const uint32_t w31 = __ldg(reinterpret_cast<const uint32_t *>(L) + 31);
                const unsigned long long fword = S == 13 ? __ldg(reinterpret_cast<const unsigned long long *>(L + 104))
                                                         : (unsigned long long)__ldg(reinterpret_cast<const uint32_t *>(L + 120));
                const bool inl = KCF_KEY_IN_LINES(key);
                uint32_t c = 0;
                if (!(inl && kcf_probe_line<S>(L, key, c))) {
                    c = 0;
                    if (S == 13 ? kcf_filter_pass64(fword, key) : kcf_filter_pass32((uint32_t)fword, key)) {
                        mask = kcf_mask_from_word31(w31);
                        if (inl && (mask & 0x7FFEu)) pending = true;
                        else if ((mask >> KCF_STASH_BIT) & 1u) c = kcf_stash_find(stash, g, key);
                    }
                }
                put(at, c);
            }
            const uint32_t pb = __ballot_sync(0xffffffffu, pending);
            if (pb) {
                if (pending) {
                    KcfXgItem it;
                    it.key = key;
                    it.home = hm;
                    it.info = (at << 16) | (mask & 0xFFFEu);
                    queue[qn + __popc(pb & ((1u << lane) - 1u))] = it;
                }
                qn += __popc(pb);
                if (qn + 32 > KCF_XG_QCAP) flush_queue();
            }
        }
        if (qn) flush_queue();
        __syncwarp();
        const uint32_t n16 = (uint32_t)min((uint64_t)32, n - i0) * (STRIDE / 16u); // bytes past a run's length are never read
        uint4 *dst = reinterpret_cast<uint4 *>(back + i0 * STRIDE);
        for (uint32_t q = lane; q < n16; q += 32) dst[q] = reinterpret_cast<const uint4 *>(stage)[q];
        __syncwarp();
    }
}

// requester: one warp per tile of the batch; a position finds its count through the run slot its head noted at send time
__global__ void __launch_bounds__(128) kcf_xg_fold_kernel(KcfXgDev X, uint64_t n_tiles, uint32_t k, int32_t min_count, KcfGap *__restrict__ tile_sum)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    constexpr int WORDS = KCF_TILE / 32; // 64
    unsigned long long sum = 0;
    uint32_t hw0 = 0, hw1 = 0;
    for (int wd = 0; wd < WORDS; ++wd) {
        const uint32_t slot = X.pos_slot[t * KCF_TILE + 32ULL * wd + lane];
        const uint32_t headmask = __ballot_sync(0xffffffffu, slot < 0xFFFFFFFEu);
        uint32_t c = 0;
        const uint32_t below = headmask & (0xFFFFFFFFu >> (31u - lane)); // run heads at or before this lane (runs never cross a word)
        const uint32_t hl = below ? 31u - __clz(below) : lane;
        const uint32_t hs = __shfl_sync(0xffffffffu, slot, hl);
        const bool has = slot != 0xFFFFFFFFu && below != 0u && hs != 0xFFFFFFFFu; // (a head that found its region full has no slot: its run reads as absent)
        if (has) {
            const uint64_t at = ((uint64_t)(hs >> 28) * X.cap + (hs & 0x0FFFFFFFu)) * X.stride + (uint64_t)(lane - hl) * X.cbytes; // lane - hl: place in the run
            c = X.cbytes == 1 ? (uint32_t)X.my_back[at] : *reinterpret_cast<const uint32_t *>(X.my_back + at);
        }
        const bool hit = has && (int32_t)c >= min_count; // Java int compare (GetVariants.java:224)
        if (hit) sum += c;
        const uint32_t hb = __ballot_sync(0xffffffffu, hit);
        if ((uint32_t)(wd & 31) == lane) {
            if (wd < 32) hw0 = hb;
            else hw1 = hb;
        }
    }
    const KcfGap a = kcf_gap_fold_warp(hw0, X.okw[t * WORDS + lane], X.start[t * WORDS + lane], lane, k);
    const KcfGap b = kcf_gap_fold_warp(hw1, X.okw[t * WORDS + 32 + lane], X.start[t * WORDS + 32 + lane], lane, k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if (lane == 0) {
        KcfGap r = kcf_gap_combine(a, b, k);
        r.sum = sum;
        tile_sum[t] = r;
    }
}

static void kcf_xg_build_dev(kcf_xg *x)
{
    KcfXgDev &d = x->dev;
    d.world = (uint32_t)x->world;
    d.me = (uint32_t)x->rank;
    d.cbytes = x->cbytes;
    d.stride = x->cbytes == 1 ? 16u : 48u;
    d.cap = x->cap;
    d.own_mul = (uint32_t)std::min<uint64_t>(((uint64_t)x->world << 32) / x->n_lines, 0xFFFFFFFFULL);
    for (int r = 0; r <= KCF_XG_MAX_WORLD; ++r) // first home line with kcf_line_owner >= r
        d.own_bound[r] = r >= x->world ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)r * x->n_lines + (uint64_t)x->world - 1) / (uint64_t)x->world);
    for (int r = 0; r < x->world; ++r) {
        uint8_t *b = x->peer[r];
        d.in_runs[r] = reinterpret_cast<uint4 *>(b + x->off_keys) + (uint64_t)x->rank * x->cap;
        d.in_count[r] = reinterpret_cast<uint32_t *>(b + x->off_count) + x->rank;
        d.back[r] = b + x->off_back + (uint64_t)x->rank * x->cap * d.stride;
    }
    d.my_runs = reinterpret_cast<const uint4 *>(x->block + x->off_keys);
    d.my_count = reinterpret_cast<const uint32_t *>(x->block + x->off_count);
    d.my_back = x->block + x->off_back;
    const uint64_t words = x->batch_positions / 32;
    d.pos_slot = reinterpret_cast<uint32_t *>(x->local);
    d.okw = d.pos_slot + x->batch_positions;
    d.start = d.okw + words;
    d.cursor = reinterpret_cast<unsigned int *>(d.start + words);
    d.flags = reinterpret_cast<uint32_t *>(d.cursor + KCF_XG_MAX_WORLD);
}

extern "C" int kcf_xg_create(kcf_ctx *ctx, kcf_db *db, int rank, int world, uint64_t batch_tiles, kcf_xg **out)
{
    if (!ctx || !db || !out || db->ctx != ctx) return KCF_ERR_ARG;
    *out = nullptr;
    if (world < 1 || world > KCF_XG_MAX_WORLD || rank < 0 || rank >= world) return kcf_fail(ctx, KCF_ERR_ARG, "exchange: rank %d of %d (at most %d ranks)", rank, world, KCF_XG_MAX_WORLD);
    if (db->part_world != world || db->part_rank != rank) return kcf_fail(ctx, KCF_ERR_ARG, "exchange: the database is slice %d of %d, not %d of %d", db->part_rank, db->part_world, rank, world);
    if (db->geom.kw != 1) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "exchange: 64-bit keys only (k <= 32)");
    if (batch_tiles == 0 || batch_tiles * KCF_TILE >= (1ULL << 31)) return kcf_fail(ctx, KCF_ERR_ARG, "exchange: batch of %llu tiles", (unsigned long long)batch_tiles);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    kcf_xg *x = new kcf_xg();
    x->ctx = ctx;
    x->rank = rank;
    x->world = world;
    x->batch_positions = batch_tiles * KCF_TILE;
    x->n_lines = db->geom.n_lines;
    // A region holds RUNS (kcf_internal.cuh).  A sender's runs spread over the owners by a hash of their minimizer, 1/world
    // each; a window of random sequence makes one run per ~5 positions (6 k-mers on average, cut at every 32nd position),
    // the worst case — every k-mer alone — one per position.  Half the positions per owner share plus a fixed slack covers
    // any real batch; an overflow is detected and reported (kcf_xg_status), never silent.
    x->cap = std::min<uint64_t>(x->batch_positions, x->batch_positions / (2 * (uint64_t)world) + 65536);
    if (x->cap >= (1ULL << 28)) { delete x; return kcf_fail(ctx, KCF_ERR_ARG, "exchange: batch too large for %d ranks (region of %llu runs)", world, (unsigned long long)x->cap); }
    x->cbytes = db->geom.cw == 1 ? 1u : 4u;
    auto up = [](uint64_t b) { return (b + 255) & ~255ULL; };
    x->off_keys = 0;
    x->off_homes = 0;
    x->off_count = up(x->off_keys + (uint64_t)world * x->cap * 16);
    x->off_back = up(x->off_count + (uint64_t)world * 4);
    x->block_bytes = up(x->off_back + (uint64_t)world * x->cap * (x->cbytes == 1 ? 16 : 48));
    const uint64_t local_bytes = x->batch_positions * 4 + 2 * (x->batch_positions / 32) * 4 + KCF_XG_MAX_WORLD * 4 + 64;
    cudaError_t e = cudaMalloc(&x->block, x->block_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&x->local, local_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(x->block, 0, x->block_bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(x->local, 0, local_bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(x->block);
        cudaFree(x->local);
        delete x;
        return kcf_fail(ctx, e == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, "exchange workspace: %s", cudaGetErrorString(e));
    }
    x->peer[rank] = x->block;
    *out = x;
    return KCF_OK;
}

extern "C" int kcf_xg_export(kcf_xg *x, void *handle64_out, void **device_ptr_out, uint64_t *bytes_out)
{
    if (!x) return KCF_ERR_ARG;
    if (handle64_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, x->block);
        if (e != cudaSuccess) return kcf_fail(x->ctx, KCF_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        memcpy(handle64_out, &h, 64);
    }
    if (device_ptr_out) *device_ptr_out = x->block;
    if (bytes_out) *bytes_out = x->block_bytes;
    return KCF_OK;
}

extern "C" int kcf_xg_connect(kcf_xg *x, const void *handles, void *const *same_process_ptrs)
{
    if (!x || (!handles && !same_process_ptrs && x->world > 1)) return KCF_ERR_ARG;
    KCF_CUDA(x->ctx, cudaSetDevice(x->ctx->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        if (same_process_ptrs) {
            x->peer[r] = (uint8_t *)same_process_ptrs[r];
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const uint8_t *)handles + 64 * r, 64);
            void *pp = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&pp, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return kcf_fail(x->ctx, KCF_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            x->peer[r] = (uint8_t *)pp;
            x->opened[r] = true;
        }
        if (!x->peer[r]) return kcf_fail(x->ctx, KCF_ERR_ARG, "exchange: no workspace pointer for rank %d", r);
    }
    kcf_xg_build_dev(x);
    x->connected = true;
    return KCF_OK;
}

extern "C" void kcf_xg_destroy(kcf_xg *x)
{
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->stream);
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->send_stream) {
        cudaStreamSynchronize(x->send_stream);
        cudaStreamDestroy(x->send_stream);
        cudaEventDestroy(x->ev_main);
        cudaEventDestroy(x->ev_sent);
    }
    cudaFree(x->block);
    cudaFree(x->local);
    delete x;
}

extern "C" int kcf_xg_send(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, kcf_xg *x, uint64_t tile_begin, uint64_t tile_end)
{
    if (!ctx || !db || !plan || !x || x->ctx != ctx || plan->ctx != ctx || db->ctx != ctx) return KCF_ERR_ARG;
    if (!x->connected) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_xg_send before kcf_xg_connect");
    if (plan->k != db->info.kmer_length) return kcf_fail(ctx, KCF_ERR_ARG, "plan built for k=%d, database has k=%d", plan->k, db->info.kmer_length);
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nt = tile_begin < tile_end ? tile_end - tile_begin : 0;
    if (nt * KCF_TILE > x->batch_positions) return kcf_fail(ctx, KCF_ERR_ARG, "exchange: batch of %llu tiles exceeds the workspace", (unsigned long long)nt);
    cudaStream_t st = ctx->stream;
    int rcs = kcf_sync_seqs(ctx); // (queues on the context's stream when the sequence table changed)
    if (rcs != KCF_OK) return rcs;
    if (x->send_stream) { // after everything queued on the context's stream so far (the fold that last read this workspace among it)
        st = x->send_stream;
        KCF_CUDA(ctx, cudaEventRecord(x->ev_main, ctx->stream));
        KCF_CUDA(ctx, cudaStreamWaitEvent(st, x->ev_main, 0));
    }
    KCF_CUDA(ctx, cudaMemsetAsync(x->dev.cursor, 0, KCF_XG_MAX_WORLD * sizeof(unsigned int), st));
    if (nt) {
        // chunks past the end of a window are never visited: their positions must read "no k-mer"
        KCF_CUDA(ctx, cudaMemsetAsync(x->dev.pos_slot, 0xFF, nt * KCF_TILE * 4, st));
        KCF_CUDA(ctx, cudaMemsetAsync(x->dev.okw, 0, nt * (KCF_TILE / 32) * 4, st));
        KCF_CUDA(ctx, cudaMemsetAsync(x->dev.start, 0, nt * (KCF_TILE / 32) * 4, st));
        int rc = kcf_launch_screen(ctx, db, plan, 1, tile_begin, tile_end, nullptr, true, nullptr, nullptr, &x->dev, x->send_stream, x->send_ctas);
        if (rc != KCF_OK) return rc;
    }
    kcf_xg_publish_kernel<<<1, 32, 0, st>>>(x->dev); // every rank publishes every batch, empty or not
    KCF_CUDA(ctx, cudaGetLastError());
    if (x->send_stream) KCF_CUDA(ctx, cudaEventRecord(x->ev_sent, st));
    return KCF_OK;
}

// Pipelined use: from now on kcf_xg_send of this workspace runs on a stream of its own, on at most send_ctas_per_sm resident
// CTAs per SM (0 = no cap), ordered after whatever the context's stream holds at the time of the call; kcf_xg_join makes the
// context's stream wait for the last send.  With two workspaces a rank sends batch b + 1 while it answers batch b.
extern "C" int kcf_xg_pipeline(kcf_xg *x, uint32_t send_ctas_per_sm)
{
    if (!x) return KCF_ERR_ARG;
    kcf_ctx *ctx = x->ctx;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!x->send_stream) {
        KCF_CUDA(ctx, cudaStreamCreateWithFlags(&x->send_stream, cudaStreamNonBlocking));
        KCF_CUDA(ctx, cudaEventCreateWithFlags(&x->ev_main, cudaEventDisableTiming));
        KCF_CUDA(ctx, cudaEventCreateWithFlags(&x->ev_sent, cudaEventDisableTiming));
        KCF_CUDA(ctx, cudaEventRecord(x->ev_sent, x->send_stream));
    }
    x->send_ctas = send_ctas_per_sm;
    return KCF_OK;
}

extern "C" int kcf_xg_join(kcf_ctx *ctx, kcf_xg *x)
{
    if (!ctx || !x || x->ctx != ctx) return KCF_ERR_ARG;
    if (!x->send_stream) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    KCF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, x->ev_sent, 0));
    return KCF_OK;
}

extern "C" int kcf_xg_answer(kcf_ctx *ctx, kcf_db *db, kcf_xg *x)
{
    if (!ctx || !db || !x || x->ctx != ctx || db->ctx != ctx || !x->connected) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const dim3 grid((unsigned)std::min<uint64_t>((x->cap + 255) / 256, (uint64_t)ctx->sm_count * 8), (unsigned)x->world); // a warp per 32 runs, grid-stride
    if (db->geom.S == 13) kcf_xg_answer_kernel<13><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, x->dev);
    else if (db->geom.S == 12) kcf_xg_answer_kernel<12><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, x->dev);
    else kcf_xg_answer_kernel<10><<<grid, 256, 0, ctx->stream>>>(db->table, db->stash, db->geom, x->dev);
    KCF_CUDA(ctx, cudaGetLastError());
    return KCF_OK;
}

extern "C" int kcf_xg_fold(kcf_ctx *ctx, kcf_plan *plan, kcf_xg *x, uint64_t tile_begin, uint64_t tile_end, int32_t min_count)
{
    if (!ctx || !plan || !x || x->ctx != ctx || plan->ctx != ctx || !x->connected) return KCF_ERR_ARG;
    if (min_count < 1) return kcf_fail(ctx, KCF_ERR_ARG, "Minimum kmer count should be at least 1");
    tile_end = std::min<uint64_t>(tile_end, plan->n_tiles);
    if (tile_begin >= tile_end) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nt = tile_end - tile_begin;
    kcf_xg_fold_kernel<<<(unsigned)((nt * 32 + 127) / 128), 128, 0, ctx->stream>>>(x->dev, nt, (uint32_t)plan->k, min_count, plan->d_tile_sum + tile_begin);
    KCF_CUDA(ctx, cudaGetLastError());
    return KCF_OK;
}

// synchronises the stream; KCF_ERR_NOMEM when a region overflowed in any batch since the last call
extern "C" int kcf_xg_status(kcf_xg *x, uint64_t *bytes_out_per_run, uint64_t *bytes_back_per_run, uint64_t *runs_sent_last_batch)
{
    if (!x) return KCF_ERR_ARG;
    kcf_ctx *ctx = x->ctx;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (x->send_stream) KCF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, x->ev_sent, 0));
    uint32_t f = 0;
    unsigned int cur[KCF_XG_MAX_WORLD] = {0};
    KCF_CUDA(ctx, cudaMemcpyAsync(&f, x->dev.flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaMemcpyAsync(cur, x->dev.cursor, sizeof cur, cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (runs_sent_last_batch) {
        *runs_sent_last_batch = 0;
        for (int r = 0; r < x->world; ++r) *runs_sent_last_batch += cur[r];
    }
    if (bytes_out_per_run) *bytes_out_per_run = 16;
    if (bytes_back_per_run) *bytes_back_per_run = x->cbytes == 1 ? 16 : 48;
    if (f) {
        cudaMemsetAsync(x->dev.flags, 0, 4, ctx->stream);
        return kcf_fail(ctx, KCF_ERR_NOMEM, "exchange: an inbox region of %llu entries overflowed; use smaller batches", (unsigned long long)x->cap);
    }
    return KCF_OK;
}
