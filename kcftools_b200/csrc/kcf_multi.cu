// kcf_multi.cu — ONE getVariations job over the GPUs of one process: the replacement for the thread-pool fan-out of
// GetVariants.java:129-151 (one pool over all windows) when several devices are present.
//
// Windows are the reference's unit of parallelism and share no results, so the job shards without a
// data-path collective: the window list is cut into n contiguous ranges balanced on Σ window length, every context
// (one per GPU, database replicated) gets one range, and a host thread per context uploads ONLY the stretches of the
// reference its windows touch — line-aligned pieces of the FASTA bytes, each registered as its own sequence — plans
// the windows whose last base lies in a piece right behind that piece's upload (the screening of piece i runs while
// piece i+1 crosses PCIe), and fetches the rows into the caller's array at their original positions.  A JVM host
// needs no communication library for this: one call, n contexts.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "kcf_internal.cuh"

// upper bound of one uploaded piece, in bases (rounded down to whole FASTA lines).  Measured on C2 (75 Mb chromosomes): 19.9 ms
// per job with 24 Mbase pieces, 18.5 with 48, 18.2 with 96 (the bare copy takes 16.5): a piece costs ~0.1 ms of host work
// (plan, launches), the tail after the last upload only the screening of that piece
#ifndef KCF_PIECE_BASES
#define KCF_PIECE_BASES (96u << 20)
#endif

extern "C" int kcf_set_upload_piece(kcf_ctx *ctx, uint64_t bases)
{
    if (!ctx) return KCF_ERR_ARG;
    if (bases != 0 && (bases < 1024 || bases > (1ULL << 30))) return kcf_fail(ctx, KCF_ERR_ARG, "upload piece of %llu bases (0 = default, else 1024 .. 2^30)", (unsigned long long)bases);
    ctx->piece_bases = bases;
    return KCF_OK;
}

extern "C" int kcf_shard_windows(const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs, uint64_t n_segs, int n_shards,
                                 uint64_t *bounds_out)
{
    if ((!wins && n_wins) || (!segs && n_segs) || n_shards < 1 || !bounds_out) return KCF_ERR_ARG;
    std::vector<uint64_t> csum(n_wins + 1, 0);
    for (uint64_t w = 0; w < n_wins; ++w) {
        if ((uint64_t)wins[w].first_seg + wins[w].n_segs > n_segs) return KCF_ERR_ARG;
        uint64_t len = 0;
        for (uint32_t s = 0; s < wins[w].n_segs; ++s) len += (uint64_t)std::max(segs[wins[w].first_seg + s].len, 0);
        csum[w + 1] = csum[w] + len;
    }
    const uint64_t total = csum[n_wins];
    bounds_out[0] = 0;
    for (int r = 1; r < n_shards; ++r) {
        // first window whose prefix sum reaches total * r / n (the same cut kcftools_b200/shard.py makes)
        const uint64_t target = (uint64_t)((unsigned __int128)total * (unsigned)r / (unsigned)n_shards);
        uint64_t c = (uint64_t)(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin());
        c = std::min(std::max(c, bounds_out[r - 1]), n_wins);
        bounds_out[r] = c;
    }
    bounds_out[n_shards] = n_wins;
    return KCF_OK;
}

namespace {
struct Piece {
    uint32_t seq;          // caller's sequence
    uint64_t base0, base1; // bases [base0, base1) of it; base0 is a multiple of line_bases
    int local_id = -1;     // id kcf_ref_add_async returned on this context
};

struct ShardJob {
    kcf_ctx *ctx;
    kcf_db *db;
    uint64_t w0, w1;
    int rc = KCF_OK;
};
} // namespace

// screen windows [w0, w1) of the job on one context; rows go to out[w0 .. w1)
static int kcf_screen_shard(kcf_ctx *ctx, kcf_db *db, const kcf_host_seq_t *seqs, uint32_t n_seqs, const kcf_window_t *wins, uint64_t w0,
                            uint64_t w1, const kcf_segment_t *segs, uint64_t n_segs, int32_t min_count, const double w[3], kcf_result_t *out)
{
    int rc = kcf_ref_clear(ctx);
    if (rc != KCF_OK || w0 >= w1) return rc;
    const uint64_t piece_bases = ctx->piece_bases ? ctx->piece_bases : (uint64_t)KCF_PIECE_BASES;
    // ---- the stretch of every sequence this shard touches
    std::vector<uint64_t> lo(n_seqs, ~0ULL), hi(n_seqs, 0);
    for (uint64_t wi = w0; wi < w1; ++wi) {
        const kcf_window_t &win = wins[wi];
        if (win.n_segs == 0) return kcf_fail(ctx, KCF_ERR_RANGE, "Fasta object is null for window %llu (no segment)", (unsigned long long)wi);
        if ((uint64_t)win.first_seg + win.n_segs > n_segs) return kcf_fail(ctx, KCF_ERR_ARG, "window %llu: segment range outside segs[]", (unsigned long long)wi);
        for (uint32_t s = 0; s < win.n_segs; ++s) {
            const kcf_segment_t &sg = segs[win.first_seg + s];
            if (sg.seq_id < 0 || (uint32_t)sg.seq_id >= n_seqs)
                return kcf_fail(ctx, KCF_ERR_RANGE, "Sequence not found in index: id %d (window %llu)", sg.seq_id, (unsigned long long)wi);
            const int64_t a = sg.start0, b = (int64_t)sg.start0 + sg.len;
            if (a < 0 || b > (int64_t)seqs[sg.seq_id].seq_len || a >= b) // FastaIndex.java:132-135
                return kcf_fail(ctx, KCF_ERR_RANGE, "Invalid range: %lld-%lld for sequence: id %d", (long long)a, (long long)b, sg.seq_id);
            lo[sg.seq_id] = std::min<uint64_t>(lo[sg.seq_id], (uint64_t)a);
            hi[sg.seq_id] = std::max<uint64_t>(hi[sg.seq_id], (uint64_t)b);
        }
    }
    // ---- pieces, in sequence order
    std::vector<Piece> pieces;
    std::vector<uint32_t> first_piece(n_seqs + 1, 0);
    for (uint32_t s = 0; s < n_seqs; ++s) {
        first_piece[s] = (uint32_t)pieces.size();
        if (hi[s] == 0) continue;
        const uint64_t lb = seqs[s].line_bases;
        if (lb == 0 || seqs[s].line_width < lb) return kcf_fail(ctx, KCF_ERR_ARG, "bad .faidx line geometry (lineBases=%u lineWidth=%u)", seqs[s].line_bases, seqs[s].line_width);
        const uint64_t step = std::max<uint64_t>(piece_bases / lb, 1) * lb;
        for (uint64_t b0 = lo[s] / lb * lb; b0 < hi[s]; b0 += step) pieces.push_back(Piece{s, b0, std::min(b0 + step, hi[s])});
    }
    first_piece[n_seqs] = (uint32_t)pieces.size();
    auto piece_of = [&](uint32_t s, uint64_t base) { // index of the piece of sequence s holding `base`
        const uint64_t lb = seqs[s].line_bases;
        const uint64_t step = std::max<uint64_t>(piece_bases / lb, 1) * lb;
        return first_piece[s] + (uint32_t)((base - lo[s] / lb * lb) / step);
    };
    // ---- windows grouped by the last piece they need; segments cut at piece boundaries and re-based
    std::vector<std::vector<uint64_t>> by_piece(pieces.size());
    std::vector<uint8_t> touched(pieces.size(), 0); // gene / transcript windows leave stretches between them unread
    for (uint64_t wi = w0; wi < w1; ++wi) {
        uint32_t last = 0;
        const kcf_window_t &win = wins[wi];
        for (uint32_t s = 0; s < win.n_segs; ++s) {
            const kcf_segment_t &sg = segs[win.first_seg + s];
            const uint32_t pa = piece_of((uint32_t)sg.seq_id, (uint64_t)sg.start0), pb = piece_of((uint32_t)sg.seq_id, (uint64_t)sg.start0 + sg.len - 1);
            for (uint32_t q = pa; q <= pb; ++q) touched[q] = 1;
            last = std::max(last, pb);
        }
        by_piece[last].push_back(wi);
    }
    std::vector<kcf_plan *> plans;
    std::vector<std::pair<size_t, size_t>> plan_rows; // (piece, windows) of every plan
    std::vector<kcf_window_t> lw;
    std::vector<kcf_segment_t> ls;
    auto cleanup = [&]() {
        for (kcf_plan *p : plans) kcf_plan_destroy(p);
    };
    for (size_t pi = 0; pi < pieces.size(); ++pi) {
        Piece &pc = pieces[pi];
        if (!touched[pi]) continue;
        const kcf_host_seq_t &sq = seqs[pc.seq];
        const uint64_t lb = sq.line_bases, lwid = sq.line_width;
        const uint64_t byte0 = pc.base0 / lb * lwid;
        // through the terminator of the line holding the piece's last base; the sequence's own end is kept as it is (a
        // missing final newline stays missing: FastaIndex.java:175-177 is then reported by kcf_plan_create)
        const uint64_t byte1 = std::min<uint64_t>(sq.n_bytes, ((pc.base1 - 1) / lb + 1) * lwid);
        if (byte0 >= byte1) return cleanup(), kcf_fail(ctx, KCF_ERR_FASTA, "sequence bytes end before base %llu", (unsigned long long)pc.base0);
        rc = kcf_ref_add_async(ctx, sq.bytes + byte0, byte1 - byte0, sq.line_bases, sq.line_width, pc.base1 - pc.base0, &pc.local_id);
        if (rc != KCF_OK) return cleanup(), rc;
        if (by_piece[pi].empty()) continue;
        lw.clear();
        ls.clear();
        for (uint64_t wi : by_piece[pi]) {
            const kcf_window_t &win = wins[wi];
            const uint32_t first = (uint32_t)ls.size();
            for (uint32_t s = 0; s < win.n_segs; ++s) {
                const kcf_segment_t &sg = segs[win.first_seg + s];
                uint64_t a = (uint64_t)sg.start0;
                const uint64_t b = a + (uint64_t)sg.len;
                while (a < b) { // one local segment per piece the segment runs through
                    const Piece &q = pieces[piece_of((uint32_t)sg.seq_id, a)];
                    const uint64_t e = std::min(b, q.base1);
                    ls.push_back(kcf_segment_t{q.local_id, (int32_t)(a - q.base0), (int32_t)(e - a)});
                    a = e;
                }
            }
            lw.push_back(kcf_window_t{first, (uint32_t)ls.size() - first});
        }
        kcf_plan *plan = nullptr;
        rc = kcf_plan_create(ctx, db->info.kmer_length, lw.data(), lw.size(), ls.data(), ls.size(), &plan);
        if (rc != KCF_OK) return cleanup(), rc;
        plans.push_back(plan);
        plan_rows.push_back({pi, lw.size()});
        rc = kcf_plan_run(ctx, db, plan, min_count, w);
        if (rc != KCF_OK) return cleanup(), rc;
    }
    // ---- all rows back with ONE wait: every plan's rows and status word are copied into a pinned landing area, then the
    // stream is synchronised once (a fetch per plan would cost a host round trip per upload piece)
    {
        size_t total = 0;
        for (auto &pr : plan_rows) total += pr.second * sizeof(kcf_result_t) + FLAG_COUNT * sizeof(uint32_t);
        if (ctx->h_rows_cap < total) {
            if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
            ctx->h_rows = nullptr;
            ctx->h_rows_cap = 0;
            const size_t cap = total + total / 4 + 4096;
            if (cudaHostAlloc(&ctx->h_rows, cap, cudaHostAllocDefault) != cudaSuccess)
                return cleanup(), kcf_fail(ctx, KCF_ERR_NOMEM, "pinned memory for %zu result bytes", total);
            ctx->h_rows_cap = cap;
        }
        size_t off = 0;
        cudaError_t e = cudaSuccess;
        for (size_t j = 0; j < plans.size() && e == cudaSuccess; ++j) {
            const size_t nb = plan_rows[j].second * sizeof(kcf_result_t);
            e = cudaMemcpyAsync(ctx->h_rows + off, plans[j]->d_out, nb, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_rows + off + nb, plans[j]->d_flags, FLAG_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
            off += nb + FLAG_COUNT * sizeof(uint32_t);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cleanup(), kcf_fail(ctx, KCF_ERR_CUDA, "fetching the rows of a sharded job: %s", cudaGetErrorString(e));
        off = 0;
        for (size_t j = 0; j < plans.size(); ++j) {
            const size_t nb = plan_rows[j].second * sizeof(kcf_result_t);
            const kcf_result_t *rows = reinterpret_cast<const kcf_result_t *>(ctx->h_rows + off);
            uint32_t flags[FLAG_COUNT];
            memcpy(flags, ctx->h_rows + off + nb, sizeof flags);
            // Data.java:101-103 — evaluated only for windows that reach the formula, left to right in double
            if (flags[FLAG_SCORE_USED] && w[0] + w[1] + w[2] != 1.0) return cleanup(), kcf_fail(ctx, KCF_ERR_WEIGHTS, "Weights should sum to 1.0");
            const std::vector<uint64_t> &ids = by_piece[plan_rows[j].first];
            for (size_t i = 0; i < ids.size(); ++i) out[ids[i]] = rows[i];
            off += nb + FLAG_COUNT * sizeof(uint32_t);
        }
    }
    cleanup();
    return KCF_OK;
}

extern "C" int kcf_screen_sharded(kcf_ctx *const *ctxs, kcf_db *const *dbs, int n, const kcf_host_seq_t *seqs, uint32_t n_seqs,
                                  const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs, uint64_t n_segs, int32_t min_count,
                                  const double w[3], kcf_result_t *out)
{
    if (!ctxs || !dbs || n < 1 || !ctxs[0]) return KCF_ERR_ARG;
    kcf_ctx *c0 = ctxs[0];
    if ((!seqs && n_seqs) || (!wins && n_wins) || (!segs && n_segs) || !w || (!out && n_wins)) return kcf_fail(c0, KCF_ERR_ARG, "kcf_screen_sharded: null argument");
    if (min_count < 1) return kcf_fail(c0, KCF_ERR_ARG, "Minimum kmer count should be at least 1"); // GetVariants.java:383-385
    for (int g = 0; g < n; ++g) {
        if (!ctxs[g] || !dbs[g] || dbs[g]->ctx != ctxs[g]) return kcf_fail(c0, KCF_ERR_ARG, "kcf_screen_sharded: dbs[%d] was not opened on ctxs[%d]", g, g);
        if (dbs[g]->part_world > 1) return kcf_fail(c0, KCF_ERR_ARG, "kcf_screen_sharded needs the whole database on every context (placement 0)");
        if (dbs[g]->info.kmer_length != dbs[0]->info.kmer_length) return kcf_fail(c0, KCF_ERR_ARG, "databases of different k on the contexts");
        for (int h = 0; h < g; ++h)
            if (ctxs[h] == ctxs[g]) return kcf_fail(c0, KCF_ERR_ARG, "kcf_screen_sharded: context %d given twice", g);
    }
    std::vector<uint64_t> bounds((size_t)n + 1);
    int rc = kcf_shard_windows(wins, n_wins, segs, n_segs, n, bounds.data());
    if (rc != KCF_OK) return kcf_fail(c0, rc, "kcf_screen_sharded: window %s", "segment range outside segs[]");
    std::vector<ShardJob> jobs((size_t)n);
    for (int g = 0; g < n; ++g) jobs[g] = ShardJob{ctxs[g], dbs[g], bounds[g], bounds[g + 1], KCF_OK};
    auto work = [&](int g) {
        ShardJob &j = jobs[g];
        j.rc = kcf_screen_shard(j.ctx, j.db, seqs, n_seqs, wins, j.w0, j.w1, segs, n_segs, min_count, w, out);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < n; ++g) th.emplace_back(work, g); // one host thread per GPU (SURVEY §8b "Threading")
    work(0);
    for (std::thread &t : th) t.join();
    for (int g = 0; g < n; ++g)
        if (jobs[g].rc != KCF_OK) {
            if (g > 0) c0->err = ctxs[g]->err; // the caller asks ctxs[0] for the message
            return jobs[g].rc;
        }
    return KCF_OK;
}
