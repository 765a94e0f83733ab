// kcf_screen.cu — K3 (rolling canonical k-mers + table probe), K4 (gap monoid reduction) and
// K5 (score): the replacement for GetVariants.processWindow (GetVariants.java:202-261) and what it
// calls per k-mer (Fasta.getKmersList Fasta.java:90-127, new Kmer(k, bothStrands) Kmer.java:57-79,
// KMC.getCount KMC.java:292-326), for all windows in one launch.
//
// Work unit = a tile of KCF_TILE consecutive positions of one window (a window is the concatenation
// of its segments, GTF.java:240-244 / Window.java:224-226), taken by ONE WARP and walked in chunks of
// KCF_CHUNK positions; lanes own consecutive k-mer end positions.  The design is described above the
// kernel.  No per-k-mer data ever goes back to HBM.
#include <algorithm>
#include <cstring>
#include "kcf_internal.cuh"
#include "kcf_lookup.cuh"
#include "kcf_gap.cuh"


struct KcfScreenParams {
    const KcfSeqDev *seqs;
    const kcf_window_t *wins;
    const kcf_segment_t *segs;
    const uint32_t *seg_off;
    const uint32_t *win_len;
    const uint64_t *tile_first;
    uint64_t n_wins;
    uint64_t tile_begin, tile_end;
    const uint8_t *table;
    const KcfStashEntry *stash;
    KcfGap *tile_sum;
    unsigned long long *tile_counter;
    int32_t min_count;
    int32_t *counts_out;   // optional: per position count of the tiles processed (-1 = no k-mer ends here)
    uint64_t counts_tile0; // tile whose position 0 maps to counts_out[0]
    // EXTRACT mode (partitioned databases): instead of probing, write every position's canonical k-mer, global home
    // line (0xFFFFFFFF where no k-mer ends) and the validity / stretch-start bitmaps, indexed from tile_begin
    unsigned long long *x_keys;
    uint32_t *x_homes, *x_okw, *x_start;
    // OWNED mode (partitioned databases, scan placement): every rank walks every tile but probes only the k-mers whose
    // home line it holds; the hit bitmap (one word per 32 positions) and the tile's Σcount go out for the reduction over
    // ranks, next to the validity / stretch-start bitmaps (identical on every rank)
    uint32_t *x_hit;
    unsigned long long *x_sum;
    // XSEND mode (partitioned databases, exchange over peer memory): every k-mer is appended to its owner's inbox
    KcfXgDev xg;
};

enum { KCF_MODE_SCREEN = 0, KCF_MODE_COUNTS = 1, KCF_MODE_EXTRACT = 2, KCF_MODE_OWNED = 3, KCF_MODE_XSEND = 4 };

// ------------------------------------------------------------------------------------------------------------
// K3/K4.  One warp takes a tile (KCF_TILE consecutive positions of one window) and walks it in chunks of KCF_CHUNK
// positions.  Inside a chunk LANES own consecutive positions (position = 32 j + lane in iteration j), so the
// ~(w+1)/2 neighbouring k-mers that share a minimizer — hence a 128-byte home line of the table — sit in the same
// load instruction and the L1 coalescer turns their probes into ONE request for that line: the de-duplication that
// makes the table's locality pay is done by the memory pipeline, not by code.  Nothing synchronises wider than a
// warp; the warps of an SM drift apart and hide each other's latencies.
//
//   stage    the two bit planes + validity bits of the chunk (+ 64-base halo): one 32-base word triple per lane,
//            cut out of the packed sequence with funnel shifts (segment junctions and window edges piecewise)
//   hash     order hash of the m-mer ending at every position: a lane owns 16 consecutive positions, holds their
//            bases in four registers and needs 2 shifts + 1 bit reversal per plane and position; then the sliding
//            minimum over w of them in registers (four minima per lane and round)
//   probe    per position: k-mer planes by funnel shift, other strand by bit reversal, minimizer -> home line, the
//            line's S low key words + filter + mask word requested together, confirm on the high word, read the
//            count; k-mers whose home mask names other lines go to a queue
//   queue    searched one item per lane, densely (continuation lines are rare per k-mer but not per warp)
//   fold     hit / valid bitmaps (one ballot per 32 positions) -> gap summary per word by bit tricks -> warp sums plus one
//            neighbour exchange for the miss runs that cross words (kcf_gap.cuh);
//            the tile's running summary lives in shared memory, not in registers
// ------------------------------------------------------------------------------------------------------------
#ifndef KCF_CHUNK
#define KCF_CHUNK 512
#endif
#ifndef KCF_MIN_WARPS
#define KCF_MIN_WARPS 40 // resident warps per SM the register allocation is held to (measured best of 32 / 40 / 48)
#endif
#define KCF_WPC 2 // independent warps per CTA (an SM holds 32 CTAs at most)
#define S_WORDS ((KCF_CHUNK + KCF_HALO) / 32 + 2) // staged 32-base words; two spare ones: the funnel shifts read past the last
#define S_HASH_WORDS (((KCF_CHUNK + KCF_HALO + 127) / 128) * 128 + 16) // padded: the 4-wide sliding-minimum rounds read 16 words per lane
#ifndef KCF_QCAP
#define KCF_QCAP 64
#endif
static_assert(KCF_CHUNK == 512 && KCF_HALO == 64, "the hash phase gives every lane 16 positions of a 512-position chunk; the halo is two words");

// a k-mer whose home line names other lines that may hold it: searched later, one item per lane
template <int KW>
struct KcfQueueItemT {
    unsigned long long key;
    uint32_t home;  // home line
    uint32_t info;  // chunk position << 16 | home mask (bit 0 cleared)
};
template <>
struct KcfQueueItemT<2> {
    unsigned long long key;
    unsigned long long key_hi; // plane 1 of a 128-bit key (k > 32)
    uint32_t home;
    uint32_t info;
};

// Per-warp shared memory.  Its size decides the L1 the SM has left: 20 CTAs of two warps at <= 9,000 bytes keep the carve-out at
// 196 KB (60 KB of L1, which the second round of a probe — high word and count of the matching slot — hits); 1 KB more per CTA
// tips it to 228 KB and costs 5 % of the step (measured, profiles/README.md).  Hence 16-byte queue items for 64-bit keys.
template <int KW>
struct __align__(16) KcfWarpSmemT {
    KcfQueueItemT<KW> queue[KCF_QCAP];
    uint32_t hash[S_HASH_WORDS];
    uint2 planes[S_WORDS];          // .x = bit 0, .y = bit 1 of the base codes; staged position q = window position o - KCF_HALO + q
    uint32_t valid[S_WORDS];
    uint32_t hit[KCF_CHUNK / 32];   // bit = k-mer observed (count >= min_count)
    uint32_t okw[KCF_CHUNK / 32];   // bit = a k-mer ends at this position
    uint32_t start[KCF_CHUNK / 32]; // bit = k-mer opens a valid stretch (EFFLEN)
    KcfGap acc;                     // summary of the tile's chunks done so far
    uint32_t xg_next[KCF_XG_MAX_WORLD], xg_end[KCF_XG_MAX_WORLD]; // XSEND: this warp's slab of entries in every owner's inbox region
};

// 32 consecutive window positions [P, P + 32) as plane / validity words (bit i = position P + i).  Positions outside the
// window read as invalid: k-mers never cross a window's edges (Window.java:224-226).  A window is the concatenation of its
// segments (GTF.java:240-244), so a word may be assembled from several of them.
__device__ __forceinline__ void kcf_stage_word(const KcfScreenParams &p, const kcf_window_t &win, uint32_t wlen, int32_t P, uint32_t &o0,
                                               uint32_t &o1, uint32_t &ov)
{
    o0 = o1 = ov = 0;
    int32_t a = max(P, 0);
    const int32_t e = (int32_t)min((int64_t)P + 32, (int64_t)wlen);
    if (a >= e) return;
    uint32_t s = 0;
    if (win.n_segs > 1) {
        uint32_t lo = 0, hi = win.n_segs; // last segment with seg_off <= a
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.seg_off[win.first_seg + mid] <= (uint32_t)a) lo = mid + 1;
            else hi = mid;
        }
        s = lo - 1;
    }
    while (a < e) {
        const uint32_t so = win.n_segs > 1 ? p.seg_off[win.first_seg + s] : 0u;
        const uint32_t send = s + 1 < win.n_segs ? p.seg_off[win.first_seg + s + 1] : wlen;
        const int32_t b = min(e, (int32_t)send);
        const kcf_segment_t sg = p.segs[win.first_seg + s];
        const KcfSeqDev sq = p.seqs[sg.seq_id];
        const uint32_t sp = (uint32_t)sg.start0 + ((uint32_t)a - so);
        const uint32_t wi = sp >> 5, sh = sp & 31u;
        const uint2 c0 = __ldg(reinterpret_cast<const uint2 *>(sq.codes) + wi), c1 = __ldg(reinterpret_cast<const uint2 *>(sq.codes) + wi + 1);
        const uint32_t v0 = __ldg(sq.valid + wi), v1 = __ldg(sq.valid + wi + 1);
        const uint32_t nb = (uint32_t)(b - a), sl = (uint32_t)(a - P);
        const uint32_t m = nb >= 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
        o0 |= (__funnelshift_r(c0.x, c1.x, sh) & m) << sl;
        o1 |= (__funnelshift_r(c0.y, c1.y, sh) & m) << sl;
        ov |= (__funnelshift_r(v0, v1, sh) & m) << sl;
        a = b;
        ++s;
    }
}

// order hashes of the m-mers ending at the 16 staged positions q0 .. q0 + 15 (q0 a multiple of 16, q0 >= m - 1): the lane
// holds the bases it needs — staged bits [q0 - m + 1, q0 + 16) of both planes — in two 64-bit windows
template <typename WS>
__device__ __forceinline__ void kcf_hash16(WS &W, uint32_t q0, uint32_t m, uint32_t mm)
{
    const uint32_t s0 = q0 + 1u - m, wi = s0 >> 5, off = s0 & 31u;
    const uint2 a = W.planes[wi], b = W.planes[wi + 1], c = W.planes[wi + 2];
    const uint32_t lo0 = __funnelshift_r(a.x, b.x, off), hi0 = __funnelshift_r(b.x, c.x, off);
    const uint32_t lo1 = __funnelshift_r(a.y, b.y, off), hi1 = __funnelshift_r(b.y, c.y, off);
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = kcf_mmer_order(__funnelshift_r(lo0, hi0, i), __funnelshift_r(lo1, hi1, i), m, mm);
    uint4 *dst = reinterpret_cast<uint4 *>(&W.hash[q0]);
    dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
    dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
    dst[2] = make_uint4(h[8], h[9], h[10], h[11]);
    dst[3] = make_uint4(h[12], h[13], h[14], h[15]);
}

// SPEC = the common geometry compiled in: both-strands database and 4 <= w <= 13 (register sliding minimum); SPEC = false
// reads both from the geometry at run time (any m, non-canonical databases)
template <int S, int MODE, bool SPEC>
__global__ void __launch_bounds__(32 * KCF_WPC, KCF_MIN_WARPS / KCF_WPC) kcf_screen_kernel(const KcfScreenParams p, const KcfTableGeom g)
{
    constexpr bool COUNTS = MODE == KCF_MODE_COUNTS, XSEND = MODE == KCF_MODE_XSEND, EXTRACT = MODE == KCF_MODE_EXTRACT || XSEND, OWNED = MODE == KCF_MODE_OWNED;
    constexpr int KW = S <= 7 ? 2 : 1; // 128-bit keys (k = 33 .. 64): 7 / 6 slots per line, home line by a hash of the key
    static_assert(KW == 1 || (!EXTRACT && !OWNED), "partitioned tables move 64-bit keys");
    __shared__ __align__(16) KcfWarpSmemT<KW> kcf_warp_smem[KCF_WPC];
    KcfWarpSmemT<KW> &W = kcf_warp_smem[threadIdx.x >> 5]; // the warps of a CTA share nothing

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t k = g.k;
    const bool FASTMIN = SPEC || (g.w >= 4 && g.w <= 13); // register sliding minimum (warp uniform)
    const bool BOTH = SPEC || g.both_strands != 0;
    uint32_t P2 = 1; // largest power of two <= w: the general sliding minimum is built by doubling up to it
    if (!SPEC)
        while (2 * P2 <= g.w) P2 *= 2;

    if (XSEND && lane < KCF_XG_MAX_WORLD) W.xg_next[lane] = W.xg_end[lane] = 0;
    __syncwarp();
    for (;;) {
        // ---- take a tile: KCF_TILE consecutive positions of one window ----
        uint64_t tile = 0;
        uint32_t w = 0;
        if (lane == 0) {
            tile = p.tile_begin + atomicAdd(p.tile_counter, 1ULL);
            if (tile < p.tile_end) {
                uint64_t lo = 0, hi = p.n_wins; // window owning the tile: last w with tile_first[w] <= tile
                while (lo < hi) {
                    uint64_t mid = (lo + hi) >> 1;
                    if (p.tile_first[mid] <= tile) lo = mid + 1;
                    else hi = mid;
                }
                w = (uint32_t)(lo - 1);
            }
        }
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= p.tile_end) break;
        w = __shfl_sync(0xffffffffu, w, 0);
        const kcf_window_t win = p.wins[w];
        const uint32_t wlen = p.win_len[w];
        const int32_t o_tile = (int32_t)((tile - p.tile_first[w]) * KCF_TILE); // window position of the tile's first k-mer end

        if (lane == 0) W.acc = kcf_gap_zero();
        uint32_t carry_hash = 0; // order hash of staged position KCF_CHUNK + 32 + lane of the previous chunk = position 32 + lane of this one
        unsigned long long owned_sum = 0; // OWNED mode: Σcount of this rank's hits in the tile

        for (uint32_t chunk = 0; chunk < KCF_TILE / KCF_CHUNK; ++chunk) {
            const int32_t o = o_tile + (int32_t)(chunk * KCF_CHUNK);
            if (o >= (int32_t)wlen) break;

            // ---- stage window positions [o - HALO, o + CHUNK): one word triple per lane; the halo of a later chunk is the
            // previous chunk's tail ----
            {
                uint2 tp = make_uint2(0u, 0u);
                uint32_t tv = 0;
                if (chunk > 0 && lane < KCF_HALO / 32) {
                    tp = W.planes[KCF_CHUNK / 32 + lane];
                    tv = W.valid[KCF_CHUNK / 32 + lane];
                }
                __syncwarp();
                if (lane < S_WORDS) {
                    if (lane >= (KCF_CHUNK + KCF_HALO) / 32) tp = make_uint2(0u, 0u), tv = 0; // the spare words
                    else if (chunk == 0 || lane >= KCF_HALO / 32) kcf_stage_word(p, win, wlen, o - KCF_HALO + 32 * (int32_t)lane, tp.x, tp.y, tv);
                    W.planes[lane] = tp;
                    W.valid[lane] = tv;
                }
            }
            __syncwarp();

            // ---- order hash of the m-mer ending at every staged position from 32 on (the minimizer window of the chunk's
            // first k-mer starts at 65 - w >= 33) ----
            if (KW == 1) {
            if (chunk > 0) W.hash[32 + lane] = carry_hash;
#pragma unroll 1
            for (uint32_t r = chunk == 0 ? 0u : 1u; r < 2; ++r) // r = 0: the halo positions 32 .. 63 of a tile's first chunk (two lanes)
                if (r == 1 || lane < 2) kcf_hash16(W, r == 0 ? 32 + 16 * lane : KCF_HALO + 16 * lane, g.m, g.mm);
            __syncwarp();
            carry_hash = W.hash[KCF_CHUNK + 32 + lane];
            // sliding minimum over w consecutive order hashes, in place: afterwards hash[q] = min over [q, q + w) (fast path) or
            // min over [q, q + P2) (general path; the probe then combines two of them)
            if (FASTMIN) {
                // 4 <= w <= 13 (the automatic choice w = S - 2 always is): a lane produces 4 consecutive minima per round from
                // the 13 hashes it holds in registers plus 3 more; the part common to the four windows is reduced once.
                // A round rewrites [128 r, 128 r + 128) after every lane has read what it needs from it; later rounds only
                // read higher positions.
                const uint32_t wv = g.w;
#pragma unroll 1
                for (uint32_t r0 = 0; r0 < KCF_CHUNK + KCF_HALO; r0 += 128) {
                    const uint32_t base = r0 + 4 * lane;
                    const uint4 va = *reinterpret_cast<const uint4 *>(&W.hash[base]);
                    const uint4 vb = *reinterpret_cast<const uint4 *>(&W.hash[base + 4]);
                    const uint4 vc = *reinterpret_cast<const uint4 *>(&W.hash[base + 8]);
                    const uint32_t v12 = W.hash[base + 12];
                    const uint32_t t0 = W.hash[base + wv], t1 = W.hash[base + wv + 1], t2 = W.hash[base + wv + 2];
                    uint32_t core = va.w; // positions 3 .. w-1 belong to all four windows
                    if (wv > 4) core = min(core, vb.x);
                    if (wv > 5) core = min(core, vb.y);
                    if (wv > 6) core = min(core, vb.z);
                    if (wv > 7) core = min(core, vb.w);
                    if (wv > 8) core = min(core, vc.x);
                    if (wv > 9) core = min(core, vc.y);
                    if (wv > 10) core = min(core, vc.z);
                    if (wv > 11) core = min(core, vc.w);
                    if (wv > 12) core = min(core, v12);
                    uint4 ov;
                    ov.x = min(min(core, va.x), min(va.y, va.z));
                    ov.y = min(min(core, va.y), min(va.z, t0));
                    ov.z = min(min(core, va.z), min(t0, t1));
                    ov.w = min(min(core, t0), min(t1, t2));
                    __syncwarp();
                    *reinterpret_cast<uint4 *>(&W.hash[base]) = ov;
                    __syncwarp();
                }
            } else {
                uint32_t sd = 1;
                while (sd < P2) {
                    const bool four = 4 * sd <= P2;
                    const uint32_t span = four ? 3 * sd : sd;
                    __syncwarp();
#pragma unroll 1
                    for (uint32_t q = lane; q + span < KCF_CHUNK + KCF_HALO; q += 32) {
                        uint32_t v = min(W.hash[q], W.hash[q + sd]);
                        if (four) v = min(v, min(W.hash[q + 2 * sd], W.hash[q + 3 * sd]));
                        __syncwarp(__activemask());
                        W.hash[q] = v;
                    }
                    sd *= four ? 4 : 2;
                }
            }
            } // KW == 1
            // validity of every k-mer of the chunk, 32 positions per lane: bit q of `run` is set iff the k staged validity
            // bits q-k+1 .. q are all set (Fasta.java:99-104), built from runs of 1, 2, 4, .. bits; a k-mer opens a valid
            // stretch when the one ending one position earlier is not valid (EFFLEN, Fasta.java:140-167)
            if (lane < KCF_CHUNK / 32) {
                if (KW == 1) {
                    const uint64_t v64 = ((uint64_t)W.valid[KCF_HALO / 32 + lane] << 32) | W.valid[KCF_HALO / 32 - 1 + lane]; // top half = this lane's 32 positions
                    uint64_t a = v64, run = ~0ULL;
                    uint32_t pos = 0;
#pragma unroll
                    for (uint32_t b = 0; b < 6; ++b) {
                        if ((k >> b) & 1u) {
                            run &= a << pos;
                            pos += 1u << b;
                        }
                        a &= a << (1u << b);
                    }
                    W.okw[lane] = (uint32_t)(run >> 32);
                    W.start[lane] = (uint32_t)((run & ~(run << 1)) >> 32);
                } else { // k up to 64 looks back 63 positions: the two halo words and this lane's word, 96 bits
                    typedef unsigned __int128 u128;
                    const u128 v = ((u128)W.valid[KCF_HALO / 32 + lane] << 64) | ((u128)W.valid[KCF_HALO / 32 - 1 + lane] << 32) | W.valid[KCF_HALO / 32 - 2 + lane];
                    u128 a = v, run = ~(u128)0;
                    uint32_t pos = 0;
#pragma unroll
                    for (uint32_t b = 0; b < 7; ++b) {
                        if ((k >> b) & 1u) {
                            run &= a << pos;
                            pos += 1u << b;
                        }
                        a &= a << (1u << b);
                    }
                    W.okw[lane] = (uint32_t)(run >> 64);
                    W.start[lane] = (uint32_t)((run & ~(run << 1)) >> 64);
                }
                W.hit[lane] = 0;
            }
            __syncwarp();

            const uint32_t npos = (uint32_t)min((int32_t)KCF_CHUNK, (int32_t)wlen - o);
            const uint32_t J = (npos + 31) / 32;
            unsigned long long sum = 0; // Σ count over this lane's observed k-mers
            uint32_t qn = 0;            // queue length (warp uniform)

            // search the queued k-mers in the lines their home masks name; one item per lane
            auto flush_queue = [&]() {
                __syncwarp();
#pragma unroll 1
                for (uint32_t t = lane; t < qn; t += 32) {
                    const KcfQueueItemT<KW> it = W.queue[t];
                    uint32_t m2 = it.info & 0x7FFEu, c2 = 0;
                    bool found = false;
                    while (m2 && !found) {
                        const uint32_t d = __ffs(m2) - 1;
                        m2 &= m2 - 1;
                        const uint8_t *ol = p.table + (uint64_t)kcf_line_wrap(it.home, d, g) * KCF_LINE_BYTES;
                        if constexpr (KW == 2) found = kcf_probe_line2<S>(ol, KcfKey2{it.key, it.key_hi}, g, c2);
                        else found = kcf_probe_line<S>(ol, it.key, c2);
                    }
                    if (!found && (it.info & (1u << KCF_STASH_BIT))) {
                        if constexpr (KW == 2) c2 = kcf_stash_find(p.stash, g, it.key, it.key_hi);
                        else c2 = kcf_stash_find(p.stash, g, it.key);
                    }
                    const uint32_t pc = it.info >> 16;
                    if ((int32_t)c2 >= p.min_count) {
                        sum += c2;
                        atomicOr(&W.hit[pc >> 5], 1u << (pc & 31u));
                    }
                    if (COUNTS) p.counts_out[(tile - p.counts_tile0) * KCF_TILE + chunk * KCF_CHUNK + pc] = (int32_t)c2;
                }
                qn = 0;
                __syncwarp();
            };

            // ---- probe: lanes own consecutive positions ----
            // minimizer of the k-mer ending at staged position q = min over the w m-mers ending at q-w+1 .. q -> home line
            auto home_of = [&](uint32_t q) {
                const uint32_t h0 = q - g.w + 1;
                return kcf_home_line(FASTMIN ? W.hash[h0] : min(W.hash[h0], W.hash[h0 + g.w - P2]), g);
            };
#pragma unroll 1
            for (uint32_t j = 0; j < J; ++j) {
                const uint32_t cpos = 32 * j + lane;  // chunk position of this lane's k-mer end
                const uint32_t q = KCF_HALO + cpos;   // the same in staged coordinates
                const bool ok = (W.okw[j] >> lane) & 1u; // a k-mer ends here (bitmap built once per chunk)
                // table key: the k-mer's bit planes; for a both-strands database the smaller strand (kcf_lookup.cuh) — what the
                // loader stored the record spelling this k-mer's canonical form under (Kmer.java:57-79)
                uint32_t klo, khi;
                uint64_t key, key_hi = 0;
                uint32_t home;
                if (KW == 1) {
                    const uint32_t b0 = q - k + 1, wi = b0 >> 5, sh = b0 & 31u;
                    const uint2 a = W.planes[wi], b = W.planes[wi + 1];
                    klo = __funnelshift_r(a.x, b.x, sh) & g.km;
                    khi = __funnelshift_r(a.y, b.y, sh) & g.km;
                    if (BOTH) kcf_plane_canonical(klo, khi, kcf_plane_rc(klo, k, g.km), kcf_plane_rc(khi, k, g.km), klo, khi);
                    key = ((uint64_t)khi << 32) | klo;
                    home = home_of(q);
                } else { // two 64-bit planes cut out of three staged words each; the home line is a hash of the whole key
                    const uint32_t b0 = q - k + 1, wi = b0 >> 5, sh = b0 & 31u;
                    const uint2 a = W.planes[wi], b = W.planes[wi + 1], c = W.planes[wi + 2];
                    KcfKey2 f;
                    f.p0 = (((uint64_t)__funnelshift_r(b.x, c.x, sh) << 32) | __funnelshift_r(a.x, b.x, sh)) & g.km64;
                    f.p1 = (((uint64_t)__funnelshift_r(b.y, c.y, sh) << 32) | __funnelshift_r(a.y, b.y, sh)) & g.km64;
                    if (BOTH) {
                        const uint64_t r0 = kcf_plane_rc64(f.p0, k, g.km64), r1 = kcf_plane_rc64(f.p1, k, g.km64);
                        if (r1 < f.p1 || (r1 == f.p1 && r0 < f.p0)) {
                            f.p0 = r0;
                            f.p1 = r1;
                        }
                    }
                    key = f.p0;
                    key_hi = f.p1;
                    klo = (uint32_t)f.p0;
                    khi = 0;
                    home = kcf_home_line2(f, g);
                }
                if (XSEND) {
                    // RUNS of k-mers sharing their home line go straight into the owners' inboxes (kcf_internal.cuh): a run's
                    // head lane packs the bases the run spans and appends one 16-byte entry; the lanes bound for one owner take
                    // consecutive entries of that owner's region for this sender (one cursor atomic per owner and warp step)
                    const uint64_t base = (tile - p.tile_begin) * KCF_TILE + (uint64_t)chunk * KCF_CHUNK;
                    const uint32_t prev_home = __shfl_up_sync(0xffffffffu, home, 1);
                    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                    const bool brk = ok && (lane == 0 || !((okmask >> (lane - 1)) & 1u) || prev_home != home); // a new home line (or the first k-mer after a gap)
                    const uint32_t brkmask = __ballot_sync(0xffffffffu, brk);
                    // runs longer than KCF_XG_RUN (two minimizers in a row hashing to one line) are cut
                    const uint32_t since = ok ? lane - (31u - __clz(brkmask & (0xFFFFFFFFu >> (31u - lane)))) : 0u;
                    const bool head = ok && (since % KCF_XG_RUN) == 0u;
                    const uint32_t headmask = __ballot_sync(0xffffffffu, head);
                    uint32_t slot = ok ? 0xFFFFFFFEu : 0xFFFFFFFFu;
                    if (head) {
                        const uint32_t above = lane == 31u ? 0u : ((headmask | ~okmask) & (0xFFFFFFFEu << lane)); // next head or next position without a k-mer
                        const uint32_t len = (above ? __ffs(above) - 1u : 32u) - lane;
                        // the k + len - 1 bases from the run's first base on, both planes
                        const uint32_t b0 = q - k + 1, wi = b0 >> 5, sh = b0 & 31u;
                        const uint2 pa = W.planes[wi], pb = W.planes[wi + 1], pc = W.planes[wi + 2];
                        const uint64_t nm = (1ULL << (k + len - 1u)) - 1ULL;
                        const uint64_t p0 = (((uint64_t)__funnelshift_r(pb.x, pc.x, sh) << 32) | __funnelshift_r(pa.x, pb.x, sh)) & nm;
                        const uint64_t p1 = (((uint64_t)__funnelshift_r(pb.y, pc.y, sh) << 32) | __funnelshift_r(pa.y, pb.y, sh)) & nm;
                        const uint32_t owner = kcf_line_owner_mapped(home, p.xg);
                        // the heads bound for one owner find each other with one match instruction (a loop over the owners
                        // present cost one round of shuffles and ballots per owner) and take
                        // consecutive entries of this warp's slab of that owner's region; a slab that cannot take the step's
                        // runs is closed (its rest marked empty) and a new one drawn from the region's cursor: one global
                        // atomic per KCF_XG_SLAB runs and owner instead of one per warp step — every warp of the GPU adds to the
                        // same `world` counters, and same-address atomics serialise (measured: 68 ms of a 98 ms step)
                        const uint32_t heads = headmask; // every run head of this step is here
                        const uint32_t same = __match_any_sync(heads, owner);
                        const uint32_t leader = __ffs(same) - 1, need = __popc(same), rk = __popc(same & ((1u << lane) - 1u));
                        uint32_t off = W.xg_next[owner];
                        const uint32_t end = W.xg_end[owner];
                        if (off + need > end) { // uniform over the owner's heads
                            for (uint32_t h = off + rk; h < end; h += need)
                                if (h < p.xg.cap) p.xg.in_runs[owner][h] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                            uint32_t nb = 0;
                            if (lane == leader) nb = atomicAdd(&p.xg.cursor[owner], (unsigned int)KCF_XG_SLAB);
                            off = __shfl_sync(same, nb, leader); // also: every head of the group has read the old slab bounds by now
                            if (lane == leader) W.xg_end[owner] = off + KCF_XG_SLAB;
                        }
                        __syncwarp(heads); // the bounds are read by all before any leader moves them
                        if (lane == leader) W.xg_next[owner] = off + need;
                        const uint32_t idx = off + rk;
                        if (idx < p.xg.cap) {
                            uint64_t ea, eb;
                            kcf_xg_pack_run(p0, p1, len, home, ea, eb);
                            p.xg.in_runs[owner][idx] = make_uint4((uint32_t)ea, (uint32_t)(ea >> 32), (uint32_t)eb, (uint32_t)(eb >> 32));
                            slot = (owner << 28) | idx;
                        } else {
                            p.xg.flags[0] = 1u; // region full: the host reports it (the run reads as absent meanwhile)
                            slot = 0xFFFFFFFFu;
                        }
                    }
                    __syncwarp();
                    p.xg.pos_slot[base + cpos] = slot;
                    if (lane == 0) {
                        p.xg.okw[(base >> 5) + j] = W.okw[j];
                        p.xg.start[(base >> 5) + j] = W.start[j];
                    }
                    continue;
                }
                if (EXTRACT) {
                    const uint64_t base = (tile - p.tile_begin) * KCF_TILE + (uint64_t)chunk * KCF_CHUNK;
                    p.x_keys[base + cpos] = key;
                    p.x_homes[base + cpos] = ok ? home : 0xFFFFFFFFu;
                    if (lane == 0) {
                        p.x_okw[(base >> 5) + j] = W.okw[j];
                        p.x_start[(base >> 5) + j] = W.start[j];
                    }
                    continue;
                }
                const uint32_t lhome = OWNED ? home - (uint32_t)g.line_lo : home; // local line index
                const bool mine = !OWNED || lhome < (uint32_t)g.n_local;          // this rank holds the k-mer's home line
                const uint8_t *L = p.table + (uint64_t)lhome * KCF_LINE_BYTES;
                uint32_t cnt = 0, mask = 0;
                bool pending = false;
                if (ok && mine) {
                    const bool inl = klo != KCF_EMPTY_LO;
                    // filter and mask words of the line travel with its key words: a miss needs them, and asking for them
                    // only after the compare would add a dependent round trip
                    const uint32_t w31 = __ldg(reinterpret_cast<const uint32_t *>(L) + 31);
                    const unsigned long long fword = S == 13 ? __ldg(reinterpret_cast<const unsigned long long *>(L + 104))
                                                             : (unsigned long long)__ldg(reinterpret_cast<const uint32_t *>(L + 120));
                    bool found;
                    if constexpr (KW == 2) found = inl && kcf_probe_line2<S>(L, KcfKey2{key, key_hi}, g, cnt);
                    else found = inl && kcf_probe_line<S>(L, key, cnt);
                    if (!found) {
                        cnt = 0;
                        // absent unless the home line's filter says a key like this one lives outside it
                        bool maybe;
                        if (KW == 2) maybe = g.fbits == 0 || kcf_filter_pass32((uint32_t)fword, key ^ (key_hi * 0x9E3779B97F4A7C15ULL));
                        else maybe = S == 13 ? kcf_filter_pass64(fword, key) : kcf_filter_pass32((uint32_t)fword, key);
                        if (maybe) {
                            mask = kcf_mask_from_word31(w31);
                            if (inl && (mask & 0x7FFEu)) pending = true;
                            else if ((mask >> KCF_STASH_BIT) & 1u) cnt = kcf_stash_find(p.stash, g, key, key_hi);
                        }
                    }
                }
                const bool hit = ok && (int32_t)cnt >= p.min_count; // Java int compare (GetVariants.java:224)
                if (hit) sum += cnt;
                const uint32_t hb = __ballot_sync(0xffffffffu, hit);
                const uint32_t pb = __ballot_sync(0xffffffffu, pending);
                if (lane == 0) W.hit[j] = hb; // the queue flush ORs late hits into it, after this store
                if (COUNTS) p.counts_out[(tile - p.counts_tile0) * KCF_TILE + chunk * KCF_CHUNK + cpos] = ok ? (int32_t)cnt : -1;
                if (pb) {
                    if (pending) {
                        KcfQueueItemT<KW> it;
                        it.key = key;
                        if constexpr (KW == 2) it.key_hi = key_hi;
                        it.home = home;
                        it.info = (cpos << 16) | (mask & 0xFFFEu);
                        W.queue[qn + __popc(pb & ((1u << lane) - 1u))] = it;
                    }
                    qn += __popc(pb);
                    if (qn + 32 > KCF_QCAP) flush_queue(); // nearly full: search it now
                }
            }
            if (EXTRACT) continue;
            if (qn) flush_queue();
            __syncwarp();
            if (OWNED) {
                // this rank's share of the chunk: bitmaps out, Σcount kept per tile; the gap summaries are folded after the
                // reduction over ranks (kcf_scan_fold)
                const uint64_t wbase = (tile - p.tile_begin) * (KCF_TILE / 32) + (uint64_t)chunk * (KCF_CHUNK / 32);
                if (lane < KCF_CHUNK / 32) {
                    p.x_hit[wbase + lane] = W.hit[lane];
                    p.x_okw[wbase + lane] = W.okw[lane];
                    p.x_start[wbase + lane] = W.start[lane];
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
                owned_sum += sum;
                __syncwarp();
                continue;
            }

            // ---- fold: lane j summarises positions [32 j, 32 j + 32), one ordered shuffle reduction per chunk ----
            constexpr uint32_t NWORDS = KCF_CHUNK / 32;
            KcfGap a = kcf_gap_fold_warp(lane < NWORDS ? W.hit[lane] : 0u, lane < NWORDS ? W.okw[lane] : 0u, lane < NWORDS ? W.start[lane] : 0u, lane, k);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d); // Σ count is not tied to positions: plain warp sum
            if (lane == 0) {
                a.sum = sum;
                W.acc = kcf_gap_combine(W.acc, a, k);
            }
            __syncwarp(); // the bitmaps are rewritten by the next chunk
        }
        if (OWNED) {
            if (lane == 0) p.x_sum[tile - p.tile_begin] = owned_sum;
        } else if (!EXTRACT && !COUNTS && lane == 0) p.tile_sum[tile] = W.acc; // the COUNTS pass (min_count = 1) leaves the run's tile sums alone
        __syncwarp();
    }
    if (XSEND) { // the unused rest of this warp's slabs reads "no run"
        __syncwarp();
        for (uint32_t o = 0; o < p.xg.world; ++o)
            for (uint32_t h = W.xg_next[o] + lane; h < W.xg_end[o]; h += 32)
                if (h < p.xg.cap) p.xg.in_runs[o][h] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
}

// K4 tail + K5: combine a window's tiles in order and emit the KCF integers and the score
// (GetVariants.java:247-258, Data.java:70-107).  No FMA contraction: explicit _rn intrinsics.
__global__ void kcf_finalize_kernel(const KcfGap *__restrict__ tile_sum, const uint64_t *__restrict__ tile_first, uint64_t n_wins,
                                    uint32_t k, double wi, double wt, double wr, kcf_result_t *__restrict__ out, uint32_t *flags)
{
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_wins) return;
    KcfGap r;
    r.n = r.obs = r.lead = r.trail = r.vin = r.inner = r.has = r.starts = 0;
    r.sum = 0;
    for (uint64_t t = tile_first[w]; t < tile_first[w + 1]; ++t) r = kcf_gap_combine(r, tile_sum[t], k);
    kcf_result_t o;
    o.total_kmers = (int32_t)r.n;
    o.eff_len = (int32_t)(r.n + (k - 1) * r.starts); // Fasta.java:140-167: stretches of >= k valid bases
    o.obs = (int32_t)r.obs;
    o.kmer_count_sum = (int64_t)r.sum;
    o._pad = 0;
    if (r.has) {
        o.left = (int32_t)r.lead;
        o.right = (int32_t)r.trail;
        o.inner = (int32_t)r.inner;
        o.variations = (int32_t)(r.vin + (r.lead > 0) + (r.trail > 0));
    } else { // no hit at all: one trailing gap (GetVariants.java:247-251)
        o.left = 0;
        o.right = (int32_t)r.n;
        o.inner = 0;
        o.variations = r.n > 0 ? 1 : 0;
    }
    double score = 0.0;
    if (!(o.obs == 0 || o.total_kmers == 0 || o.eff_len == 0)) { // Data.java:96-98
        atomicOr(&flags[FLAG_SCORE_USED], 1u);
        const double eff = (double)o.eff_len;
        const double ta = __dmul_rn(wr, __ddiv_rn((double)o.obs, (double)o.total_kmers));
        const double tb = __dmul_rn(wi, __dsub_rn(1.0, __ddiv_rn((double)o.inner, eff)));
        const double tc = __dmul_rn(wt, __dsub_rn(1.0, __ddiv_rn((double)(o.left + o.right), eff)));
        score = __dmul_rn(__dadd_rn(__dadd_rn(ta, tb), tc), 100.0); // Data.java:104-106
    }
    o.score = score;
    out[w] = o;
}

// one 32-byte sector with a single 256-bit load that bypasses L1 allocation
__device__ __forceinline__ void kcf_ld_bucket(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// ---- random 32-byte sector gather microbenchmark (the random-access roofline of SURVEY §8(d)) ----
__global__ void __launch_bounds__(256) kcf_random_sector_kernel(const uint64_t *__restrict__ buf, uint64_t n_sectors, uint64_t per_thread,
                                                                uint64_t seed, uint64_t *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ULL ^ seed;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < per_thread; it += 8) {
        uint64_t a[8], b[8], c[8], d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x ^= x >> 29;
            x *= 0xBF58476D1CE4E5B9ULL;
            x ^= x >> 32;
            uint64_t s = __umul64hi(x, n_sectors);
            kcf_ld_bucket(buf + 4 * s, a[j], b[j], c[j], d[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += a[j] ^ b[j] ^ c[j] ^ d[j];
    }
    if (acc == 0x1234567ULL) *sink = acc; // keeps the loads alive
}

extern "C" int kcf_measure_random_sector_gbps(kcf_ctx *ctx, uint64_t n_bytes, uint64_t n_loads, int repeats, double *gbps_out)
{
    if (!ctx || !gbps_out || n_bytes < 4096 || n_loads == 0) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *buf = nullptr, *sink = nullptr;
    const uint64_t n_sectors = n_bytes / 32;
    KCF_CUDA(ctx, cudaMalloc(&buf, n_sectors * 32));
    cudaError_t e = cudaMalloc(&sink, 8);
    if (e != cudaSuccess) { cudaFree(buf); return kcf_fail(ctx, KCF_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(buf, 1, n_sectors * 32, ctx->stream);
    const uint64_t per_thread = 64;
    const uint64_t threads = (n_loads + per_thread - 1) / per_thread;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int r = 0; r < std::max(repeats, 1) + 1; ++r) {
        cudaEventRecord(a, ctx->stream);
        kcf_random_sector_kernel<<<grid, 256, 0, ctx->stream>>>(buf, n_sectors, per_thread, 0x51ED27ULL * (r + 1), sink);
        cudaEventRecord(b, ctx->stream);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double gbps = (double)grid * 256.0 * per_thread * 32.0 / (ms * 1e-3) / 1e9;
        if (r > 0 && gbps > best) best = gbps; // first launch is the warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    cudaFree(sink);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "random sector kernel: %s", cudaGetErrorString(e));
    *gbps_out = best;
    return KCF_OK;
}

// ---- random 128-byte LINE gather microbenchmark: the memory-side bound of THIS table (one coalesced line request per run
// of k-mers sharing a minimizer).  Groups of 4 lanes read the 4 sectors of one uniformly random line in one instruction, 8
// lines in flight per group (profiles/r1b_randline.txt, pattern T3).
__global__ void __launch_bounds__(256) kcf_random_line_kernel(const uint64_t *__restrict__ buf, uint64_t n_lines, uint64_t iters, uint64_t seed,
                                                              uint64_t *sink)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = ((t >> 2) + 1) * 0x9E3779B97F4A7C15ULL ^ seed; // the 4 lanes of a group draw the same lines
    uint64_t acc = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t a[8], b[8], c[8], d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x ^= x >> 29;
            x *= 0xBF58476D1CE4E5B9ULL;
            x ^= x >> 32;
            const uint64_t l = __umul64hi(x, n_lines);
            kcf_ld_bucket(buf + l * 16 + (t & 3) * 4, a[j], b[j], c[j], d[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += a[j] ^ b[j] ^ c[j] ^ d[j];
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

extern "C" int kcf_measure_random_line_rate(kcf_ctx *ctx, uint64_t n_bytes, uint64_t n_lines_read, int repeats, double *lines_per_s_out)
{
    if (!ctx || !lines_per_s_out || n_bytes < 4096 || n_lines_read == 0) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *buf = nullptr, *sink = nullptr;
    const uint64_t n_lines = n_bytes / 128;
    KCF_CUDA(ctx, cudaMalloc(&buf, n_lines * 128));
    cudaError_t e = cudaMalloc(&sink, 8);
    if (e != cudaSuccess) { cudaFree(buf); return kcf_fail(ctx, KCF_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(buf, 1, n_lines * 128, ctx->stream);
    const uint64_t iters = 8; // 64 lines per group of 4 lanes
    const uint64_t groups = (n_lines_read + iters * 8 - 1) / (iters * 8);
    const unsigned grid = (unsigned)((groups * 4 + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int r = 0; r < std::max(repeats, 1) + 1; ++r) {
        cudaEventRecord(a, ctx->stream);
        kcf_random_line_kernel<<<grid, 256, 0, ctx->stream>>>(buf, n_lines, iters, 0x51ED27ULL * (r + 1), sink);
        cudaEventRecord(b, ctx->stream);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)grid * 64.0 * iters * 8.0 / (ms * 1e-3); // 64 groups per CTA
        if (r > 0 && rate > best) best = rate; // first launch is the warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    cudaFree(sink);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "random line kernel: %s", cudaGetErrorString(e));
    *lines_per_s_out = best;
    return KCF_OK;
}

// ---- host side: plans ---------------------------------------------------------------------------
// bring the device-side sequence table up to date, stream-ordered and without a host synchronisation (a host that
// uploads one chromosome after the other and screens each as it arrives must not stall on earlier work)
int kcf_sync_seqs(kcf_ctx *ctx)
{
    const size_t n = ctx->seqs.size();
    if (ctx->d_seqs_cap < std::max<size_t>(n, 1)) {
        const size_t cap = std::max<size_t>(4096, 2 * n);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->d_seqs) cudaFree(ctx->d_seqs);
        if (ctx->h_seqs) cudaFreeHost(ctx->h_seqs);
        ctx->d_seqs = nullptr;
        ctx->h_seqs = nullptr;
        ctx->d_seqs_cap = 0;
        KCF_CUDA(ctx, cudaMalloc(&ctx->d_seqs, cap * sizeof(KcfSeqDev)));
        KCF_CUDA(ctx, cudaHostAlloc(&ctx->h_seqs, cap * sizeof(KcfSeqDev), cudaHostAllocDefault));
        ctx->d_seqs_cap = cap;
        ctx->seqs_uploaded = 0;
    }
    if (n < ctx->seqs_uploaded) ctx->seqs_uploaded = 0; // cleared since (kcf_ref_clear synchronises)
    if (n > ctx->seqs_uploaded) {
        for (size_t i = ctx->seqs_uploaded; i < n; ++i) {
            ctx->h_seqs[i].codes = ctx->seqs[i].codes;
            ctx->h_seqs[i].valid = ctx->seqs[i].valid;
            ctx->h_seqs[i].len = (uint32_t)ctx->seqs[i].len;
            ctx->h_seqs[i]._pad = 0;
        }
        KCF_CUDA(ctx, cudaMemcpyAsync(ctx->d_seqs + ctx->seqs_uploaded, ctx->h_seqs + ctx->seqs_uploaded,
                                      (n - ctx->seqs_uploaded) * sizeof(KcfSeqDev), cudaMemcpyHostToDevice, ctx->stream));
        ctx->seqs_uploaded = n;
    }
    ctx->seqs_dirty = false;
    return KCF_OK;
}

extern "C" void kcf_plan_destroy(kcf_plan *plan)
{
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    cudaStreamSynchronize(plan->ctx->stream);
    kcf_pool_put(plan->ctx, plan->d_block, plan->d_block_bytes); // d_wins .. d_out live in it; the stream is idle (synchronised above)
    cudaFree(plan->x_keys);
    cudaFree(plan->x_homes);
    cudaFree(plan->x_okw);
    cudaFree(plan->x_start);
    cudaFree(plan->x_cnt);
    cudaFree(plan->x_cursor);
    cudaFree(plan->s_okw);
    cudaFree(plan->s_start);
    delete plan;
}

extern "C" int kcf_plan_create(kcf_ctx *ctx, int32_t kmer_length, const kcf_window_t *wins, uint64_t n_wins,
                               const kcf_segment_t *segs, uint64_t n_segs, kcf_plan **out)
{
    if (!ctx || !out || (!wins && n_wins) || (!segs && n_segs)) return KCF_ERR_ARG;
    *out = nullptr;
    if (kmer_length < 3 || kmer_length > 64) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "k=%d outside 3..64", kmer_length);
    if (n_wins >= (1ULL << 32) || n_segs >= (1ULL << 32)) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "too many windows / segments");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> seg_off(std::max<uint64_t>(n_segs, 1), 0);
    std::vector<uint32_t> win_len(std::max<uint64_t>(n_wins, 1), 0);
    std::vector<uint64_t> tile_first(n_wins + 1, 0);
    uint64_t tiles = 0, positions = 0;
    for (uint64_t w = 0; w < n_wins; ++w) {
        const kcf_window_t &win = wins[w];
        if (win.n_segs == 0) // fasta == null (GetVariants.java:213-216)
            return kcf_fail(ctx, KCF_ERR_RANGE, "Fasta object is null for window %llu (no segment)", (unsigned long long)w);
        if ((uint64_t)win.first_seg + win.n_segs > n_segs) return kcf_fail(ctx, KCF_ERR_ARG, "window %llu: segment range outside segs[]", (unsigned long long)w);
        uint64_t len = 0;
        for (uint32_t s = 0; s < win.n_segs; ++s) {
            const kcf_segment_t &sg = segs[win.first_seg + s];
            if (sg.seq_id < 0 || (size_t)sg.seq_id >= ctx->seqs.size())
                return kcf_fail(ctx, KCF_ERR_RANGE, "Sequence not found in index: id %d (window %llu)", sg.seq_id, (unsigned long long)w);
            const KcfSeqHost &sq = ctx->seqs[sg.seq_id];
            const int64_t start = sg.start0, end = (int64_t)sg.start0 + sg.len;
            if (start < 0 || end > (int64_t)sq.len || start >= end) // FastaIndex.java:132-135
                return kcf_fail(ctx, KCF_ERR_RANGE, "Invalid range: %lld-%lld for sequence: id %d", (long long)start, (long long)end, sg.seq_id);
            // FastaIndex.java:169,175-177: after the last chunk the buffer position still moves past the line terminator
            const uint64_t last = (uint64_t)((end - 1) / sq.line_bases) * sq.line_width + (uint64_t)((end - 1) % sq.line_bases);
            if (last + 1 + (sq.line_width - sq.line_bases) > sq.n_bytes)
                return kcf_fail(ctx, KCF_ERR_FASTA, "Error reading sequence: id %d range %lld-%lld runs past the mapped bytes (missing trailing newline?)",
                                sg.seq_id, (long long)start, (long long)end);
            seg_off[win.first_seg + s] = (uint32_t)len;
            len += (uint64_t)sg.len;
        }
        if (len >= (1ULL << 31)) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "window %llu longer than 2^31 bases", (unsigned long long)w);
        win_len[w] = (uint32_t)len;
        tile_first[w] = tiles;
        tiles += (len + KCF_TILE - 1) / KCF_TILE;
        positions += len;
    }
    tile_first[n_wins] = tiles;
    kcf_plan *plan = new kcf_plan();
    plan->ctx = ctx;
    plan->k = kmer_length;
    plan->n_wins = n_wins;
    plan->n_segs = n_segs;
    plan->n_tiles = tiles;
    plan->n_positions = positions;
    plan->ref_generation = ctx->ref_generation;
    int rc = KCF_OK;
#define PL_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess && rc == KCF_OK)                                                         \
            rc = kcf_fail(ctx, e__ == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)
    {
        // one device block for all arrays of the plan (one cudaMalloc on the upload path instead of seven)
        auto up = [](uint64_t b) { return (b + 255) & ~255ULL; };
        const uint64_t nw = std::max<uint64_t>(n_wins, 1), ns = std::max<uint64_t>(n_segs, 1);
        const uint64_t b_wins = up(nw * sizeof(kcf_window_t)), b_segs = up(ns * sizeof(kcf_segment_t)), b_off = up(ns * 4), b_len = up(nw * 4),
                       b_tf = up((n_wins + 1) * 8), b_ts = up(std::max<uint64_t>(tiles, 1) * sizeof(KcfGap) + 8), // +8: the tile counter lives at the end
                       b_out = up(nw * sizeof(kcf_result_t)), b_flags = up(FLAG_COUNT * sizeof(uint32_t));
        const size_t total = b_wins + b_segs + b_off + b_len + b_tf + b_ts + b_out + b_flags;
        uint8_t *base = (uint8_t *)kcf_pool_get(ctx, total);
        if (!base && rc == KCF_OK) rc = kcf_fail(ctx, KCF_ERR_NOMEM, "device memory for a plan of %llu windows", (unsigned long long)n_wins);
        plan->d_block = base;
        plan->d_block_bytes = total;
        if (base) {
            plan->d_wins = reinterpret_cast<kcf_window_t *>(base);
            base += b_wins;
            plan->d_segs = reinterpret_cast<kcf_segment_t *>(base);
            base += b_segs;
            plan->d_seg_off = reinterpret_cast<uint32_t *>(base);
            base += b_off;
            plan->d_win_len = reinterpret_cast<uint32_t *>(base);
            base += b_len;
            plan->d_tile_first = reinterpret_cast<uint64_t *>(base);
            base += b_tf;
            plan->d_tile_sum = reinterpret_cast<KcfGap *>(base);
            base += b_ts;
            plan->d_out = reinterpret_cast<kcf_result_t *>(base);
            base += b_out;
            plan->d_flags = reinterpret_cast<uint32_t *>(base);
        }
    }
    if (rc == KCF_OK && n_wins) {
        // descriptors travel on their own stream: creating a plan waits neither for kernels queued on the main stream nor
        // for a sequence upload in flight on the copy stream (the host goes on to queue the next upload at once)
        PL_CUDA(cudaMemcpyAsync(plan->d_wins, wins, n_wins * sizeof(kcf_window_t), cudaMemcpyHostToDevice, ctx->desc_stream));
        PL_CUDA(cudaMemcpyAsync(plan->d_segs, segs, n_segs * sizeof(kcf_segment_t), cudaMemcpyHostToDevice, ctx->desc_stream));
        PL_CUDA(cudaMemcpyAsync(plan->d_seg_off, seg_off.data(), n_segs * 4, cudaMemcpyHostToDevice, ctx->desc_stream));
        PL_CUDA(cudaMemcpyAsync(plan->d_win_len, win_len.data(), n_wins * 4, cudaMemcpyHostToDevice, ctx->desc_stream));
    }
    if (rc == KCF_OK) PL_CUDA(cudaMemcpyAsync(plan->d_tile_first, tile_first.data(), (n_wins + 1) * 8, cudaMemcpyHostToDevice, ctx->desc_stream));
    if (rc == KCF_OK) PL_CUDA(cudaStreamSynchronize(ctx->desc_stream)); // host vectors go out of scope; caller may free wins/segs
#undef PL_CUDA
    if (rc != KCF_OK) {
        kcf_plan_destroy(plan);
        return rc;
    }
    plan->h_tile_first.swap(tile_first);
    plan->h_win_len.swap(win_len);
    *out = plan;
    return KCF_OK;
}

int kcf_launch_screen(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, int32_t min_count, uint64_t tile_begin, uint64_t tile_end,
                      int32_t *d_counts, bool extract, uint32_t *d_owned_hit, unsigned long long *d_owned_sum, const KcfXgDev *xsend,
                      cudaStream_t on_stream, uint32_t ctas_per_sm_cap)
{
    const cudaStream_t stream = on_stream ? on_stream : ctx->stream; // the exchange's send may run beside the answers of the batch before
    const bool owned = d_owned_hit != nullptr;
    if (xsend) extract = true;
    if (!extract && !owned && db->part_world > 1)
        return kcf_fail(ctx, KCF_ERR_ARG, "this database holds slice %d of %d: screen it through the exchange calls (kcf_xchg_*)", db->part_rank, db->part_world);
    if (plan->ref_generation != ctx->ref_generation)
        return kcf_fail(ctx, KCF_ERR_ARG, "plan was created before kcf_ref_clear: its segments refer to sequences that no longer exist");
    int rc = kcf_sync_seqs(ctx);
    if (rc != KCF_OK) return rc;
    unsigned long long *d_counter = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<uint8_t *>(plan->d_tile_sum) + std::max<uint64_t>(plan->n_tiles, 1) * sizeof(KcfGap));
    KCF_CUDA(ctx, cudaMemsetAsync(d_counter, 0, 8, stream));
    KcfScreenParams p{};
    p.seqs = ctx->d_seqs;
    p.wins = plan->d_wins;
    p.segs = plan->d_segs;
    p.seg_off = plan->d_seg_off;
    p.win_len = plan->d_win_len;
    p.tile_first = plan->d_tile_first;
    p.n_wins = plan->n_wins;
    p.tile_begin = tile_begin;
    p.tile_end = tile_end;
    p.table = db->table;
    p.stash = db->stash;
    p.tile_sum = plan->d_tile_sum;
    p.tile_counter = d_counter;
    p.min_count = min_count;
    p.counts_out = d_counts;
    p.counts_tile0 = tile_begin;
    p.x_keys = plan->x_keys;
    p.x_homes = plan->x_homes;
    p.x_okw = plan->x_okw;
    p.x_start = plan->x_start;
    if (xsend) p.xg = *xsend;
    if (owned) {
        p.x_okw = plan->s_okw;
        p.x_start = plan->s_start;
        p.x_hit = d_owned_hit;
        p.x_sum = d_owned_sum;
    }
    const size_t smem = 0; // the per-warp buffers are static shared memory
    void (*kern)(KcfScreenParams, KcfTableGeom);
    const int S = (int)db->geom.S;
    if (db->geom.kw == 2 && (extract || owned)) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "k=%u > 32 with a partitioned table", db->geom.k);
    // the common geometry (both-strands database, 4 <= w <= 13) runs the kernels that have it compiled in
    const bool spec = db->geom.both_strands && db->geom.w >= 4 && db->geom.w <= 13;
#define KCF_PICK_S(MODE, SP) (S == 13 ? kcf_screen_kernel<13, MODE, SP> : (S == 12 ? kcf_screen_kernel<12, MODE, SP> : kcf_screen_kernel<10, MODE, SP>))
#define KCF_PICK(MODE) (spec ? KCF_PICK_S(MODE, true) : KCF_PICK_S(MODE, false))
    if (db->geom.kw == 2) { // 128-bit keys: no minimizer, strandedness read at run time
        if (d_counts) kern = S == 7 ? kcf_screen_kernel<7, KCF_MODE_COUNTS, false> : kcf_screen_kernel<6, KCF_MODE_COUNTS, false>;
        else kern = S == 7 ? kcf_screen_kernel<7, KCF_MODE_SCREEN, false> : kcf_screen_kernel<6, KCF_MODE_SCREEN, false>;
    } else if (xsend) kern = KCF_PICK(KCF_MODE_XSEND);
    else if (extract) kern = KCF_PICK(KCF_MODE_EXTRACT);
    else if (owned) kern = KCF_PICK(KCF_MODE_OWNED);
    else if (d_counts) kern = KCF_PICK(KCF_MODE_COUNTS);
    else kern = KCF_PICK(KCF_MODE_SCREEN);
#undef KCF_PICK
#undef KCF_PICK_S
    int per_sm = 0;
    KCF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * KCF_WPC, smem));
    const uint64_t n = tile_end - tile_begin;
    if (ctas_per_sm_cap) per_sm = std::min<int>(per_sm, (int)ctas_per_sm_cap); // a persistent grid that leaves room for another kernel
    const unsigned grid = (unsigned)std::min<uint64_t>((n + KCF_WPC - 1) / KCF_WPC, (uint64_t)ctx->sm_count * std::max(per_sm, 1));
    if (grid) kern<<<grid, 32 * KCF_WPC, smem, stream>>>(p, db->geom);
    KCF_CUDA(ctx, cudaGetLastError());
    return KCF_OK;
}

extern "C" int kcf_plan_run(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, int32_t min_count, const double w[3])
{
    if (!ctx || !db || !plan || !w || plan->ctx != ctx || db->ctx != ctx) return KCF_ERR_ARG;
    if (min_count < 1) return kcf_fail(ctx, KCF_ERR_ARG, "Minimum kmer count should be at least 1"); // GetVariants.java:383-385
    if (plan->k != db->info.kmer_length) return kcf_fail(ctx, KCF_ERR_ARG, "plan built for k=%d, database has k=%d", plan->k, db->info.kmer_length);
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->d_flags, 0, FLAG_COUNT * sizeof(uint32_t), ctx->stream)); // per plan: several may be queued before a fetch
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], ctx->stream);
    int rc = kcf_launch_screen(ctx, db, plan, min_count, 0, plan->n_tiles, nullptr, false, nullptr, nullptr);
    if (rc != KCF_OK) return rc;
    if (ctx->profiling) cudaEventRecord(ctx->ev[1], ctx->stream);
    if (plan->n_wins) {
        kcf_finalize_kernel<<<(unsigned)((plan->n_wins + 127) / 128), 128, 0, ctx->stream>>>(
            plan->d_tile_sum, plan->d_tile_first, plan->n_wins, (uint32_t)plan->k, w[0], w[1], w[2], plan->d_out, plan->d_flags);
        KCF_CUDA(ctx, cudaGetLastError());
    }
    if (ctx->profiling) {
        cudaEventRecord(ctx->ev[2], ctx->stream);
        ctx->ev_valid = true;
    }
    plan->weights[0] = w[0];
    plan->weights[1] = w[1];
    plan->weights[2] = w[2];
    plan->ran = true;
    return KCF_OK;
}

// K5 alone, over tile summaries that were produced by kcf_xchg_fold (partitioned databases)
extern "C" int kcf_plan_finalize(kcf_ctx *ctx, kcf_plan *plan, const double w[3])
{
    if (!ctx || !plan || !w || plan->ctx != ctx) return KCF_ERR_ARG;
    if (plan->ref_generation != ctx->ref_generation) return kcf_fail(ctx, KCF_ERR_ARG, "plan was created before kcf_ref_clear");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    KCF_CUDA(ctx, cudaMemsetAsync(plan->d_flags, 0, FLAG_COUNT * sizeof(uint32_t), ctx->stream));
    if (plan->n_wins) {
        kcf_finalize_kernel<<<(unsigned)((plan->n_wins + 127) / 128), 128, 0, ctx->stream>>>(
            plan->d_tile_sum, plan->d_tile_first, plan->n_wins, (uint32_t)plan->k, w[0], w[1], w[2], plan->d_out, plan->d_flags);
        KCF_CUDA(ctx, cudaGetLastError());
    }
    plan->weights[0] = w[0];
    plan->weights[1] = w[1];
    plan->weights[2] = w[2];
    plan->ran = true;
    return KCF_OK;
}

extern "C" int kcf_plan_fetch(kcf_ctx *ctx, kcf_plan *plan, kcf_result_t *out)
{
    if (!ctx || !plan || plan->ctx != ctx || (!out && plan->n_wins)) return KCF_ERR_ARG;
    if (!plan->ran) return kcf_fail(ctx, KCF_ERR_ARG, "kcf_plan_fetch before kcf_plan_run");
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t flags[FLAG_COUNT] = {0};
    if (plan->n_wins) KCF_CUDA(ctx, cudaMemcpyAsync(out, plan->d_out, plan->n_wins * sizeof(kcf_result_t), cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaMemcpyAsync(flags, plan->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    KCF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // Data.java:101-103 — evaluated only for windows that reach the formula, left to right in double
    if (flags[FLAG_SCORE_USED] && plan->weights[0] + plan->weights[1] + plan->weights[2] != 1.0)
        return kcf_fail(ctx, KCF_ERR_WEIGHTS, "Weights should sum to 1.0");
    return KCF_OK;
}

extern "C" int kcf_plan_stats(kcf_plan *plan, uint64_t *n_tiles, uint64_t *n_positions, uint32_t *kernels_per_run)
{
    if (!plan) return KCF_ERR_ARG;
    if (n_tiles) *n_tiles = plan->n_tiles;
    if (n_positions) *n_positions = plan->n_positions;
    if (kernels_per_run) *kernels_per_run = plan->n_wins ? 2u : 0u;
    return KCF_OK;
}

extern "C" int kcf_screen(kcf_ctx *ctx, kcf_db *db, const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs,
                          uint64_t n_segs, int32_t min_count, const double w[3], kcf_result_t *out)
{
    if (!ctx || !db) return KCF_ERR_ARG;
    if (min_count < 1) return kcf_fail(ctx, KCF_ERR_ARG, "Minimum kmer count should be at least 1");
    kcf_plan *plan = nullptr;
    int rc = kcf_plan_create(ctx, db->info.kmer_length, wins, n_wins, segs, n_segs, &plan);
    if (rc != KCF_OK) return rc;
    rc = kcf_plan_run(ctx, db, plan, min_count, w);
    if (rc == KCF_OK) rc = kcf_plan_fetch(ctx, plan, out);
    kcf_plan_destroy(plan);
    return rc;
}

extern "C" int kcf_window_counts(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t window, int32_t *counts_out, uint64_t cap,
                                 uint64_t *n_out)
{
    if (!ctx || !db || !plan || !n_out || plan->ctx != ctx || window >= plan->n_wins) return KCF_ERR_ARG;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t t0 = plan->h_tile_first[window], t1 = plan->h_tile_first[window + 1];
    const uint64_t npos = (t1 - t0) * KCF_TILE;
    int32_t *d = nullptr;
    KCF_CUDA(ctx, cudaMalloc(&d, std::max<uint64_t>(npos, 1) * 4));
    int rc = kcf_launch_screen(ctx, db, plan, 1, t0, t1, d, false, nullptr, nullptr);
    std::vector<int32_t> h(npos);
    if (rc == KCF_OK) {
        cudaMemcpyAsync(h.data(), d, npos * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = kcf_fail(ctx, KCF_ERR_CUDA, "kcf_window_counts: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    if (rc != KCF_OK) return rc;
    uint64_t n = 0;
    const uint64_t wlen = plan->h_win_len[window];
    for (uint64_t i = 0; i < wlen && i < npos; ++i)
        if (h[i] != -1) { // -1 marks "no k-mer ends here" (ambiguous only for a 4-byte counter of 0xFFFFFFFF)
            if (n < cap && counts_out) counts_out[n] = h[i];
            ++n;
        }
    *n_out = n;
    return KCF_OK;
}
