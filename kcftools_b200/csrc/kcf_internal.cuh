// kcf_internal.cuh — shared host/device definitions of libkcfgpu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/kcf_b200.h"

// ------------------------------------------------------------------------------------------
// Device-resident database geometry.
//
// The reference finds a k-mer by signature -> bin -> prefix LUT -> binary search over 7-11 byte
// records (KMC.java:292-326): ~23 table lookups plus ~log2(range) dependent random reads.  Here
// the records are re-keyed once at load time into a table whose unit is one 128-byte line — the
// granularity at which B200's HBM serves a random access (profiles/README.md) — and whose home
// line is chosen by the k-mer's MINIMIZER, so that the ~(w+1)/2 consecutive reference k-mers
// sharing a minimizer probe the same line:
//
//   o(x)   = hash(smaller strand of x)         order hash of an m-mer, strand symmetric (kcf_lookup.cuh)
//   mu(K)  = min o(x) over the w = k-m+1 m-mers of K
//   home   = floor(mix32(mu ^ c) * n_lines / 2^32)
//   line   = S key low words | S key high words | filter | S counts | 16-bit mask
//            (S = 13 slots and a 64-bit filter for 1-byte counts; 12 / 10 slots and 32 bits for 2 / 4-byte counts;
//             key = the k-mer's bit planes, plane1 << 32 | plane0 (kcf_lookup.cuh), stored in full; low word
//             0xFFFFFFFF = empty slot)
//
// A key lives in its home line or, when that was full at insertion time (no deletions), in one of 14 lines of a
// separate OVERFLOW region starting at a hash of the home line — home lines only ever hold their own keys, so one
// crowded line does not push its neighbours' keys out.  Bit 0 of the HOME line's mask says "a key homed here lives
// here", bit d >= 1 "... in overflow line ov(home) + d - 1", bit 15 "... in the stash"; a lookup therefore knows after
// reading the home line exactly which other lines (if any) can hold its key.  The filter is a 2-hash Bloom filter over the keys homed
// here that live elsewhere: a k-mer that misses in its home line and fails the filter is absent, and no other line
// is read for it (absent k-mers are a quarter of a typical reference scan).  Keys are stored in full: a probe is
// exact; the filter can only cause extra probes, never a wrong answer.
// ------------------------------------------------------------------------------------------
#define KCF_LINE_BYTES 128
#define KCF_MAX_DISP 14
#define KCF_STASH_BIT 15
#define KCF_EMPTY_LO 0xFFFFFFFFu // low key word of a free slot
// keys whose low word equals the marker live in the stash
#define KCF_KEY_IN_LINES(key) ((uint32_t)(key) != KCF_EMPTY_LO)

struct KcfTableGeom {
    uint64_t n_lines;    // < 2^32 - 1; the GLOBAL number of home lines (home lines are computed against it)
    uint64_t line_lo;    // global index of local home line 0 (0 unless the database is partitioned)
    uint64_t n_local;    // home lines held by this table (== n_lines unless partitioned)
    uint64_t n_ov;       // overflow lines, stored after the home lines: keys that do not fit their home line
    uint64_t kmask;      // 2k one-bits (the reference's k-mer value)
    uint64_t stash_mask; // stash capacity - 1 (power of two), 0 when the stash is empty
    uint32_t k;
    uint32_t kshift;     // 64 - 2k
    uint32_t m;          // minimizer length, 1..24, <= k
    uint32_t w;          // k - m + 1 (1..32)
    uint32_t km;         // k one-bits (one bit plane of a k-mer)
    uint32_t mm;         // m one-bits (one bit plane of an m-mer)
    uint32_t S;          // slots per line: 13 / 12 / 10 for count width 1 / 2 / 4
    uint32_t cw;         // bytes per stored count: 1, 2 or 4 (0-byte counters store nothing)
    uint32_t coff;       // byte offset of the counts inside a line: 112 / 96 / 80
    uint32_t foff;       // byte offset of the filter: 104 (64 bits) / 120 (32 bits)
    uint32_t fbits;      // 64 or 32
    uint32_t both_strands;
    uint32_t kw;         // key width in 64-bit words: 1 (k <= 32) or 2 (k <= 64: two 64-bit planes, see below)
    uint64_t km64;       // k one-bits as a 64-bit plane mask (kw = 2)
};

// k-mers of 33 .. 64 bases (kw = 2): the key is two 64-bit planes (p0 = bit 0 of every base code, p1 = bit 1, base j in
// bit j) and a line holds S = 7 (1- and 2-byte counts) or 6 (4-byte counts) of them:
//   S low words (p0 & 0xFFFFFFFF; 0xFFFFFFFF = empty) | S x 3 remaining key words (p0 >> 32, p1 low, p1 high) | counts |
//   32-bit filter at byte 120 (absent for 2-byte counts: every miss then follows the mask) | 16-bit mask
// The home line is chosen by a hash of the whole key, not by a minimizer (m = w = 0): a run of consecutive k-mers would
// outnumber the slots of a line, so the locality trick of the short-k layout does not carry over.
struct KcfKey2 {
    uint64_t p0, p1;
};

struct KcfStashEntry {
    uint64_t key;    // table key (kw = 2: plane 0)
    uint64_t key_hi; // kw = 2: plane 1
    uint64_t meta;   // bit 63 = occupied, low 32 bits = count
};

// bijective 32-bit mixer (multiply / xorshift rounds)
__host__ __device__ __forceinline__ uint32_t kcf_mix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// the two filter bit positions of a key (6 bits each; the 32-bit filters use 5)
__host__ __device__ __forceinline__ uint32_t kcf_filter_hash(uint64_t key)
{
    return ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
}

// 64-bit mixer used for the stash only
__host__ __device__ __forceinline__ uint64_t kcf_mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// ------------------------------------------------------------------------------------------
// Reference sequences: two bit planes of the 2-bit base codes (A = 0, C = 1, G = 2, T = 3, Kmer.java:286-294) and a
// validity bitmap, 32 bases per word: codes[2i] = bit 0 of bases 32i .. 32i+31 (base j in bit j), codes[2i+1] = bit 1,
// valid[i] = base is one of ACGTacgt.
// ------------------------------------------------------------------------------------------
struct KcfSeqDev {
    const uint32_t *codes;
    const uint32_t *valid;
    uint32_t len;
    uint32_t _pad;
};

struct KcfSeqHost {
    uint32_t *codes = nullptr;
    uint32_t *valid = nullptr;
    uint64_t len = 0;
    uint64_t n_bytes = 0;
    uint32_t line_bases = 0, line_width = 0;
};

// ------------------------------------------------------------------------------------------
// The gap monoid: summary of a run of valid k-mers, each hit or miss, closed
// under in-order concatenation.  Restates the state machine of GetVariants.java:217-252.
// ------------------------------------------------------------------------------------------
struct KcfGap {
    uint32_t n;      // valid k-mers
    uint32_t obs;    // hits
    uint32_t lead;   // misses before the first hit (== n when there is no hit)
    uint32_t trail;  // misses after the last hit  (== n when there is no hit)
    uint32_t vin;    // interior miss runs (hit on both sides)
    uint32_t inner;  // Σ getDistance over interior miss runs
    uint32_t has;    // any hit
    uint32_t starts; // k-mers that open a valid stretch (for EFFLEN)
    uint64_t sum;    // Σ count over hits
};

struct kcf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::vector<KcfSeqHost> seqs;
    KcfSeqDev *d_seqs = nullptr; // mirror of seqs on the device
    KcfSeqDev *h_seqs = nullptr; // pinned staging of the same (entries are append-only between two kcf_ref_clear calls)
    size_t d_seqs_cap = 0;
    size_t seqs_uploaded = 0;
    bool seqs_dirty = false;
    // raw FASTA bytes are staged through two device buffers: the H2D copy of one sequence (copy stream) overlaps the
    // 2-bit packing of the previous one and whatever screening is queued on the main stream
    cudaStream_t copy_stream = nullptr;
    cudaStream_t desc_stream = nullptr; // window descriptors of new plans: never queued behind a sequence upload
    uint8_t *d_raw[2] = {nullptr, nullptr};
    size_t d_raw_cap[2] = {0, 0};
    cudaEvent_t raw_free[2] = {nullptr, nullptr}; // pack kernel done: staging buffer reusable
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    cudaEvent_t plan_ready = nullptr;             // window descriptors uploaded on the copy stream
    int raw_next = 0;
    struct PoolBlock {
        void *p;
        size_t bytes;
    };
    std::vector<PoolBlock> pool; // device blocks of cleared sequences, reused by later kcf_ref_add calls
    double load_factor = 0.0;    // 0 = automatic: 0.15, denser when the table would crowd the device memory
    int minimizer_len = 0;       // 0 = chosen from the database size
    int part_rank = 0, part_world = 1; // slice kept by databases opened with placement 1 (kcf_set_partition)
    int sm_count = 148;
    int profiling = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_screen_ms = 0.f, last_finalize_ms = 0.f;
    bool ev_valid = false;
    uint32_t *d_flags = nullptr; // device error / status flags (database load)
    // database ingest: ring of pinned host / device staging buffers, kept across kcf_db_open calls
    uint8_t *ing_h[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t *ing_d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ing_free[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // kernel that read slot j done
    cudaEvent_t ing_copied[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // H2D copy of slot j done
    size_t ing_slot_bytes = 0;
    cudaStream_t ing_stream = nullptr;                 // insert kernels: they may run under the proof kernel of the next chunk
    cudaEvent_t ing_proved[2] = {nullptr, nullptr};    // proof kernel that filled staged buffer j done
    cudaEvent_t ing_inserted[2] = {nullptr, nullptr};  // insert kernel that read staged buffer j done
    uint64_t piece_bases = 0;  // upload piece of a sharded job, 0 = default (kcf_set_upload_piece)
    uint8_t *h_rows = nullptr; // pinned landing area for the rows of a sharded job (kcf_multi.cu): all plans fetched with one wait
    size_t h_rows_cap = 0;
    uint64_t ref_generation = 1; // bumped by kcf_ref_clear: plans remember the generation they were built against
};

struct kcf_db {
    kcf_ctx *ctx = nullptr;
    int part_rank = 0, part_world = 1; // > 1: this table holds one slice of the line space (placement 1)
    kcf_db_info_t info{};
    KcfTableGeom geom{};
    uint8_t *table = nullptr;        // n_lines * 128 bytes, from the context's block pool
    size_t table_bytes = 0;
    KcfStashEntry *stash = nullptr;  // stash_mask + 1 entries
};

struct kcf_plan {
    kcf_ctx *ctx = nullptr;
    int32_t k = 0;
    uint64_t n_wins = 0, n_segs = 0, n_tiles = 0, n_positions = 0;
    void *d_block = nullptr;          // the one device block (from the context's pool) holding the arrays below
    size_t d_block_bytes = 0;
    kcf_window_t *d_wins = nullptr;
    kcf_segment_t *d_segs = nullptr;
    uint32_t *d_seg_off = nullptr;    // offset of each segment inside its window
    uint32_t *d_win_len = nullptr;    // Σ segment lengths per window
    uint64_t *d_tile_first = nullptr; // n_wins + 1: first tile of each window
    KcfGap *d_tile_sum = nullptr;     // one summary per tile
    kcf_result_t *d_out = nullptr;
    uint32_t *d_flags = nullptr;      // this plan's status word(s): FLAG_SCORE_USED of ITS last run (several plans may be queued)
    uint64_t ref_generation = 0;      // ctx->ref_generation at creation: the sequences its segments point into
    std::vector<uint32_t> h_win_len;
    std::vector<uint64_t> h_tile_first;
    bool ran = false;
    double weights[3] = {0, 0, 0};
    // scratch of the exchange path (partitioned databases), sized for one batch of tiles
    uint64_t x_cap = 0;                    // positions
    unsigned long long *x_keys = nullptr;  // canonical k-mer of every position of the batch
    uint32_t *x_homes = nullptr;           // its global home line (0xFFFFFFFF: no k-mer ends here)
    uint32_t *x_okw = nullptr, *x_start = nullptr; // validity / stretch-start bitmaps, one word per 32 positions
    uint32_t *x_cnt = nullptr;             // counts returned by the owners, by position
    unsigned long long *x_cursor = nullptr; // per-owner counters (device)
    // scratch of the scan path (partitioned databases, every rank walks every tile): bitmaps, one word per 32 positions
    uint64_t s_cap = 0; // words
    uint32_t *s_okw = nullptr, *s_start = nullptr;
};

// ------------------------------------------------------------------------------------------
// k-mer exchange over peer memory (partitioned table).  Every rank owns one workspace block that its
// peers map (CUDA IPC across processes, plain pointers inside one): an INBOX with one region per sender and a BACK
// area with one region per owner.  The unit on the wire is not a k-mer but a RUN: up to 11 consecutive k-mers of a
// window that share their home line (they share their minimizer, which is what picks the home line and therefore the
// owner).  A run travels as the k + len - 1 bases it spans — two bit planes of at most 41 bits — plus its length and
// home line, 16 bytes in all (2.7 bytes per k-mer at the average run of 6, against 12 for key + home line per k-mer);
// the owner fetches the run's line once, cuts the len k-mers out of the planes, and stores their counts as one 16-byte
// slot (1-byte counters) into the requester's BACK region.  The screening kernel appends straight into the owners'
// inboxes over NVLink (no local copy, no grouping pass), and the requester folds through the run slot it noted per
// position.  Nothing crosses the host between the three kernels of a batch but two barriers.
// ------------------------------------------------------------------------------------------
#define KCF_XG_MAX_WORLD 16
#define KCF_XG_RUN 11          // k-mers per run at most: k + 10 <= 42 bases fit the entry's planes for k <= 32
#define KCF_XG_SLAB 32u        // entries a warp reserves in an owner's region at a time (>= the runs of one warp step)
struct KcfXgDev {
    uint32_t world, me, cbytes, stride;  // cbytes: bytes per count on the wire (1 for 1-byte counters, else 4); stride: bytes per BACK slot (16 / 48)
    uint64_t cap;                        // runs per (sender, owner) region
    uint4 *in_runs[KCF_XG_MAX_WORLD];    // [o]: rank o's inbox, region of sender `me`
    uint32_t *in_count[KCF_XG_MAX_WORLD];// [o]: where rank o reads how many runs `me` sent
    uint8_t *back[KCF_XG_MAX_WORLD];     // [s]: rank s's BACK area, region of owner `me`
    const uint4 *my_runs;                // this rank's inbox (regions indexed by sender)
    const uint32_t *my_count;
    const uint8_t *my_back;              // this rank's BACK area (regions indexed by owner)
    uint32_t *pos_slot;                  // per position of the batch: run head: owner << 28 | run index; 0xFFFFFFFE: member of the run that
                                         // heads to its left; 0xFFFFFFFF: no k-mer ends here
    uint32_t *okw, *start;               // validity / stretch-start bitmaps of the batch
    unsigned int *cursor;                // [world]: runs appended per owner so far
    uint32_t *flags;                     // [0]: a region overflowed
    // owner of a home line without the 64-bit division of kcf_line_owner: own_mul = floor(2^32 world / n_lines) gives an
    // estimate that is never too high, own_bound[r] = the first home line of rank r's slice corrects it
    uint32_t own_mul, own_bound[KCF_XG_MAX_WORLD + 1];
};

// one run on the wire: a = plane 0 (k + len - 1 bits) | (len - 1) << 42 | (home & 0x3FFFF) << 46, b = plane 1 | (home >> 18) << 42
__host__ __device__ __forceinline__ void kcf_xg_pack_run(uint64_t p0, uint64_t p1, uint32_t len, uint32_t home, uint64_t &a, uint64_t &b)
{
    a = p0 | ((uint64_t)(len - 1u) << 42) | ((uint64_t)(home & 0x3FFFFu) << 46);
    b = p1 | ((uint64_t)(home >> 18) << 42);
}
__host__ __device__ __forceinline__ void kcf_xg_unpack_run(uint64_t a, uint64_t b, uint64_t &p0, uint64_t &p1, uint32_t &len, uint32_t &home)
{
    p0 = a & ((1ULL << 42) - 1ULL);
    p1 = b & ((1ULL << 42) - 1ULL);
    len = (uint32_t)((a >> 42) & 15u) + 1u;
    home = (uint32_t)(a >> 46) | ((uint32_t)(b >> 42) << 18);
}

#define KCF_TILE 2048          // positions per tile = the unit of work one warp takes
#define KCF_HALO 64            // bases staged before a chunk (>= k-1 + the minimizer window, word aligned)

// indices into kcf_ctx::d_flags
enum { FLAG_LUT_BAD = 0, FLAG_ORDER_BAD = 1, FLAG_SCORE_USED = 2, FLAG_COUNT = 8 };

int kcf_fail(kcf_ctx *ctx, int code, const char *fmt, ...);
// device blocks are recycled through the context (cudaMalloc / cudaFree are slow, far slower once a communication library
// has enabled peer access, and cudaFree synchronises the device): sequences and plans take their memory from here
void *kcf_pool_get(kcf_ctx *ctx, size_t bytes);
void kcf_pool_put(kcf_ctx *ctx, void *p, size_t bytes);
void kcf_pool_trim(kcf_ctx *ctx); // free every pooled block (called when a large allocation fails)
int kcf_sync_seqs(kcf_ctx *ctx); // uploads the sequence table when it changed (on the context's stream)
int kcf_launch_screen(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, int32_t min_count, uint64_t tile_begin, uint64_t tile_end,
                      int32_t *d_counts, bool extract, uint32_t *d_owned_hit, unsigned long long *d_owned_sum, const KcfXgDev *xsend = nullptr,
                      cudaStream_t on_stream = nullptr, uint32_t ctas_per_sm_cap = 0);
#define KCF_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return kcf_fail((ctx), e__ == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, \
                            "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
