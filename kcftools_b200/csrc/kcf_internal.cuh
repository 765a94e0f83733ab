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
// the records are re-keyed once at load time into an open-addressing table whose unit is one
// 32-byte DRAM sector:
//
//   h      = mix(canonical k-mer)            bijection on 2k bits
//   bucket = floor(h * n_buckets / 4^k)      home bucket
//   slot   = [1 | disp:3 | rem:r | count:cbits]   64 bits, 4 slots per bucket
//
// rem = low r bits of h; the h values of one home bucket form an interval shorter than 2^r, so
// (home bucket, rem) identifies h and therefore the k-mer.  An entry that does not fit its home
// bucket goes to bucket home+disp (disp <= 7); beyond that to a small stash.  A probe stops at the
// first bucket with a free slot (there are no deletions).
// ------------------------------------------------------------------------------------------
#define KCF_SLOTS_PER_BUCKET 4
#define KCF_DISP_BITS 3
#define KCF_MAX_DISP 7

struct KcfTableGeom {
    uint64_t n_buckets;
    uint64_t kmask;      // 2k one-bits
    uint64_t rmask;      // r one-bits
    uint64_t cmask;      // cbits one-bits
    uint64_t stash_mask; // stash capacity - 1 (power of two), 0 when the stash is empty
    uint32_t k;
    uint32_t kshift;     // 64 - 2k
    uint32_t rbits;
    uint32_t cbits;
    uint32_t both_strands;
    uint32_t s1, s2;     // xor-shift distances of the mixer
};

struct KcfStashEntry {
    uint64_t key;   // canonical k-mer value (right aligned)
    uint64_t meta;  // bit 63 = occupied, low 32 bits = count
};

// bijective mixer on 2k-bit values: xorshift / odd multiply rounds, all modulo 2^(2k)
__host__ __device__ __forceinline__ uint64_t kcf_mix(uint64_t x, const KcfTableGeom &g)
{
    x ^= x >> g.s1;
    x = (x * 0xff51afd7ed558ccdULL) & g.kmask;
    x ^= x >> g.s2;
    x = (x * 0xc4ceb9fe1a85ec53ULL) & g.kmask;
    x ^= x >> g.s1;
    return x;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t kcf_home_bucket(uint64_t h, const KcfTableGeom &g)
{
    return __umul64hi(h << g.kshift, g.n_buckets);
}
#endif

// ------------------------------------------------------------------------------------------
// Reference sequences: 2-bit codes (16 bases per u32, base j of a word in bits 2j..2j+1) and a
// validity bitmap (32 bases per u32).  A = 0, C = 1, G = 2, T = 3 (Kmer.java:286-294).
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
    size_t d_seqs_cap = 0;
    bool seqs_dirty = false;
    uint8_t *d_raw = nullptr;    // staging for raw FASTA bytes
    size_t d_raw_cap = 0;
    double load_factor = 0.5;
    int sm_count = 148;
    int profiling = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_screen_ms = 0.f, last_finalize_ms = 0.f;
    bool ev_valid = false;
    uint32_t *d_flags = nullptr; // device error / status flags
};

struct kcf_db {
    kcf_ctx *ctx = nullptr;
    kcf_db_info_t info{};
    KcfTableGeom geom{};
    uint64_t *table = nullptr;       // n_buckets * 4 slots
    KcfStashEntry *stash = nullptr;  // stash_mask + 1 entries
};

struct kcf_plan {
    kcf_ctx *ctx = nullptr;
    int32_t k = 0;
    uint64_t n_wins = 0, n_segs = 0, n_tiles = 0, n_positions = 0;
    kcf_window_t *d_wins = nullptr;
    kcf_segment_t *d_segs = nullptr;
    uint32_t *d_seg_off = nullptr;    // offset of each segment inside its window
    uint32_t *d_win_len = nullptr;    // Σ segment lengths per window
    uint64_t *d_tile_first = nullptr; // n_wins + 1: first tile of each window
    KcfGap *d_tile_sum = nullptr;     // one summary per tile
    kcf_result_t *d_out = nullptr;
    std::vector<uint32_t> h_win_len;
    std::vector<uint64_t> h_tile_first;
    bool ran = false;
    double weights[3] = {0, 0, 0};
};

#define KCF_TILE 2048          // positions per tile (256 threads x 8)
#define KCF_THREADS 256
#define KCF_PER_THREAD 8
#define KCF_HALO 32            // bases staged before the tile (>= k-1, word aligned)

// indices into kcf_ctx::d_flags
enum { FLAG_LUT_BAD = 0, FLAG_ORDER_BAD = 1, FLAG_SCORE_USED = 2, FLAG_COUNT = 8 };

int kcf_fail(kcf_ctx *ctx, int code, const char *fmt, ...);
#define KCF_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return kcf_fail((ctx), e__ == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, \
                            "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
